"""DBRX (reference models/dbrx/modeling_dbrx.py:1-308): fused Wqkv with ``clip_qkv``, bias-free LayerNorm, MoE v1
(softmax router, top-4 of 16, renormalised), GLU experts stored as w1/v1/w2 ``[E*I, H]``."""
from __future__ import annotations

import torch
import torch.nn as nn

from ...config import InferenceConfig, MoENeuronConfig
from ...modules.attention import AttentionBase
from ...modules.moe import initialize_moe_module
from ...modules.norm import LayerNorm
from ...modules.rope import RotaryEmbedding
from ...parallel.layers import ColumnParallelLinear, ParallelEmbedding
from ..application_base import NeuronBaseForCausalLM
from ..model_base import NeuronBaseModel


class DbrxInferenceConfig(InferenceConfig):
    def get_required_attributes(self):
        return ["d_model", "n_heads", "n_layers", "vocab_size", "max_seq_len", "attn_config", "ffn_config"]

    def add_derived_config(self):
        self.num_cores_per_group = 1
        ac, fc = _as_dict(self.attn_config), _as_dict(self.ffn_config)
        self.hidden_size, self.num_attention_heads, self.num_hidden_layers = self.d_model, self.n_heads, self.n_layers
        self.num_key_value_heads = ac.get("kv_n_heads", self.n_heads)
        self.head_dim = self.d_model // self.n_heads
        self.clip_qkv = ac.get("clip_qkv")
        self.rope_theta = ac.get("rope_theta", 10000.0)
        self.intermediate_size = fc.get("ffn_hidden_size")
        self.num_local_experts = fc.get("moe_num_experts")
        self.num_experts_per_tok = fc.get("moe_top_k")
        self.hidden_act = (fc.get("ffn_act_fn") or {}).get("name", "silu")
        self.max_position_embeddings = self.max_seq_len
        self.rms_norm_eps = 1e-5

    @classmethod
    def get_neuron_config_cls(cls):
        return MoENeuronConfig


def _as_dict(x):
    return x if isinstance(x, dict) else (x.to_dict() if hasattr(x, "to_dict") else vars(x))


class DbrxBlock(nn.Module):
    mlp_is_moe = True

    @property
    def mlp(self):
        return self.ffn

    def __init__(self, config, i, rotary, device=None):
        super().__init__()
        dt = config.neuron_config.torch_dtype
        self.norm_1 = LayerNorm(config.hidden_size, 1e-5, bias=False, dtype=dt, device=device)
        self.norm_2 = LayerNorm(config.hidden_size, 1e-5, bias=False, dtype=dt, device=device)
        self.self_attn = AttentionBase(config, hidden_size=config.hidden_size, num_attention_heads=config.num_attention_heads,
                                       num_key_value_heads=config.num_key_value_heads, head_dim=config.head_dim,
                                       rotary_emb=rotary, clip_qkv=config.clip_qkv, layer_idx=i, device=device)
        self.ffn = initialize_moe_module(config, device=device)

    def forward(self, h, meta, kv_mgr, lora=None):
        h = self.self_attn(self.norm_1(h), meta, kv_mgr, residual=h)
        return self.ffn(self.norm_2(h), residual=h)


class NeuronDbrxModel(NeuronBaseModel):
    graph_safe = False            # the torch expert dispatch synchronises ...
    moe_decode_graph_safe = True  # ... but decode (T <= 8) runs the moe_decode kernels: CUDA graphs allowed when they apply

    def setup_attr_for_model(self, config):
        nc = config.neuron_config
        self.tp_degree, self.hidden_size = nc.tp_degree, config.hidden_size
        self.num_attention_heads, self.num_key_value_heads = config.num_attention_heads, config.num_key_value_heads
        self.max_batch_size, self.buckets = nc.max_batch_size, nc.buckets

    def init_model(self, config):
        nc = config.neuron_config
        dev, dt = self.device_, nc.torch_dtype
        self.embed_tokens = ParallelEmbedding(config.vocab_size, config.hidden_size, None, dtype=dt, device=dev,
                                              shard_across_embedding=not nc.vocab_parallel, pad=True,
                                              tensor_model_parallel_group=self.tp_group)
        rotary = RotaryEmbedding(config.head_dim, max(config.max_seq_len, nc.seq_len), config.rope_theta, device=dev)
        self.layers = nn.ModuleList([DbrxBlock(config, i, rotary, dev) for i in range(config.num_hidden_layers)])
        self.norm = LayerNorm(config.hidden_size, 1e-5, bias=False, dtype=dt, device=dev)
        self.lm_head = ColumnParallelLinear(config.hidden_size, config.vocab_size, bias=False, gather_output=False, dtype=dt,
                                            device=dev, pad=True, tensor_model_parallel_group=self.tp_group)

    def compute_logits(self, h):
        return self.lm_head(self.norm(h))


class NeuronDbrxForCausalLM(NeuronBaseForCausalLM):
    _model_cls = NeuronDbrxModel
    _STATE_DICT_MODEL_PREFIX = "transformer."

    @classmethod
    def get_config_cls(cls):
        return DbrxInferenceConfig

    @staticmethod
    def convert_hf_to_neuron_state_dict(state_dict, config):
        """reference :51-112."""
        E, I, H = config.num_local_experts, config.intermediate_size, config.hidden_size
        out = {}
        for k, v in state_dict.items():
            k = k.replace("blocks.", "layers.")
            if k == "wte.weight":
                out["embed_tokens.weight"] = v
            elif k == "norm_f.weight":
                out["norm.weight"] = v
            elif ".norm_attn_norm.norm_1." in k:
                out[k.replace(".norm_attn_norm.norm_1.", ".norm_1.")] = v
            elif ".norm_attn_norm.norm_2." in k:
                out[k.replace(".norm_attn_norm.norm_2.", ".norm_2.")] = v
            elif ".norm_attn_norm.attn.Wqkv." in k:
                out[k.replace(".norm_attn_norm.attn.Wqkv.", ".self_attn.qkv_proj.")] = v
            elif ".norm_attn_norm.attn.out_proj." in k:
                out[k.replace(".norm_attn_norm.attn.out_proj.", ".self_attn.o_proj.")] = v
            elif k.endswith(".ffn.router.layer.weight"):
                out[k.replace(".ffn.router.layer.weight", ".ffn.router.linear_router.weight")] = v.float()
            elif k.endswith(".ffn.experts.mlp.w1"):
                base = k[: -len(".ffn.experts.mlp.w1")]
                w1 = v.view(E, I, H)
                v1 = state_dict[base.replace("layers.", "blocks.") + ".ffn.experts.mlp.v1"].view(E, I, H)
                w2 = state_dict[base.replace("layers.", "blocks.") + ".ffn.experts.mlp.w2"].view(E, I, H)
                out[base + ".ffn.expert_mlps.gate_up_proj"] = torch.cat([w1, v1], 1).contiguous()   # [E, 2I, H]
                out[base + ".ffn.expert_mlps.down_proj"] = w2.transpose(1, 2).contiguous()           # [E, H, I]
            elif k.endswith(".ffn.experts.mlp.v1") or k.endswith(".ffn.experts.mlp.w2"):
                continue
            else:
                out[k] = v
        return out
