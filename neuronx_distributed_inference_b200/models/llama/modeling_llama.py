"""Llama family (Llama-2/3/3.1/3.2/3.3, TinyLlama, open_llama ...).

reference: models/llama/modeling_llama.py:1-1378.  Per layer: fused-QKV GQA attention with RoPE
(llama3 frequency scaling), SwiGLU MLP with fused gate/up, RMSNorm; vocab-parallel lm_head; optional
tied embeddings.  The reference's NKI kernel toggles (qkv/mlp/attn kernels, fused residual add,
skip-gamma folding) have no analogue: on B200 the fused kernels are simply what the layers run.
"""
from __future__ import annotations

from typing import List, Type

import torch
import torch.nn as nn

from ...config import InferenceConfig, NeuronConfig
from ...modules.attention import AttentionBase
from ...modules.mlp import GatedMLP
from ...modules.norm import IdentityNorm, RMSNorm
from ...modules.rope import RotaryEmbedding
from ...parallel.layers import ColumnParallelLinear, ParallelEmbedding
from ..application_base import NeuronBaseForCausalLM
from ..model_base import DecoderLayer, NeuronBaseModel
from ..state_dict_utils import fuse_qkv_and_gate_up


class LlamaInferenceConfig(InferenceConfig):
    def get_required_attributes(self) -> List[str]:
        return ["hidden_size", "num_attention_heads", "num_hidden_layers", "num_key_value_heads",
                "vocab_size", "max_position_embeddings", "rms_norm_eps", "intermediate_size"]

    def add_derived_config(self):
        self.num_cores_per_group = 1
        if not hasattr(self, "head_dim") or self.head_dim is None:
            self.head_dim = self.hidden_size // self.num_attention_heads
        if not hasattr(self, "hidden_act"):
            self.hidden_act = "silu"
        nc_ = self.neuron_config
        if nc_.flash_decoding_enabled or nc_.attention_dp_degree > 1 or nc_.cp_degree > 1:
            from ...modules.flashdecode import calculate_num_cores_per_group
            self.num_cores_per_group = calculate_num_cores_per_group(
                self.num_attention_heads, self.num_key_value_heads, self.neuron_config.tp_degree)

    @classmethod
    def get_neuron_config_cls(cls) -> Type[NeuronConfig]:
        return NeuronConfig


def rope_scaling_of(config):
    rs = getattr(config, "rope_scaling", None)
    if rs is None:
        rp = getattr(config, "rope_parameters", None)
        if isinstance(rp, dict) and rp.get("rope_type", "default") != "default":
            rs = rp
    return rs


def rope_theta_of(config, default=10000.0):
    th = getattr(config, "rope_theta", None)
    if th is None:
        rp = getattr(config, "rope_parameters", None)
        if isinstance(rp, dict):
            th = rp.get("rope_theta")
    return float(th) if th is not None else default


class NeuronLlamaAttention(AttentionBase):
    def __init__(self, config, layer_idx: int, rotary_emb, device=None, **over):
        kw = dict(hidden_size=config.hidden_size, num_attention_heads=config.num_attention_heads,
                  num_key_value_heads=config.num_key_value_heads, head_dim=config.head_dim, rotary_emb=rotary_emb,
                  qkv_bias=getattr(config, "attention_bias", False), o_bias=getattr(config, "attention_bias", False),
                  layer_idx=layer_idx, rms_norm_eps=config.rms_norm_eps, device=device)
        kw.update(over)
        super().__init__(config, **kw)


class NeuronLlamaMLP(GatedMLP):
    def __init__(self, config, device=None):
        nc = config.neuron_config
        super().__init__(config.hidden_size, config.intermediate_size, config.hidden_act, nc.torch_dtype,
                         bias=getattr(config, "mlp_bias", False), device=device,
                         sequence_parallel_enabled=nc.sequence_parallel_enabled, reduce_dtype=nc.rpl_reduce_dtype)


class ResBlock(nn.Module):
    """Medusa residual block ``x + silu(W x + b)`` (reference modeling_llama.py:1059-1095)."""

    def __init__(self, hidden_size, dtype, device=None):
        super().__init__()
        self.linear = nn.Linear(hidden_size, hidden_size, dtype=dtype, device=device)
        nn.init.zeros_(self.linear.weight)
        for p in self.parameters():
            p.requires_grad_(False)

    def forward(self, x):
        return x + torch.nn.functional.silu(self.linear(x))


class MedusaHead(nn.Module):
    """``num_layers`` ResBlocks followed by a vocab-parallel projection (reference modeling_llama.py:1172-1187)."""

    def __init__(self, config, tp_group, device=None, num_layers: int = 1):
        super().__init__()
        dt = config.neuron_config.torch_dtype
        self.blocks = nn.ModuleList([ResBlock(config.hidden_size, dt, device) for _ in range(num_layers)])
        self.proj = ColumnParallelLinear(config.hidden_size, config.vocab_size, bias=False, gather_output=False, dtype=dt,
                                         device=device, pad=True, tensor_model_parallel_group=tp_group)

    def forward(self, x):
        for b in self.blocks:
            x = b(x)
        return self.proj(x)


class Eagle3DecoderLayer(DecoderLayer):
    """EAGLE-3 draft block: attention reads ``[norm(embedding) | norm(feature)]`` (2H wide), the residual stream is the
    feature half (reference modeling_llama.py:917-922, 974-1001)."""

    def __init__(self, attn, mlp, input_layernorm, post_attention_layernorm, hidden_norm, layer_idx=0):
        super().__init__(attn, mlp, input_layernorm, post_attention_layernorm, layer_idx)
        self.hidden_norm = hidden_norm

    def forward(self, h2, meta, kv_mgr, lora=None):
        H = h2.shape[-1] // 2
        emb, feat = h2[..., :H], h2[..., H:]
        x = torch.cat([self.input_layernorm(emb), self.hidden_norm(feat)], -1)
        h = self.self_attn(x, meta, kv_mgr, residual=feat.contiguous())
        n2 = self.post_attention_layernorm
        return self.mlp(h, norm_weight=n2.weight, norm_eps=n2.variance_epsilon, norm_offset=n2.offset, residual=h)


class NeuronLlamaModel(NeuronBaseModel):
    attention_cls = NeuronLlamaAttention
    mlp_cls = NeuronLlamaMLP

    def setup_attr_for_model(self, config):
        nc = config.neuron_config
        self.tp_degree = nc.tp_degree
        self.hidden_size = config.hidden_size
        self.num_attention_heads = config.num_attention_heads
        self.num_key_value_heads = config.num_key_value_heads
        self.max_batch_size = nc.max_batch_size
        self.buckets = nc.buckets

    def make_rotary(self, config, device):
        rot_dim = int(config.head_dim * getattr(config, "partial_rotary_factor", 1.0))
        return RotaryEmbedding(rot_dim, max(config.max_position_embeddings, config.neuron_config.seq_len),
                               rope_theta_of(config), rope_scaling_of(config), device=device)

    def make_layer(self, config, i, rotary, device):
        nc = config.neuron_config
        if nc.is_eagle_draft and nc.is_eagle3:
            norm = lambda: RMSNorm(config.hidden_size, config.rms_norm_eps, nc.torch_dtype, device=device)  # noqa: E731
            return Eagle3DecoderLayer(self.attention_cls(config, i, rotary, device=device, qkv_input_size=2 * config.hidden_size),
                                      self.mlp_cls(config, device=device), norm(), norm(), norm(), i)
        if nc.is_eagle_draft and not nc.enable_eagle_draft_input_norm and i == 0:
            return DecoderLayer(self.attention_cls(config, i, rotary, device=device), self.mlp_cls(config, device=device),
                                IdentityNorm(), RMSNorm(config.hidden_size, config.rms_norm_eps, nc.torch_dtype, device=device), i)
        return DecoderLayer(self.attention_cls(config, i, rotary, device=device), self.mlp_cls(config, device=device),
                            RMSNorm(config.hidden_size, config.rms_norm_eps, nc.torch_dtype, device=device),
                            RMSNorm(config.hidden_size, config.rms_norm_eps, nc.torch_dtype, device=device), i)

    def init_model(self, config):
        nc = config.neuron_config
        dev, dt = self.device_, nc.torch_dtype
        self.embed_tokens = ParallelEmbedding(config.vocab_size, config.hidden_size, getattr(config, "pad_token_id", None),
                                              dtype=dt, device=dev, shard_across_embedding=not nc.vocab_parallel,
                                              pad=True, tensor_model_parallel_group=self.tp_group)
        rotary = self.make_rotary(config, dev)
        self.rotary_emb = rotary
        self.layers = nn.ModuleList([self.make_layer(config, i, rotary, dev) for i in range(config.num_hidden_layers)])
        eagle1_draft = nc.is_eagle_draft and not nc.is_eagle3
        self.norm = IdentityNorm() if eagle1_draft else RMSNorm(config.hidden_size, config.rms_norm_eps, dt, device=dev)
        out_vocab = getattr(config, "draft_vocab_size", None) if nc.is_eagle_draft else None
        self.lm_head = ColumnParallelLinear(config.hidden_size, out_vocab or config.vocab_size, bias=False,
                                            gather_output=False, dtype=dt, device=dev, pad=True,
                                            tensor_model_parallel_group=self.tp_group)
        if nc.is_eagle_draft:
            # feature fusion: [embedding | target feature] -> H (EAGLE-1/2) or [low | mid | high] -> H (EAGLE-3).  The
            # reference gathers this weight on the fly (WeightGatheredColumnParallel, modeling_llama.py:1161-1167); at
            # 2H x H it is 64 MB for an 8B model, so on B200 it is simply replicated.
            n_in = 3 if nc.is_eagle3 else 2
            self.fc = nn.Linear(n_in * config.hidden_size, config.hidden_size, bias=bool(getattr(config, "fc_bias", False)),
                                dtype=dt, device=dev)
            for p in self.fc.parameters():
                p.requires_grad_(False)
            if out_vocab:
                self.register_buffer("d2t", torch.zeros(out_vocab, dtype=torch.long, device=dev))
        if nc.is_eagle3 and not nc.is_eagle_draft:
            L = config.num_hidden_layers
            self.aux_hidden_layers = tuple(getattr(config, "eagle3_aux_layers", None) or (min(2, L - 1), L // 2, max(L - 3, 0)))
        if nc.is_medusa:
            self.medusa_heads = nn.ModuleList([MedusaHead(config, self.tp_group, dev, getattr(config, "medusa_num_layers", 1))
                                               for _ in range(nc.num_medusa_heads)])

    # ---- EAGLE draft hooks -----------------------------------------------------------------------------------------
    def fuse_prev_hidden(self, emb, prev_hidden):
        """EAGLE-1/2: ``fc([emb | feature])``.  EAGLE-3: features coming from the target are 3H wide and go through ``fc``
        first; the draft's own features (H wide) are used as they are; the layer then consumes ``[emb | feature]``."""
        nc = self.neuron_config
        prev_hidden = prev_hidden.to(emb.dtype)
        if not nc.is_eagle3:
            return self.fc(torch.cat([emb, prev_hidden], -1))
        if prev_hidden.shape[-1] == 3 * self.hidden_size:
            prev_hidden = self.fc(prev_hidden)
        return torch.cat([emb, prev_hidden], -1)

    def final_hidden(self, h):
        nc = self.neuron_config
        if nc.is_eagle_draft and nc.is_eagle3:
            return h            # EAGLE-3 feeds the pre-norm state back; ``norm`` is applied only in front of lm_head
        return self.norm(h)

    def map_draft_tokens(self, tokens):
        """EAGLE-3 reduced draft vocabulary: target id = draft id + d2t[draft id]."""
        d2t = getattr(self, "d2t", None)
        return tokens if d2t is None else tokens + d2t[tokens]


class NeuronLlamaForCausalLM(NeuronBaseForCausalLM):
    _model_cls = NeuronLlamaModel

    @classmethod
    def get_config_cls(cls):
        return LlamaInferenceConfig

    @staticmethod
    def convert_hf_to_neuron_state_dict(state_dict: dict, config) -> dict:
        nc = config.neuron_config
        if nc.is_medusa:
            # Medusa checkpoints: medusa_head.<i>.<j>.linear.{weight,bias}, medusa_head.<i>.<n>.weight (vocab projection)
            n_blocks = getattr(config, "medusa_num_layers", 1)
            for k in [k for k in state_dict if k.startswith("medusa_head.")]:
                parts = k.split(".")
                i, j = int(parts[1]), int(parts[2])
                new = f"medusa_heads.{i}.proj.{parts[-1]}" if j == n_blocks else f"medusa_heads.{i}.blocks.{j}." + ".".join(parts[3:])
                state_dict[new] = state_dict.pop(k)
        if nc.is_eagle_draft:
            # EAGLE drafts ship a single "midlayer" (EAGLE-3) or layers.* without embeddings / lm_head of their own
            for k in [k for k in state_dict if k.startswith("midlayer.")]:
                state_dict["layers.0." + k[len("midlayer."):]] = state_dict.pop(k)
        return fuse_qkv_and_gate_up(state_dict, config.num_hidden_layers)

    @staticmethod
    def update_state_dict_for_tied_weights(state_dict):
        state_dict["lm_head.weight"] = state_dict["embed_tokens.weight"].clone()
