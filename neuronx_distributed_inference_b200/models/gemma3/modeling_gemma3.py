"""Gemma-3 text (reference models/gemma3/modeling_gemma3.py:1-424): offset RMSNorm (1 + w), four norms per layer
(pre/post attention, pre/post feed-forward), q/k RMSNorm, mixed sliding-window / global layers with two rope
bases, sqrt(hidden) embedding scale, GeGLU (tanh), query_pre_attn_scalar softmax scale, final logit soft-cap."""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from ...modules.attention import AttentionBase
from ...modules.mlp import GatedMLP
from ...modules.norm import RMSNorm
from ...modules.rope import RotaryEmbedding
from ...parallel.layers import ColumnParallelLinear, ParallelEmbedding
from ..llama.modeling_llama import LlamaInferenceConfig, NeuronLlamaForCausalLM, rope_scaling_of, rope_theta_of
from ..model_base import NeuronBaseModel
from ..state_dict_utils import fuse_qkv_and_gate_up


class Gemma3InferenceConfig(LlamaInferenceConfig):
    def add_derived_config(self):
        super().add_derived_config()
        if not hasattr(self, "hidden_act"):
            self.hidden_act = getattr(self, "hidden_activation", "gelu_pytorch_tanh")
        if getattr(self, "hidden_activation", None):
            self.hidden_act = self.hidden_activation


def _is_sliding(config, i: int) -> bool:
    lt = getattr(config, "layer_types", None)
    if lt:
        return lt[i] == "sliding_attention"
    pat = getattr(config, "sliding_window_pattern", 6)
    return bool((i + 1) % pat)


class Gemma3DecoderLayer(nn.Module):
    def __init__(self, config, i, rope_local, rope_global, device=None):
        super().__init__()
        nc = config.neuron_config
        dt = nc.torch_dtype
        sliding = _is_sliding(config, i)
        scalar = getattr(config, "query_pre_attn_scalar", config.head_dim)
        self.self_attn = AttentionBase(
            config, hidden_size=config.hidden_size, num_attention_heads=config.num_attention_heads,
            num_key_value_heads=config.num_key_value_heads, head_dim=config.head_dim,
            rotary_emb=rope_local if sliding else rope_global, qkv_bias=getattr(config, "attention_bias", False),
            o_bias=getattr(config, "attention_bias", False), sliding_window=config.sliding_window if sliding else None,
            qk_norm="rms_pre_rope", qk_norm_eps=config.rms_norm_eps, softmax_scale=1.0 / math.sqrt(scalar),
            logit_softcap=getattr(config, "attn_logit_softcapping", None), layer_idx=i, rms_norm_eps=config.rms_norm_eps,
            device=device)
        self.self_attn.q_layernorm.offset = 1.0
        self.self_attn.k_layernorm.offset = 1.0
        self.mlp = GatedMLP(config.hidden_size, config.intermediate_size, config.hidden_act, dt, device=device)
        mk = lambda: RMSNorm(config.hidden_size, config.rms_norm_eps, dt, offset=1.0, device=device)
        self.input_layernorm, self.post_attention_layernorm = mk(), mk()
        self.pre_feedforward_layernorm, self.post_feedforward_layernorm = mk(), mk()
        self.layer_idx = i

    def forward(self, h, meta, kv_mgr, lora=None):
        n = self.input_layernorm
        a = self.self_attn(h, meta, kv_mgr, norm_weight=n.weight, norm_eps=n.variance_epsilon, norm_offset=n.offset)
        h = h + self.post_attention_layernorm(a)
        n = self.pre_feedforward_layernorm
        m = self.mlp(h, norm_weight=n.weight, norm_eps=n.variance_epsilon, norm_offset=n.offset)
        return h + self.post_feedforward_layernorm(m)


class NeuronGemma3Model(NeuronBaseModel):
    def setup_attr_for_model(self, config):
        nc = config.neuron_config
        self.tp_degree, self.hidden_size = nc.tp_degree, config.hidden_size
        self.num_attention_heads, self.num_key_value_heads = config.num_attention_heads, config.num_key_value_heads
        self.max_batch_size, self.buckets = nc.max_batch_size, nc.buckets

    def init_model(self, config):
        nc = config.neuron_config
        dev, dt = self.device_, nc.torch_dtype
        self.embed_tokens = ParallelEmbedding(config.vocab_size, config.hidden_size, getattr(config, "pad_token_id", None),
                                              dtype=dt, device=dev, shard_across_embedding=not nc.vocab_parallel, pad=True,
                                              tensor_model_parallel_group=self.tp_group)
        self.embed_scale = float(torch.tensor(config.hidden_size ** 0.5, dtype=dt))
        maxpos = max(config.max_position_embeddings, nc.seq_len)
        rope_global = RotaryEmbedding(config.head_dim, maxpos, rope_theta_of(config, 1e6), rope_scaling_of(config), device=dev)
        rope_local = RotaryEmbedding(config.head_dim, maxpos, float(getattr(config, "rope_local_base_freq", 10000.0)), None,
                                     device=dev)
        self.layers = nn.ModuleList([Gemma3DecoderLayer(config, i, rope_local, rope_global, dev)
                                     for i in range(config.num_hidden_layers)])
        self.norm = RMSNorm(config.hidden_size, config.rms_norm_eps, dt, offset=1.0, device=dev)
        self.lm_head = ColumnParallelLinear(config.hidden_size, config.vocab_size, bias=False, gather_output=False, dtype=dt,
                                            device=dev, pad=True, tensor_model_parallel_group=self.tp_group)
        self.final_logit_softcap = getattr(config, "final_logit_softcapping", None)


class NeuronGemma3ForCausalLM(NeuronLlamaForCausalLM):
    _model_cls = NeuronGemma3Model

    @classmethod
    def get_config_cls(cls):
        return Gemma3InferenceConfig

    @staticmethod
    def convert_hf_to_neuron_state_dict(state_dict, config):
        sd = {k.replace("language_model.", ""): v for k, v in state_dict.items() if "vision_tower" not in k
              and "multi_modal_projector" not in k}
        sd = fuse_qkv_and_gate_up(sd, config.num_hidden_layers)
        return {k.replace("self_attn.q_norm.", "self_attn.q_layernorm.").replace("self_attn.k_norm.", "self_attn.k_layernorm."): v
                for k, v in sd.items()}
