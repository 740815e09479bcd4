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
from ...modules.norm import RMSNorm
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
        if self.neuron_config.flash_decoding_enabled:
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
        self.norm = RMSNorm(config.hidden_size, config.rms_norm_eps, dt, device=dev)
        self.lm_head = ColumnParallelLinear(config.hidden_size, config.vocab_size, bias=False,
                                            gather_output=False, dtype=dt, device=dev, pad=True,
                                            tensor_model_parallel_group=self.tp_group)


class NeuronLlamaForCausalLM(NeuronBaseForCausalLM):
    _model_cls = NeuronLlamaModel

    @classmethod
    def get_config_cls(cls):
        return LlamaInferenceConfig

    @staticmethod
    def convert_hf_to_neuron_state_dict(state_dict: dict, config) -> dict:
        return fuse_qkv_and_gate_up(state_dict, config.num_hidden_layers)

    @staticmethod
    def update_state_dict_for_tied_weights(state_dict):
        state_dict["lm_head.weight"] = state_dict["embed_tokens.weight"].clone()
