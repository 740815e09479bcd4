"""Qwen3 dense (reference models/qwen3/modeling_qwen3.py:1-290): per-head RMSNorm on q and k before RoPE — fused
into the rope+kv-append kernel on the decode path (csrc/rope_kv.cu)."""
from __future__ import annotations

from ..llama.modeling_llama import (LlamaInferenceConfig, NeuronLlamaAttention, NeuronLlamaForCausalLM, NeuronLlamaModel)
from ..state_dict_utils import fuse_qkv_and_gate_up


class Qwen3InferenceConfig(LlamaInferenceConfig):
    pass


class NeuronQwen3Attention(NeuronLlamaAttention):
    def __init__(self, config, layer_idx, rotary_emb, device=None, **over):
        super().__init__(config, layer_idx, rotary_emb, device=device, qk_norm="rms_pre_rope",
                         qk_norm_eps=config.rms_norm_eps, qkv_bias=getattr(config, "attention_bias", False), **over)


class NeuronQwen3Model(NeuronLlamaModel):
    attention_cls = NeuronQwen3Attention


class NeuronQwen3ForCausalLM(NeuronLlamaForCausalLM):
    _model_cls = NeuronQwen3Model

    @classmethod
    def get_config_cls(cls):
        return Qwen3InferenceConfig

    @staticmethod
    def convert_hf_to_neuron_state_dict(state_dict, config):
        sd = fuse_qkv_and_gate_up(state_dict, config.num_hidden_layers)
        out = {}
        for k, v in sd.items():
            k = k.replace("self_attn.q_norm.", "self_attn.q_layernorm.").replace("self_attn.k_norm.", "self_attn.k_layernorm.")
            out[k] = v
        return out
