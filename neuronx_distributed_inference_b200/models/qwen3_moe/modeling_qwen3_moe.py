"""Qwen3-MoE (reference models/qwen3_moe/modeling_qwen3_moe.py:1-543): Qwen3 attention (q/k RMSNorm) + MoE with an
fp32 softmax router, ``norm_topk_prob``; dense layers where ``mlp_only_layers`` / ``decoder_sparse_step`` say so;
HF block-fp8 checkpoints are dequantised on load (:103-118)."""
from __future__ import annotations

import torch

from ...config import MoENeuronConfig
from ...modules.moe import initialize_moe_module
from ...modules.norm import RMSNorm
from ..llama.modeling_llama import LlamaInferenceConfig, NeuronLlamaForCausalLM, NeuronLlamaMLP, NeuronLlamaModel
from ..model_base import DecoderLayer
from ..qwen3.modeling_qwen3 import NeuronQwen3Attention
from ..state_dict_utils import convert_moe_experts, dequantize_block_fp8, fuse_qkv_and_gate_up


class Qwen3MoeInferenceConfig(LlamaInferenceConfig):
    def get_required_attributes(self):
        return super().get_required_attributes() + ["num_experts", "num_experts_per_tok", "moe_intermediate_size"]

    @classmethod
    def get_neuron_config_cls(cls):
        return MoENeuronConfig


def _is_moe_layer(config, i):
    if i in (getattr(config, "mlp_only_layers", None) or []):
        return False
    step = getattr(config, "decoder_sparse_step", 1) or 1
    return config.num_experts > 0 and (i + 1) % step == 0


class NeuronQwen3MoeModel(NeuronLlamaModel):
    graph_safe = False            # the torch expert dispatch synchronises ...
    moe_decode_graph_safe = True  # ... but decode (T <= 8) runs the moe_decode kernels: CUDA graphs allowed when they apply

    def make_layer(self, config, i, rotary, device):
        nc = config.neuron_config
        attn = NeuronQwen3Attention(config, i, rotary, device=device)
        if _is_moe_layer(config, i):
            mlp = initialize_moe_module(config, device=device, intermediate_size=config.moe_intermediate_size,
                                        normalize=bool(getattr(config, "norm_topk_prob", True)))
        else:
            mlp = NeuronLlamaMLP(config, device=device)
        return DecoderLayer(attn, mlp, RMSNorm(config.hidden_size, config.rms_norm_eps, nc.torch_dtype, device=device),
                            RMSNorm(config.hidden_size, config.rms_norm_eps, nc.torch_dtype, device=device), i)


class NeuronQwen3MoeForCausalLM(NeuronLlamaForCausalLM):
    _model_cls = NeuronQwen3MoeModel

    @classmethod
    def get_config_cls(cls):
        return Qwen3MoeInferenceConfig

    @staticmethod
    def convert_hf_to_neuron_state_dict(state_dict, config):
        sd = dequantize_block_fp8(state_dict, config.neuron_config.torch_dtype)
        moe_layers = [i for i in range(config.num_hidden_layers) if _is_moe_layer(config, i)]
        sd = fuse_qkv_and_gate_up(sd, config.num_hidden_layers, fuse_mlp=True)
        sd = convert_moe_experts(sd, config.num_hidden_layers, config.num_experts, moe_prefixes=("mlp",), gate_names=("gate",),
                                 w_names=("gate_proj", "up_proj", "down_proj"), layers=moe_layers)
        return {k.replace("self_attn.q_norm.", "self_attn.q_layernorm.").replace("self_attn.k_norm.", "self_attn.k_layernorm."): v
                for k, v in sd.items()}
