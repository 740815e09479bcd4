"""Mixtral 8x7B / 8x22B (reference models/mixtral/modeling_mixtral.py:1-330): Llama attention + MoE (softmax router,
top-2, renormalised)."""
from __future__ import annotations

import torch

from ...config import MoENeuronConfig
from ...modules.moe import initialize_moe_module
from ...modules.norm import RMSNorm
from ..llama.modeling_llama import LlamaInferenceConfig, NeuronLlamaAttention, NeuronLlamaForCausalLM, NeuronLlamaModel
from ..model_base import DecoderLayer
from ..state_dict_utils import convert_moe_experts, fuse_qkv_and_gate_up


class MixtralInferenceConfig(LlamaInferenceConfig):
    def get_required_attributes(self):
        return super().get_required_attributes() + ["num_local_experts", "num_experts_per_tok"]

    @classmethod
    def get_neuron_config_cls(cls):
        return MoENeuronConfig


class NeuronMixtralModel(NeuronLlamaModel):
    graph_safe = False            # the torch expert dispatch synchronises ...
    moe_decode_graph_safe = True  # ... but decode (T <= 8) runs the moe_decode kernels: CUDA graphs allowed when they apply

    def make_layer(self, config, i, rotary, device):
        nc = config.neuron_config
        attn = NeuronLlamaAttention(config, i, rotary, device=device, sliding_window=getattr(config, "sliding_window", None))
        moe = initialize_moe_module(config, device=device)
        return DecoderLayer(attn, moe, RMSNorm(config.hidden_size, config.rms_norm_eps, nc.torch_dtype, device=device),
                            RMSNorm(config.hidden_size, config.rms_norm_eps, nc.torch_dtype, device=device), i, mlp_is_moe=True)


class NeuronMixtralForCausalLM(NeuronLlamaForCausalLM):
    _model_cls = NeuronMixtralModel

    @classmethod
    def get_config_cls(cls):
        return MixtralInferenceConfig

    @staticmethod
    def convert_hf_to_neuron_state_dict(state_dict, config):
        sd = fuse_qkv_and_gate_up(state_dict, config.num_hidden_layers, fuse_mlp=False)
        return convert_moe_experts(sd, config.num_hidden_layers, config.num_local_experts,
                                   moe_prefixes=("mlp", "block_sparse_moe"), gate_names=("gate",), w_names=("w1", "w3", "w2"))
