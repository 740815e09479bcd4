"""GPT-OSS (reference models/gpt_oss/modeling_gpt_oss.py:1-1217): learned attention sinks, alternating
sliding-window / full layers, biased q/k/v/o, YaRN RoPE, MoE with a biased router that soft-maxes over the
selected top-k, clamped SwiGLU experts with biases (``(up+1) * gate*sigmoid(1.702*gate)``).
MXFP4 expert checkpoints are de-quantised at load (``mx_layout_transform.py`` of the reference repacks them for the
Neuron engines; on B200 the block-scaled path would feed ``tcgen05.mma.kind::mxf4`` — see DESIGN.md)."""
from __future__ import annotations

import torch

from ...config import MoENeuronConfig
from ...modules.moe import initialize_moe_module
from ...modules.norm import RMSNorm
from ..llama.modeling_llama import LlamaInferenceConfig, NeuronLlamaAttention, NeuronLlamaForCausalLM, NeuronLlamaModel
from ..model_base import DecoderLayer
from ..state_dict_utils import fuse_qkv_and_gate_up

ALPHA, LIMIT = 1.702, 7.0


def gpt_oss_glu(h: torch.Tensor) -> torch.Tensor:
    """h = [gate | up] halves (our layout) -> clamped SwiGLU of the reference (:GptOssExperts._apply_gate)."""
    gate, up = h.float().chunk(2, -1)
    gate = gate.clamp(max=LIMIT)
    up = up.clamp(-LIMIT, LIMIT)
    return ((up + 1) * gate * torch.sigmoid(gate * ALPHA)).to(h.dtype)


gpt_oss_glu.kernel_act = "gpt_oss_glu"      # epilogue of the grouped tcgen05 GEMM (csrc/gemm_tcgen05.cu, act 4)


class GptOssInferenceConfig(LlamaInferenceConfig):
    def get_required_attributes(self):
        return super().get_required_attributes() + ["num_local_experts", "num_experts_per_tok"]

    @classmethod
    def get_neuron_config_cls(cls):
        return MoENeuronConfig


def _sliding(config, i):
    lt = getattr(config, "layer_types", None)
    return (lt[i] == "sliding_attention") if lt else (i % 2 == 0)


class NeuronGptOssModel(NeuronLlamaModel):
    graph_safe = False            # the PyTorch expert dispatch synchronises ...
    moe_decode_graph_safe = True  # ... the grouped-GEMM path (biases + clamped SwiGLU in the epilogue) does not

    def make_layer(self, config, i, rotary, device):
        nc = config.neuron_config
        attn = NeuronLlamaAttention(config, i, rotary, device=device, qkv_bias=True, o_bias=True, learned_sinks=True,
                                    sliding_window=config.sliding_window if _sliding(config, i) else None)
        moe = initialize_moe_module(config, device=device, router_bias=True, expert_bias=True, act_fn=gpt_oss_glu,
                                    apply_act_fn_over_topk=True)
        return DecoderLayer(attn, moe, RMSNorm(config.hidden_size, config.rms_norm_eps, nc.torch_dtype, device=device),
                            RMSNorm(config.hidden_size, config.rms_norm_eps, nc.torch_dtype, device=device), i, mlp_is_moe=True)


class NeuronGptOssForCausalLM(NeuronLlamaForCausalLM):
    _model_cls = NeuronGptOssModel

    @classmethod
    def get_config_cls(cls):
        return GptOssInferenceConfig

    @staticmethod
    def convert_hf_to_neuron_state_dict(state_dict, config):
        if any(k.endswith("_blocks") for k in state_dict):      # released MXFP4 checkpoints
            from .mx_layout_transform import dequantize_mxfp4_state_dict
            state_dict = dequantize_mxfp4_state_dict(state_dict, config.neuron_config.torch_dtype)
        sd = fuse_qkv_and_gate_up(state_dict, config.num_hidden_layers, fuse_mlp=False)
        out = {}
        for k, v in sd.items():
            if k.endswith(".mlp.router.weight"):
                out[k.replace(".mlp.router.weight", ".mlp.router.linear_router.weight")] = v.float()
            elif k.endswith(".mlp.router.bias"):
                out[k.replace(".mlp.router.bias", ".mlp.router.linear_router.bias")] = v.float()
            elif k.endswith(".mlp.experts.gate_up_proj"):      # [E, H, 2I] interleaved (gate, up, gate, up ...)
                w = v.transpose(1, 2)                           # [E, 2I, H]
                out[k.replace(".mlp.experts.gate_up_proj", ".mlp.expert_mlps.gate_up_proj")] = \
                    torch.cat([w[:, 0::2], w[:, 1::2]], 1).contiguous()
            elif k.endswith(".mlp.experts.gate_up_proj_bias"):  # [E, 2I] interleaved
                out[k.replace(".mlp.experts.gate_up_proj_bias", ".mlp.expert_mlps.gate_up_bias")] = \
                    torch.cat([v[:, 0::2], v[:, 1::2]], 1).contiguous()
            elif k.endswith(".mlp.experts.down_proj"):          # [E, I, H] -> [E, H, I]
                out[k.replace(".mlp.experts.down_proj", ".mlp.expert_mlps.down_proj")] = v.transpose(1, 2).contiguous()
            elif k.endswith(".mlp.experts.down_proj_bias"):
                out[k.replace(".mlp.experts.down_proj_bias", ".mlp.expert_mlps.down_bias")] = v
            else:
                out[k] = v
        return out
