"""Families added after the reference hub was cut, each a few lines over the shared blocks:

* **Ministral / CWM (Code World Model)** — the Llama block with ``layer_types`` choosing sliding-window or full attention per layer
  (combinable with the window-sized rolling KV cache, ``rolling_sliding_window_cache``).
* **OLMo-1** — sequential pre-norm block with PARAMETER-FREE LayerNorms (unit weights are injected at load), SwiGLU, optional
  ``clip_qkv`` clamp on the fused projection output.
Each is checked against Hugging Face in fp32 (tests/test_contrib_cpu.py)."""
from __future__ import annotations

import torch

from ...models.llama.modeling_llama import NeuronLlamaAttention, NeuronLlamaForCausalLM, NeuronLlamaModel
from ...models.state_dict_utils import fuse_qkv_and_gate_up
from .classic_family import NeuronClassicModel, _ClassicCausalLM


class _LayerTypeAttention(NeuronLlamaAttention):
    def __init__(self, config, layer_idx, rotary_emb, device=None, **over):
        lt = getattr(config, "layer_types", None)
        sliding = lt[layer_idx] == "sliding_attention" if lt else getattr(config, "sliding_window", None) is not None
        b = bool(getattr(config, "attention_bias", False))
        super().__init__(config, layer_idx, rotary_emb, device=device, qkv_bias=b, o_bias=b,
                         sliding_window=getattr(config, "sliding_window", None) if sliding else None, **over)


class NeuronMinistralModel(NeuronLlamaModel):
    attention_cls = _LayerTypeAttention
    graph_safe = False


class NeuronMinistralForCausalLM(NeuronLlamaForCausalLM):
    _model_cls = NeuronMinistralModel


NeuronCwmForCausalLM = NeuronMinistralForCausalLM


class NeuronOlmoModel(NeuronClassicModel):
    SPEC = dict(NeuronClassicModel.SPEC, norm_bias=False)

    def layer_spec(self, config, i):
        return dict(parallel=False, norm_bias=False, mlp="gated", act=getattr(config, "hidden_act", "silu"), qkv_bias=False, o_bias=False,
                    mlp_bias=False, clip_qkv=getattr(config, "clip_qkv", None))


class NeuronOlmoForCausalLM(_ClassicCausalLM):
    _model_cls = NeuronOlmoModel

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        sd = fuse_qkv_and_gate_up(sd, config.num_hidden_layers)
        ones = torch.ones(config.hidden_size, dtype=next(iter(sd.values())).dtype)
        for i in range(config.num_hidden_layers):
            sd[f"layers.{i}.input_layernorm.weight"] = ones.clone()
            sd[f"layers.{i}.post_attention_layernorm.weight"] = ones.clone()
        sd["norm.weight"] = ones.clone()
        return sd


RECENT_MODEL_TYPES = {"ministral": NeuronMinistralForCausalLM, "cwm": NeuronCwmForCausalLM, "olmo": NeuronOlmoForCausalLM}


# ---- HunYuan (dense and MoE): per-head q/k RMSNorm applied AFTER the rotation ---------------------------------------------------------
class _HunYuanAttention(NeuronLlamaAttention):
    def __init__(self, config, layer_idx, rotary_emb, device=None, **over):
        b = bool(getattr(config, "attention_bias", False))
        qk = "rms_post_rope" if getattr(config, "use_qk_norm", True) else None
        super().__init__(config, layer_idx, rotary_emb, device=device, qkv_bias=b, o_bias=b, qk_norm=qk, qk_norm_eps=config.rms_norm_eps, **over)


def _hunyuan_names(sd):
    return {k.replace(".self_attn.query_layernorm.", ".self_attn.q_layernorm.").replace(".self_attn.key_layernorm.", ".self_attn.k_layernorm."): v
            for k, v in sd.items()}


class NeuronHunYuanDenseModel(NeuronLlamaModel):
    attention_cls = _HunYuanAttention
    graph_safe = False


class NeuronHunYuanDenseForCausalLM(NeuronLlamaForCausalLM):
    _model_cls = NeuronHunYuanDenseModel

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        return _hunyuan_names(NeuronLlamaForCausalLM.convert_hf_to_neuron_state_dict(sd, config))


RECENT_MODEL_TYPES["hunyuan_v1_dense"] = NeuronHunYuanDenseForCausalLM
