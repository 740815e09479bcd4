"""Qwen2 / Qwen2.5 (reference models/qwen2/modeling_qwen2.py:1-283): Llama block with q/k/v biases; optional
sliding window on the upper layers (``use_sliding_window`` / ``max_window_layers``)."""
from __future__ import annotations

from ..llama.modeling_llama import (LlamaInferenceConfig, NeuronLlamaAttention, NeuronLlamaForCausalLM, NeuronLlamaModel)


class Qwen2InferenceConfig(LlamaInferenceConfig):
    pass


class NeuronQwen2Attention(NeuronLlamaAttention):
    def __init__(self, config, layer_idx, rotary_emb, device=None, **over):
        sw = None
        if getattr(config, "use_sliding_window", False) and layer_idx >= getattr(config, "max_window_layers", 0):
            sw = getattr(config, "sliding_window", None)
        super().__init__(config, layer_idx, rotary_emb, device=device, qkv_bias=True, o_bias=False,
                         sliding_window=sw, **over)


class NeuronQwen2Model(NeuronLlamaModel):
    attention_cls = NeuronQwen2Attention


class NeuronQwen2ForCausalLM(NeuronLlamaForCausalLM):
    _model_cls = NeuronQwen2Model

    @classmethod
    def get_config_cls(cls):
        return Qwen2InferenceConfig
