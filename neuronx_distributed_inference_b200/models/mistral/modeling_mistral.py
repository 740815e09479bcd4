"""Mistral (reference models/mistral/modeling_mistral.py:1-229): Llama block + optional sliding-window attention."""
from __future__ import annotations

from ..llama.modeling_llama import (LlamaInferenceConfig, NeuronLlamaAttention, NeuronLlamaForCausalLM, NeuronLlamaModel)


class MistralInferenceConfig(LlamaInferenceConfig):
    pass


class NeuronMistralAttention(NeuronLlamaAttention):
    def __init__(self, config, layer_idx, rotary_emb, device=None, **over):
        sw = getattr(config, "sliding_window", None)
        super().__init__(config, layer_idx, rotary_emb, device=device, sliding_window=sw, **over)


class NeuronMistralModel(NeuronLlamaModel):
    attention_cls = NeuronMistralAttention


class NeuronMistralForCausalLM(NeuronLlamaForCausalLM):
    _model_cls = NeuronMistralModel

    @classmethod
    def get_config_cls(cls):
        return MistralInferenceConfig
