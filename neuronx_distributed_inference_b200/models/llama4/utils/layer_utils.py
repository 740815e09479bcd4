"""Which Llama-4 layers neighbour a NoPE (global-attention) layer (reference models/llama4/utils/layer_utils.py).  ``no_rope_layers[i] == 0``
marks layer ``i`` as NoPE (Hugging Face convention: 1 = rotary)."""


def is_before_nope_layer(config, layer_idx: int) -> bool:
    nxt = layer_idx + 1
    return nxt < len(config.no_rope_layers) and config.no_rope_layers[nxt] == 0


def is_after_nope_layer(config, layer_idx: int) -> bool:
    return layer_idx == 0 or config.no_rope_layers[layer_idx - 1] == 0
