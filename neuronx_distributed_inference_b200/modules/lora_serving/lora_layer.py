"""reference modules/lora_serving/lora_layer.py:10-358 (per-layer stacked A/B with per-row adapter selection)."""
from ..lora import LoraLayer  # noqa: F401
