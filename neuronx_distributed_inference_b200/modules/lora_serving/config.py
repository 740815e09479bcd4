"""reference modules/lora_serving/config.py:9-224 — ``LoraServingConfig`` lives with the other config objects."""
from ...config import LoraServingConfig  # noqa: F401
