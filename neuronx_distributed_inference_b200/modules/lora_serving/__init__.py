"""reference layout modules/lora_serving/* — implementation in modules/lora.py."""
from ..lora import AdapterCache, LoraLayer, LoraModel, LoraModelManager  # noqa: F401
from ...config import LoraServingConfig  # noqa: F401
