"""reference layout modules/lora_serving/{config,lora_checkpoint,lora_layer,lora_module,lora_model}.py — the layers and the
dynamic-adapter machinery are implemented in modules/lora.py."""
from ..lora import AdapterCache, LoraLayer, LoraModel, LoraModelManager  # noqa: F401
from .config import LoraServingConfig  # noqa: F401
from .lora_checkpoint import LoraCheckpoint  # noqa: F401
from .lora_model import LoraWeightManager  # noqa: F401
