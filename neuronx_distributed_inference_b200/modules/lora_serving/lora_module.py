"""reference modules/lora_serving/lora_module.py:16-208 wraps Linear / Column / RowParallel modules one by one.  The engine's
projections are fused (``qkv_proj``, ``gate_up_proj``), so the wrapping unit is the fused projection: ``TARGETS`` lists them and
``HF_TO_FUSED`` says which row block of the fused delta each PEFT target (``q_proj`` ...) writes."""
from ..lora import TARGETS, _HF_TO_FUSED as HF_TO_FUSED, _LayerHook as LoraModuleHook  # noqa: F401
