"""reference models/pixtral/utils/input_processor.py."""
from ...qwen2_vl.utils.input_processor import prepare_generation_inputs_hf  # noqa: F401
