"""Import path of the reference (models/qwen3_vl/modeling_qwen3_vl_text.py)."""
from .modeling_qwen3_vl import NeuronQwen3VLForCausalLM, NeuronQwen3VLTextModel, Qwen3VLInferenceConfig  # noqa: F401

NeuronQwen3VLTextForCausalLM = NeuronQwen3VLForCausalLM
