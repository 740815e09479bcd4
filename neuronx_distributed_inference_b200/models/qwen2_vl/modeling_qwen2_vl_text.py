"""Import path of the reference (models/qwen2_vl/modeling_qwen2_vl_text.py): the text decoder with multimodal RoPE.  One application
serves text-only and image prompts here, so the text names alias it."""
from .modeling_qwen2_vl import (NeuronQwen2VLForCausalLM, NeuronQwen2VLTextModel, Qwen2VLInferenceConfig, get_rope_index,  # noqa: F401
                                mrope_section_of)

NeuronQwen2VLTextForCausalLM = NeuronQwen2VLForCausalLM
