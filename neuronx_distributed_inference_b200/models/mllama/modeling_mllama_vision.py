"""Import path of the reference (models/mllama/modeling_mllama_vision.py)."""
from .modeling_mllama import MllamaVisionLayer, NeuronMllamaVisionModel  # noqa: F401

ImageTransformerBlock = MllamaVisionLayer
