"""reference import path ``neuronx_distributed_inference.models.config`` -> package-level ``config`` module."""
from ..config import *  # noqa: F401,F403
from ..config import (ChunkedPrefillConfig, FusedSpecNeuronConfig, InferenceConfig, KVQuantizationConfig, LoraServingConfig,  # noqa: F401
                      MoENeuronConfig, NeuronConfig, OnDeviceSamplingConfig, to_torch_dtype)
