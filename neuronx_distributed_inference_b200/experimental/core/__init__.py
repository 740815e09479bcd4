"""Experimental core (reference experimental/core/**): YAML configuration with per-model-tag overrides, build flows (static-shape
"built" models with one entry per tag), padding helpers, the bucketing processor, greedy generation, sharded-safetensors loading and
the logit-validation algorithm."""
from .config import NeuronConfigHandler, load_yaml_config  # noqa: F401
from .generate import GenerateResult, generate  # noqa: F401
from .processor import BucketingProcessor  # noqa: F401
