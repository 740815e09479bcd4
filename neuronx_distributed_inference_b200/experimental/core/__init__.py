"""Experimental core (reference experimental/core/**): YAML configuration with per-sub-model overrides, bucketing processor,
greedy generation and the logit-validation algorithm."""
from .config import NeuronConfigHandler, load_yaml_config  # noqa: F401
from .generate import generate  # noqa: F401
from .processor import BucketingProcessor  # noqa: F401
