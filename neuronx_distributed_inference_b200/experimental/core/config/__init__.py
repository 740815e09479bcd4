"""Configuration objects of the experimental flow (reference experimental/core/config/**)."""
from .._legacy_config import NeuronConfigHandler  # noqa: F401
from .attention import AttentionConfig  # noqa: F401
from .build import BuildConfig  # noqa: F401
from .neuron_config_handler import (Cfg, get_config_for_model_tag, load_neuron_config, load_yaml_config,  # noqa: F401
                                    parse_config_with_model_tags_overrides)
