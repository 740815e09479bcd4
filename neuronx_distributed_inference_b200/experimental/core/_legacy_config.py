"""YAML configuration with per-tag overrides (reference experimental/core/config/neuron_config_handler.py:1-122).

    common: {tp_degree: 8, batch_size: 4, seq_len: 4096, torch_dtype: bfloat16}
    context_encoding_model: {buckets: [512, 1024, 4096]}
    token_generation_model: {buckets: [1024, 4096], cuda_graphs: true}
"""
from __future__ import annotations

import copy
from typing import Dict

from ...config import NeuronConfig


def load_yaml_config(path_or_text: str) -> Dict:
    import os
    import yaml
    if os.path.exists(path_or_text):
        with open(path_or_text) as f:
            return yaml.safe_load(f) or {}
    return yaml.safe_load(path_or_text) or {}


class NeuronConfigHandler:
    def __init__(self, cfg: Dict, neuron_config_cls=NeuronConfig):
        self.raw = cfg
        self.cls = neuron_config_cls
        self.common = dict(cfg.get("common", {}))

    def tags(self):
        return [k for k in self.raw if k != "common"]

    def for_tag(self, tag: str) -> NeuronConfig:
        kw = copy.deepcopy(self.common)
        kw.update(self.raw.get(tag, {}))
        return self.cls(**kw)

    def common_config(self) -> NeuronConfig:
        return self.cls(**self.common)
