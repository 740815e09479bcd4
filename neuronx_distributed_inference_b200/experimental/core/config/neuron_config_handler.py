"""YAML configuration with per-model-tag overrides (reference experimental/core/config/neuron_config_handler.py:22-122).

    model: {name: llama3, path: /ckpt}
    build: {batch_size: 2, sequence_length: 1024, compiler_args: ["--a"]}
    attention: {try_using_kernel: true}
    dtype: ${torch_dtype:bfloat16}
    config_override:
      - model_tags: [prefill_1024, prefill_4096]
        build: {compiler_args: ["--b"]}
      - model_tags: [decode_4096]
        attention: {try_using_kernel: false}

``load_neuron_config`` returns ``{model, default, <tag>...}``: every tag starts from the default sections and merges its overrides
(dict sections merge key by key; ``build.compiler_args`` EXTENDS the default list, every other list is replaced).  The reference uses
OmegaConf; this is a small attribute-dict with the same access patterns (``cfg.build.batch_size``, ``"x" in cfg``, ``cfg.get``) and the
``${torch_dtype:NAME}`` resolver."""
from __future__ import annotations

import copy
import os
import re
from typing import Any

import torch


class Cfg(dict):
    """dict with attribute access, recursively applied to nested mappings / lists."""

    def __init__(self, d=None):
        super().__init__()
        for k, v in (d or {}).items():
            self[k] = _wrap(v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = _wrap(v)

    def __delattr__(self, k):
        del self[k]

    def __deepcopy__(self, memo):
        return Cfg({k: copy.deepcopy(v, memo) for k, v in self.items()})


_RESOLVER = re.compile(r"^\$\{torch_dtype:\s*([A-Za-z0-9_]+)\s*\}$")


def _wrap(v: Any):
    if isinstance(v, Cfg):
        return v
    if isinstance(v, dict):
        return Cfg(v)
    if isinstance(v, (list, tuple)):
        return [_wrap(x) for x in v]
    if isinstance(v, str):
        m = _RESOLVER.match(v.strip())
        if m:
            return getattr(torch, m.group(1))
    return v


def load_yaml_config(cfg_path: str) -> Cfg:
    """A file path or YAML text -> :class:`Cfg`."""
    import yaml
    if os.path.exists(cfg_path):
        with open(cfg_path) as f:
            return Cfg(yaml.safe_load(f) or {})
    return Cfg(yaml.safe_load(cfg_path) or {})


def _merge(base, upd):
    if isinstance(base, dict) and isinstance(upd, dict):
        out = Cfg(base)
        for k, v in upd.items():
            out[k] = _merge(out[k], v) if k in out else _wrap(copy.deepcopy(v))
        return out
    return _wrap(copy.deepcopy(upd))


def parse_config_with_model_tags_overrides(config: Cfg) -> Cfg:
    default = Cfg({k: copy.deepcopy(v) for k, v in config.items() if k not in ("config_override", "model")})
    parsed = Cfg({"model": config.get("model"), "default": default})
    default_args = list((default.get("build") or {}).get("compiler_args", []) or [])
    for override in config.get("config_override", []) or []:
        sections = {k: v for k, v in override.items() if k != "model_tags"}
        for tag in override["model_tags"]:
            cur = parsed[tag] if tag in parsed else copy.deepcopy(default)
            for name, upd in sections.items():
                cur[name] = _merge(cur[name], upd) if name in cur else _wrap(copy.deepcopy(upd))
                if name == "build":      # lists are replaced by a merge; compiler arguments accumulate instead
                    cur[name]["compiler_args"] = default_args + list((upd or {}).get("compiler_args", []) or [])
            parsed[tag] = cur
    return parsed


def load_neuron_config(config_file_path: str) -> Cfg:
    config = load_yaml_config(config_file_path)
    return parse_config_with_model_tags_overrides(config) if "config_override" in config else config


def get_config_for_model_tag(config: Cfg, model_tag: str) -> Cfg:
    return config[model_tag] if model_tag in config else config["default"]
