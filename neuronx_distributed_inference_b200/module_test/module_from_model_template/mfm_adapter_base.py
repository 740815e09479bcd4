"""Module-from-model adapters (reference module_test/module_from_model_template/mfm_adapter_base.py:34-407): cut named sub-modules out of a
COMPLETE model — a Hugging Face ``PreTrainedModel`` or an engine application — by their paths, wrap them in a small module whose
``forward`` is supplied by the test, and load only their weights.

``build_prefixes_map(prefixes, names, layer_id, default)`` -> ``{name: prefix}`` where a prefix containing ``"layer"`` gets the layer id
appended (``"model.layers"`` -> ``"model.layers.3"``), others are used as they are (``"model"`` for e.g. a rotary embedding)."""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional

import torch
import torch.nn as nn

from ..base_template.adapter_base import B200DeviceAdapterBase, HFAdapterBase, NxDISingleRankCPUAdapterBase


def build_prefixes_map(prefixes: Optional[List[str]], needed_module_names: List[str], layer_id: int, default_prefix: str) -> Dict[str, str]:
    if prefixes is None:
        prefixes = [default_prefix] * len(needed_module_names)
    if len(prefixes) != len(needed_module_names):
        raise ValueError(f"one prefix per module: {len(prefixes)} prefixes for {len(needed_module_names)} modules")
    return {n: (f"{p}.{layer_id}" if "layer" in p else p) for n, p in zip(needed_module_names, prefixes)}


def _join(prefix: str, name: str) -> str:
    return f"{prefix}.{name}" if prefix else name


def extract_submodules_by_prefixes(prefixes_map: Dict[str, str], all_modules: Dict[str, nn.Module]) -> Dict[str, nn.Module]:
    out = {}
    for name, pref in prefixes_map.items():
        path = _join(pref, name)
        if path not in all_modules:
            raise KeyError(f"no module {path!r} in the complete model")
        out[name] = all_modules[path]
    return out


def extract_subweights_by_prefixes(prefixes_map: Dict[str, str], full_sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Weights under ``<prefix>.<name>.`` with the prefix stripped (keys then start with the module name, as in the partial module)."""
    out = {}
    for name, pref in prefixes_map.items():
        root = _join(pref, name)
        strip = len(pref) + 1 if pref else 0
        for k, v in full_sd.items():
            if k == root or k.startswith(root + "."):
                out[k[strip:]] = v
    return out


def _partial_cls(sub: Dict[str, nn.Module], forward_fn: Callable) -> type:
    class Partial(nn.Module):
        def __init__(self):
            super().__init__()
            for k, m in sub.items():
                setattr(self, k, m)
    Partial.forward = forward_fn
    return Partial


class MFMHFAdapter(HFAdapterBase):
    """``forward_fn(self, **inputs)`` runs over attributes named like ``module_names`` taken from ``complete_model_cls.from_pretrained``."""

    def __init__(self, forward_fn: Callable[..., Any], complete_model_cls, module_names: List[str], layer_id: int = 0,
                 prefixes: Optional[List[str]] = None):
        super().__init__()
        self.forward_fn, self.complete_model_cls = forward_fn, complete_model_cls
        self.prefixes_map = build_prefixes_map(prefixes, module_names, layer_id, "model.layers")
        self._model = None

    def _complete(self):
        if self._model is None:
            assert self.hf_ckpt_path, "the orchestrator sets hf_ckpt_path"
            self._model = self.complete_model_cls.from_pretrained(self.hf_ckpt_path).eval()
        return self._model

    def define_module_cls(self):
        sub = extract_submodules_by_prefixes(self.prefixes_map, dict(self._complete().named_modules()))
        self.module_cls = _partial_cls(sub, self.forward_fn)

    def get_state_dict(self):
        return extract_subweights_by_prefixes(self.prefixes_map, self._complete().state_dict())

    def free_resources(self):
        self._model = None
        super().free_resources()


class _MFMEngineMixin:
    """Shared by the CPU and the device flavour: the complete model is an engine application built from the checkpoint with
    ``app_factory(ckpt_path, device_str)``; the sub-modules keep the weights the application loaded (already converted / sharded)."""

    def _init_mfm(self, forward_fn, app_factory, needed_module_names, layer_id, prefixes):
        self.forward_fn, self.app_factory = forward_fn, app_factory
        self.prefixes_map = build_prefixes_map(prefixes, needed_module_names, layer_id, "layers")
        self.app = None

    def _application(self):
        if self.app is None:
            assert self.hf_ckpt_path, "the orchestrator sets hf_ckpt_path"
            self.app = self.app_factory(self.hf_ckpt_path, self.device.type)
        return self.app

    def define_module_cls(self):
        model = self._application().model
        sub = extract_submodules_by_prefixes(self.prefixes_map, dict(model.named_modules()))
        self.module_cls = _partial_cls(sub, self.forward_fn)

    def instantiate_module(self, example_inputs=None):
        # the application already holds the (converted) weights on the right device: the partial module shares them
        self.module = self.module_cls().eval()
        self.torch_dtype = next(self.module.parameters()).dtype

    def load_kv_cache(self, hf_kv_cache, layer_idx: int = 0):
        """HF layout ``[B, H_kv, S, D]`` -> cache lines ``0..B-1`` of the application's cache manager (single rank)."""
        k, v = hf_kv_cache
        mgr = self._application().model.kv_mgr
        mgr.reset()
        kc, vc = mgr.get_kv_by_layer_id(layer_idx)
        B, _, S, D = k.shape
        kc[:B, :, :S, :D] = k.to(kc.device, kc.dtype)
        vc[:B, :, :S, :D] = v.to(vc.device, vc.dtype)

    def free_resources(self):
        self.app = None
        super().free_resources()


class MFMNxDICPUSingleRankAdapter(_MFMEngineMixin, NxDISingleRankCPUAdapterBase):
    def __init__(self, forward_fn: Callable[..., Any], app_factory: Callable[[str, str], Any], needed_module_names: List[str],
                 layer_id: int = 0, prefixes: Optional[List[str]] = None):
        NxDISingleRankCPUAdapterBase.__init__(self)
        self._init_mfm(forward_fn, app_factory, needed_module_names, layer_id, prefixes)


class MFMB200DeviceAdapter(_MFMEngineMixin, B200DeviceAdapterBase):
    def __init__(self, forward_fn: Callable[..., Any], app_factory: Callable[[str, str], Any], needed_module_names: List[str],
                 layer_id: int = 0, prefixes: Optional[List[str]] = None, tp_degree: int = 1, world_size: int = 1):
        B200DeviceAdapterBase.__init__(self, tp_degree, world_size)
        self._init_mfm(forward_fn, app_factory, needed_module_names, layer_id, prefixes)

    def instantiate_module(self, example_inputs=None):
        if self.device.type != "cuda":
            raise RuntimeError("the device adapter needs a GPU")
        _MFMEngineMixin.instantiate_module(self, example_inputs)


MFMNxDINeuronAdapter = MFMB200DeviceAdapter
