"""Tensor replacement: teacher-force intermediate activations from golden tensors to localise a numerical divergence
(reference utils/tensor_replacement/registry.py:148-577).  ``module_map`` maps module paths of the device model to golden
tensors (or to ``.pt`` files under ``ref_dir``); while active, the module's output is replaced by the golden value
(cast/moved to the live dtype/device), for the selected step indices only."""
from __future__ import annotations

import os
from typing import Dict, Iterable, Optional, Union

import torch
import torch.nn as nn


class TensorReplacementRegistry:
    def __init__(self, model: nn.Module, module_map: Dict[str, Union[torch.Tensor, str]], ref_dir: Optional[str] = None,
                 steps: Optional[Iterable[int]] = None):
        self.model = model
        self.map = dict(module_map)
        self.ref_dir = ref_dir
        self.steps = set(steps) if steps is not None else None
        self.step = 0
        self.replaced = []
        self._handles = []

    def _golden(self, name):
        g = self.map[name]
        if isinstance(g, str):
            path = g if os.path.isabs(g) or self.ref_dir is None else os.path.join(self.ref_dir, g)
            g = torch.load(path, map_location="cpu")
            self.map[name] = g
        return g

    def __enter__(self):
        mods = dict(self.model.named_modules())
        for n in self.map:
            if n not in mods:
                raise KeyError(f"module '{n}' not found")
            self._handles.append(mods[n].register_forward_hook(self._hook(n)))
        return self

    def __exit__(self, *exc):
        for h in self._handles:
            h.remove()
        self._handles = []

    def next_step(self):
        self.step += 1

    def _hook(self, name):
        def fn(mod, inputs, output):
            if self.steps is not None and self.step not in self.steps:
                return None
            live = output[0] if isinstance(output, (tuple, list)) else output
            g = self._golden(name).to(device=live.device, dtype=live.dtype)
            if g.shape != live.shape:
                raise ValueError(f"golden tensor for {name} has shape {tuple(g.shape)}, live output {tuple(live.shape)}")
            self.replaced.append((self.step, name))
            if isinstance(output, tuple):
                return (g,) + tuple(output[1:])
            if isinstance(output, list):
                return [g] + list(output[1:])
            return g
        return fn


def replace_tensors(app, module_map, *forward_args, ref_dir=None, **forward_kwargs):
    with TensorReplacementRegistry(app.model, module_map, ref_dir) as reg:
        out = app(*forward_args, **forward_kwargs)
    return out, reg.replaced
