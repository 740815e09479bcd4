"""Tensor capture: intermediate activations as extra outputs of a forward (reference utils/tensor_capture_utils.py:22-426,
model side model_base.py:1043-1149, config TensorCaptureConfig).

The reference must thread captured tensors out of a traced graph as extra outputs; in an eager engine a capture is a forward
hook.  ``modules_to_capture`` are module paths relative to the device model (``layers.0.self_attn``, ``layers.1.mlp`` ...);
``capture_inputs`` also records the module inputs; ``auto_capture_moe_tensors`` records router logits / expert indices of
every MoE block (config.py:1135-1144).  Captures are CPU copies keyed ``<module>.outputs`` / ``<module>.inputs``."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn as nn


class TensorCapture:
    def __init__(self, model: nn.Module, modules_to_capture: List[str], capture_inputs: bool = False,
                 max_intermediate_tensors: Optional[int] = None, auto_capture_moe_tensors: bool = False):
        self.model = model
        self.names = list(modules_to_capture)
        self.capture_inputs = capture_inputs
        self.max = max_intermediate_tensors
        self.auto_moe = auto_capture_moe_tensors
        self.captured: Dict[str, torch.Tensor] = {}
        self._handles = []

    def __enter__(self):
        mods = dict(self.model.named_modules())
        for n in self.names:
            if n not in mods:
                raise KeyError(f"module '{n}' not found; available e.g. {list(mods)[:8]}")
            self._handles.append(mods[n].register_forward_hook(self._hook(n)))
        if self.auto_moe:
            from ..modules.moe import MoE
            for n, m in mods.items():
                if isinstance(m, MoE):
                    m.return_router_logits = m.return_expert_index = True
                    self._handles.append(m.register_forward_hook(self._moe_hook(n)))
        return self

    def __exit__(self, *exc):
        for h in self._handles:
            h.remove()
        self._handles = []

    def _store(self, key, t):
        if self.max is not None and len(self.captured) >= self.max:
            return
        if torch.is_tensor(t):
            self.captured[key] = t.detach().float().cpu()

    def _hook(self, name):
        def fn(mod, inputs, output):
            if self.capture_inputs:
                for i, x in enumerate(inputs):
                    self._store(f"{name}.inputs.{i}", x)
            out = output[0] if isinstance(output, (tuple, list)) else output
            self._store(f"{name}.outputs", out)
        return fn

    def _moe_hook(self, name):
        def fn(mod, inputs, output):
            self._store(f"{name}.router_logits", mod.last_router_logits)
            if mod.last_expert_index is not None:
                self._store(f"{name}.expert_index", mod.last_expert_index.float())
        return fn


def capture_model_tensors(app, modules_to_capture: List[str], *forward_args, capture_inputs: bool = False, **forward_kwargs):
    """Run ``app.forward`` once with capture hooks.  -> (output, {name: tensor})."""
    with TensorCapture(app.model, modules_to_capture, capture_inputs) as cap:
        out = app(*forward_args, **forward_kwargs)
    out.captured_tensors = cap.captured
    return out, cap.captured


def get_available_modules(app) -> List[str]:
    return [n for n, _ in app.model.named_modules() if n]
