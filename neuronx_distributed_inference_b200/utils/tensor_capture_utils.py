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


# ---- capture sessions on disk (reference :22-113, :212-229, :231-426) --------------------------------------------------------------
import json  # noqa: E402
import os  # noqa: E402
import time  # noqa: E402


class TensorCaptureMetadata:
    """One ``capture_metadata.json`` per capture directory: for every saved ``.pt`` file the generation step, the phase
    (``cte`` / ``tkg``), the module it came from, shape and dtype.  Re-opening a directory appends to the existing index."""

    FILE = "capture_metadata.json"

    def __init__(self, save_dir: str):
        self.save_dir = save_dir
        os.makedirs(save_dir, exist_ok=True)
        self.path = os.path.join(save_dir, self.FILE)
        self.metadata = {"capture_session": {"created_at": time.time(), "version": "1.0"}, "tensors": {}}
        if os.path.isfile(self.path):
            try:
                with open(self.path) as f:
                    old = json.load(f)
                self.metadata["tensors"].update(old.get("tensors", {}))
                self.metadata["capture_session"] = old.get("capture_session", self.metadata["capture_session"])
            except (OSError, ValueError):
                pass

    def add_tensor(self, filename: str, step: int, phase: str, tensor_type: str, tensor: torch.Tensor, module_name: Optional[str] = None):
        self.metadata["tensors"][filename] = dict(step=step, phase=phase, tensor_type=tensor_type, tensor_shape=list(tensor.shape),
                                                  tensor_dtype=str(tensor.dtype), module_name=module_name, timestamp=time.time())

    def save(self):
        with open(self.path, "w") as f:
            json.dump(self.metadata, f, indent=2)


def get_tensor_capture_hook(modules_to_capture: List[str], capture_indices: Optional[List[int]] = None,
                            tensor_capture_save_dir: str = "captured_tensors", capture_inputs: bool = False):
    """Returns ``hook(app, step, *forward_args, **forward_kwargs) -> output`` to call INSTEAD of ``app(...)`` inside a generation
    loop: at the steps listed in ``capture_indices`` (all steps when None) the forward runs under :class:`TensorCapture` and every
    captured tensor is written to ``<dir>/step<k>_<phase>_<module>.pt`` with an entry in the metadata index."""
    meta = TensorCaptureMetadata(tensor_capture_save_dir)
    want = None if capture_indices is None else set(capture_indices)

    def hook(app, step: int, *args, **kwargs):
        if want is not None and step not in want:
            return app(*args, **kwargs)
        out, cap = capture_model_tensors(app, modules_to_capture, *args, capture_inputs=capture_inputs, **kwargs)
        ids = args[0] if args else kwargs.get("input_ids")
        phase = "cte" if ids is not None and ids.shape[-1] > 1 else "tkg"
        for name, t in cap.items():
            fn = f"step{step}_{phase}_{name.replace('.', '_')}.pt"
            torch.save(t, os.path.join(tensor_capture_save_dir, fn))
            kind = "input" if ".inputs." in name else "output"
            meta.add_tensor(fn, step, phase, kind, t, module_name=name.rsplit(".outputs", 1)[0].split(".inputs.")[0])
        meta.save()
        return out
    return hook


def list_capturable_modules_in_application(app) -> Dict[str, List[str]]:
    """Module paths usable in ``modules_to_capture`` grouped by kind (attention / mlp / moe / norm / layer / other), per sub-model
    family the paths are identical: the CTE and TKG runners share one device model here."""
    groups: Dict[str, List[str]] = {"layer": [], "attention": [], "mlp": [], "moe": [], "norm": [], "other": []}
    from ..modules.moe import MoE
    for n, m in app.model.named_modules():
        if not n:
            continue
        leaf = n.rsplit(".", 1)[-1]
        if isinstance(m, MoE):
            groups["moe"].append(n)
        elif leaf in ("self_attn", "cross_attn", "attention"):
            groups["attention"].append(n)
        elif leaf in ("mlp", "feed_forward"):
            groups["mlp"].append(n)
        elif "norm" in leaf:
            groups["norm"].append(n)
        elif n.startswith("layers.") and n.count(".") == 1:
            groups["layer"].append(n)
        else:
            groups["other"].append(n)
    return groups


def analyze_captured_tensors(tensor_dir: str, reference_dir: Optional[str] = None, rtol: float = 1e-2, atol: float = 1e-3) -> Dict[str, dict]:
    """Statistics of every captured tensor (shape, mean, std, abs-max, NaN / Inf counts); with ``reference_dir`` (a capture of the
    same run from a trusted configuration, e.g. fp32 on CPU) also max abs / relative error and an allclose verdict, so the FIRST
    module that diverges can be read off the step-ordered report."""
    meta = TensorCaptureMetadata(tensor_dir).metadata["tensors"]
    files = sorted((f for f in os.listdir(tensor_dir) if f.endswith(".pt")), key=lambda f: (meta.get(f, {}).get("step", 0), f))
    report: Dict[str, dict] = {}
    for f in files:
        t = torch.load(os.path.join(tensor_dir, f)).float()
        r = dict(shape=list(t.shape), mean=float(t.mean()), std=float(t.std()) if t.numel() > 1 else 0.0, absmax=float(t.abs().max()),
                 nan=int(torch.isnan(t).sum()), inf=int(torch.isinf(t).sum()), **{k: meta.get(f, {}).get(k) for k in ("step", "phase", "module_name")})
        if reference_dir is not None and os.path.isfile(os.path.join(reference_dir, f)):
            g = torch.load(os.path.join(reference_dir, f)).float()
            if g.shape == t.shape:
                d = (t - g).abs()
                r.update(max_abs_err=float(d.max()), rel_err=float(d.norm() / g.norm().clamp_min(1e-12)),
                         allclose=bool(torch.allclose(t, g, rtol=rtol, atol=atol)))
            else:
                r.update(allclose=False, shape_mismatch=list(g.shape))
        report[f] = r
    return report
