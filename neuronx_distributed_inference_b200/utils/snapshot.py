"""Input snapshots: dump the exact inputs (and optionally weights / outputs) of selected requests so a numerical or runtime
problem can be reproduced offline (reference utils/snapshot.py:58-481; env-driven registration application_base.py:423-554).

Environment (same names as the reference):
  NXD_INFERENCE_CAPTURE_SNAPSHOT=1        enable
  NXD_INFERENCE_SNAPSHOT_OUTPUT_PATH      directory (default ./snapshots)
  NXD_INFERENCE_SNAPSHOT_OUTPUT_FORMAT    "pt" (torch.save) or "npy" (one .npy per tensor)
  NXD_INFERENCE_SNAPSHOT_AT_REQUESTS      comma list of request indices to capture (default 0)
  NXD_INFERENCE_SNAPSHOT_FOR_TOKENS       comma list of token-generation step indices to capture (default: all of a request)
Layout: <path>/<sub-model tag>/request<N>/[step<K>/]rank<R>/{inputs.pt|*.npy} (+ ``weights.pt`` with ``save_weights``)."""
from __future__ import annotations

import os
from typing import Dict, Iterable, Optional

import torch


class SnapshotOutputFormat:
    PT = "pt"
    NPY = "npy"


def _ints(s: Optional[str], default):
    if not s:
        return default
    return [int(x) for x in s.split(",") if x.strip()]


class SnapshotHook:
    def __init__(self, tag: str, output_path: str, fmt: str = "pt", at_requests: Iterable[int] = (0,),
                 for_tokens: Optional[Iterable[int]] = None, rank: int = 0, save_weights: bool = False, model=None):
        self.tag, self.path, self.fmt = tag, output_path, fmt
        self.at_requests = set(at_requests)
        self.for_tokens = set(for_tokens) if for_tokens is not None else None
        self.rank, self.save_weights, self.model = rank, save_weights, model
        self.request = -1
        self.step = 0

    def new_request(self):
        self.request += 1
        self.step = 0

    def __call__(self, inputs: Dict[str, torch.Tensor], is_prefill: bool):
        if is_prefill:
            self.new_request()
        else:
            self.step += 1
        if self.request not in self.at_requests:
            return None
        if not is_prefill and self.for_tokens is not None and self.step not in self.for_tokens:
            return None
        d = os.path.join(self.path, self.tag, f"request{self.request}")
        if not is_prefill:
            d = os.path.join(d, f"step{self.step}")
        d = os.path.join(d, f"rank{self.rank}")
        os.makedirs(d, exist_ok=True)
        cpu = {k: v.detach().cpu() for k, v in inputs.items() if torch.is_tensor(v)}
        if self.fmt == SnapshotOutputFormat.NPY:
            import numpy as np
            for k, v in cpu.items():
                np.save(os.path.join(d, f"{k}.npy"), v.float().numpy() if v.dtype == torch.bfloat16 else v.numpy())
        else:
            torch.save(cpu, os.path.join(d, "inputs.pt"))
        if self.save_weights and self.model is not None and is_prefill:
            torch.save({k: v.detach().cpu() for k, v in self.model.state_dict().items()}, os.path.join(d, "weights.pt"))
        return d


def register_snapshot_hooks(app, output_path: Optional[str] = None, fmt: Optional[str] = None, at_requests=None, for_tokens=None,
                            save_weights: bool = False):
    """Attach a :class:`SnapshotHook` to every sub-model runner of ``app`` (env vars supply defaults)."""
    env = os.environ
    output_path = output_path or env.get("NXD_INFERENCE_SNAPSHOT_OUTPUT_PATH", "./snapshots")
    fmt = fmt or env.get("NXD_INFERENCE_SNAPSHOT_OUTPUT_FORMAT", "pt")
    at_requests = at_requests if at_requests is not None else _ints(env.get("NXD_INFERENCE_SNAPSHOT_AT_REQUESTS"), [0])
    for_tokens = for_tokens if for_tokens is not None else _ints(env.get("NXD_INFERENCE_SNAPSHOT_FOR_TOKENS"), None)
    from ..parallel.state import get_tensor_model_parallel_group
    rank = get_tensor_model_parallel_group().rank
    hooks = []
    shared = {"request": -1}
    for r in app.models:
        h = SnapshotHook(r.tag, output_path, fmt, at_requests, for_tokens, rank, save_weights, app.model)
        r.snapshot_hook = h
        hooks.append(h)
    # request counters advance together: a prefill on the CTE runner starts a new request for every runner
    def on_prefill():
        for h in hooks:
            h.new_request()
    for r in app.models:
        r.on_new_request = on_prefill
    return hooks


def maybe_register_from_env(app):
    if os.environ.get("NXD_INFERENCE_CAPTURE_SNAPSHOT", "0") in ("1", "true", "True"):
        return register_snapshot_hooks(app)
    return []
