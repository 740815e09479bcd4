"""Profiling helpers — the role of the reference's ``neuron-profile capture/view`` wrapper (utils/profiling.py:34-122).
On B200: NVTX ranges per layer/op for Nsight, a CUPTI kernel timeline of one CUDA-graph replay (tools/trace_decode.py),
and command builders for the ncu recipes of the profiling guide."""
from __future__ import annotations

import contextlib
import shlex
from typing import List

import torch


@contextlib.contextmanager
def nvtx_range(name: str):
    if torch.cuda.is_available():
        torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        if torch.cuda.is_available():
            torch.cuda.nvtx.range_pop()


def annotate_layers(model):
    """Wrap every decoder layer's forward in an NVTX range ``layer<i>`` (the analogue of layer boundary markers,
    reference models/layer_boundary_marker.py:11-63)."""
    for i, layer in enumerate(getattr(model, "layers", [])):
        orig = layer.forward

        def fwd(*a, __orig=orig, __i=i, **k):
            with nvtx_range(f"layer{__i}"):
                return __orig(*a, **k)
        layer.forward = fwd
    return model


def ncu_launch_list_cmd(cmd: List[str], out_csv: str = "gpurun_out/launches.csv", skip: int = 0, count: int = 400) -> str:
    return " ".join(["ncu", "--metrics", "gpu__time_duration.sum", "--clock-control", "none", "-s", str(skip), "-c", str(count),
                     "--csv", "--log-file", out_csv] + [shlex.quote(c) for c in cmd])


def ncu_full_cmd(cmd: List[str], kernel_regex: str, out: str = "gpurun_out/prof", skip: int = 0, count: int = 3) -> str:
    return " ".join(["ncu", "--set", "full", "--clock-control", "none", "--import-source", "on", "-k", f"regex:{kernel_regex}",
                     "-s", str(skip), "-c", str(count), "-o", out, "-f"] + [shlex.quote(c) for c in cmd])


def kernel_timeline(fn, warmup: int = 3):
    """CUPTI timeline of ``fn()``: list of (kernel name, start_us, duration_us)."""
    from torch.profiler import ProfilerActivity, profile
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    t0 = evs[0].time_range.start if evs else 0
    return [(e.name, e.time_range.start - t0, e.time_range.end - e.time_range.start) for e in evs]
