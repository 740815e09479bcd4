"""Pipeline-level helpers for diffusion applications (reference utils/diffusers_adapter.py): latency wrapper + image export."""
from __future__ import annotations

import time

import torch


def to_uint8_images(img: torch.Tensor):
    """[B,3,H,W] in [0,1] -> list of HxWx3 uint8 arrays."""
    x = (img.clamp(0, 1) * 255).round().to(torch.uint8).permute(0, 2, 3, 1).cpu().numpy()
    return [x[i] for i in range(x.shape[0])]


def timed_generate(app, *a, **kw):
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = app(*a, **kw)
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    return out, time.perf_counter() - t0
