"""Padding helpers (reference modules/padding.py:6-85)."""
from __future__ import annotations

from typing import Sequence

import torch
import torch.nn.functional as F


def pad_tensor(t: torch.Tensor, target_shape: Sequence[int], pad_value=0, left: bool = False):
    """Pad every dim of ``t`` up to ``target_shape``; returns (padded, original slices)."""
    assert t.dim() == len(target_shape)
    pads = []
    slices = []
    for d in reversed(range(t.dim())):
        extra = target_shape[d] - t.shape[d]
        assert extra >= 0, f"dim {d}: {t.shape[d]} > {target_shape[d]}"
        pads += [extra, 0] if left else [0, extra]
    for d in range(t.dim()):
        extra = target_shape[d] - t.shape[d]
        slices.append(slice(extra, None) if left else slice(0, t.shape[d]))
    return F.pad(t, pads, value=pad_value), tuple(slices)


def unpad_tensor(t: torch.Tensor, slices):
    return t[slices]


def pad_with_first_batchline(t: torch.Tensor, target_batch: int):
    """Grow the batch dim by repeating row 0 (reference ``repeat_first_batchline``)."""
    b = t.shape[0]
    if b >= target_batch:
        return t
    rep = t[:1].expand(target_batch - b, *t.shape[1:])
    return torch.cat([t, rep], 0)
