"""Padding helpers of the experimental flow (reference experimental/core/pad.py:9-58): always pad at the END of a dimension."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def pad_at_end(input: torch.Tensor, dim: int, padded_len: int, mode: str = "constant", value=None) -> torch.Tensor:
    if not (0 <= dim < input.ndim):
        raise ValueError(f"dim {dim} out of range for a {input.ndim}-d tensor")
    cur = input.shape[dim]
    if padded_len < cur:
        raise ValueError(f"cannot pad dim {dim} of length {cur} down to {padded_len}")
    if padded_len == cur:
        return input
    spec = [0, 0] * (input.ndim - 1 - dim) + [0, padded_len - cur]      # F.pad lists the LAST dimension first
    return F.pad(input, spec, mode=mode, value=value)


def pad_to_shape(input: torch.Tensor, expected_shape, mode: str = "constant", value=None) -> torch.Tensor:
    if tuple(input.shape) == tuple(expected_shape):
        return input
    for dim, n in enumerate(expected_shape):
        input = pad_at_end(input, dim, n, mode, value)
    return input
