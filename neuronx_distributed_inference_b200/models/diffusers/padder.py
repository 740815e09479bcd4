"""Weight padding helpers of the diffusion stack (reference models/diffusers/padder.py:10-128): make a head / channel count divisible by
the tensor-parallel degree by adding zero heads, either at the END of the dimension or INTERLEAVED (every group of real heads is
followed by its share of padding, so that a contiguous TP split gives every rank the same number of real heads)."""
from __future__ import annotations

from typing import Optional, Sequence, Union

import torch
import torch.nn.functional as F


def round_up_to_divisor(value: int, divisor: int) -> int:
    return -(-value // divisor) * divisor


def pad_sizes(shape: Sequence[int], dims: Union[int, Sequence[int]], sizes: Union[int, Sequence[int]], left: bool = False):
    """``F.pad`` argument that grows ``dims`` of a tensor of ``shape`` to ``sizes`` (zeros on the right, or left)."""
    dims = [dims] if isinstance(dims, int) else list(dims)
    sizes = [sizes] if isinstance(sizes, int) else list(sizes)
    spec = [0] * (2 * len(shape))
    for d, s in zip(dims, sizes):
        d = d % len(shape)
        grow = s - shape[d]
        assert grow >= 0, f"dim {d}: cannot pad {shape[d]} down to {s}"
        spec[2 * (len(shape) - 1 - d) + (0 if left else 1)] = grow
    return tuple(spec)


def pad(tensor: Optional[torch.Tensor], dims, sizes, left: bool = False):
    if tensor is None:
        return None
    return F.pad(tensor, pad_sizes(tensor.shape, dims, sizes, left))


def pad_interleaved(tensor: torch.Tensor, dim: int, size: int, source_len_per_group: int, pad_len_per_group: int) -> torch.Tensor:
    """``[g0 (source_len), 0 x pad_len, g1, 0 x pad_len, ...]`` along ``dim`` up to ``size``, e.g. [1,2,3] with (1, 2) ->
    [1,0,0,2,0,0,3,0,0]."""
    per = source_len_per_group + pad_len_per_group
    groups = size // per
    assert groups * source_len_per_group == tensor.shape[dim] and groups * per == size
    t = tensor.movedim(dim, 0)
    t = t.reshape(groups, source_len_per_group, *t.shape[1:])
    t = torch.cat([t, t.new_zeros(groups, pad_len_per_group, *t.shape[2:])], 1)
    return t.reshape(size, *t.shape[2:]).movedim(0, dim)


class MaybePadder:
    """Callable ``(weight, dim) -> padded weight``.  ``padding="end"``: zeros appended up to ``size``.  ``padding="interleaved"``:
    the dimension is first viewed as ``split_size`` units (heads) of equal width when ``split_size`` is given, then units are padded
    in ``interleaved_factor`` groups (one group per TP rank)."""

    def __init__(self, size: int, padding: str = "end", split_size: Optional[int] = None, interleaved_factor: Optional[int] = None):
        assert padding in ("end", "interleaved"), f"Invalid padding mode {padding}"
        self.size, self.padding, self.split_size, self.interleaved_factor = size, padding, split_size, interleaved_factor

    def __call__(self, weight: Optional[torch.Tensor], dim: int):
        if weight is None:
            return None
        if self.padding == "end":
            return pad(weight, dim, self.size)
        assert self.interleaved_factor, "interleaved_factor is not provided"
        dim = dim % weight.dim()
        n = weight.shape[dim]
        unit = 1
        if self.split_size:
            assert n % self.split_size == 0, f"dim {dim} of size {n} is not divisible by split_size {self.split_size}"
            unit = n // self.split_size
            weight = weight.reshape(*weight.shape[:dim], self.split_size, unit, *weight.shape[dim + 1:])
        units, target = weight.shape[dim], self.size // unit
        f = self.interleaved_factor
        assert units % f == 0 and (target - units) % f == 0
        out = pad_interleaved(weight, dim, target, units // f, (target - units) // f)
        return out.reshape(*out.shape[:dim], target * unit, *out.shape[dim + 2:]) if self.split_size else out
