"""reference experimental/functional/parallel/tensor_ops.py:9-57."""
from __future__ import annotations

import torch


def split_along_dim(x: torch.Tensor, dim: int, rank: int, num_partitions: int) -> torch.Tensor:
    """The ``rank``-th of ``num_partitions`` equal contiguous slices of ``x`` along ``dim`` (a view)."""
    n = x.shape[dim]
    if n % num_partitions != 0:
        raise ValueError(f"dimension {dim} of size {n} does not split into {num_partitions} equal parts")
    if not 0 <= int(rank) < num_partitions:
        raise ValueError(f"rank {rank} outside [0, {num_partitions})")
    per = n // num_partitions
    return x.narrow(dim, int(rank) * per, per)
