"""Batch split of an activation for attention data parallelism (reference experimental/functional/attention/data_parallel.py:8-32): DP rank
``d`` of a ``world_size / dp_degree``-wide tensor-parallel mesh keeps batch rows ``[d * B / dp, (d + 1) * B / dp)``."""
from __future__ import annotations

import torch

from ..parallel.tensor_ops import split_along_dim
from ..pg.data_parallel import get_dp_rank


def split_input_for_data_parallel(x: torch.Tensor, dim: int, world_size: int, dp_degree: int, rank) -> torch.Tensor:
    """``rank``: this process's global rank (int or 0-d / 1-element tensor, the reference's SPMDRank value)."""
    d = int(get_dp_rank(torch.as_tensor(rank), world_size, dp_degree))
    return split_along_dim(x, dim, d, dp_degree)
