"""reference experimental/functional/pg/data_parallel.py:6-33 — attention data parallelism uses the same mesh shape as context
parallelism (``dp_degree`` contiguous TP groups)."""
from __future__ import annotations

import torch


def get_dp_rank(rank: torch.Tensor, world_size: int, dp_degree: int) -> torch.Tensor:
    if dp_degree < 1 or world_size % dp_degree != 0:
        raise ValueError(f"dp_degree {dp_degree} must divide world_size {world_size}")
    return torch.div(torch.as_tensor(rank), world_size // dp_degree, rounding_mode="floor").to(torch.int32)
