"""reference models/qwen3_vl/utils/slicing.py: per-image slices of the flat patch sequence."""
from __future__ import annotations

import torch


def slice_by_image_hw(x: torch.Tensor, image_grid_thw: torch.Tensor):
    """x ``[sum(t*h*w), ...]`` -> list of per-image tensors ``[t*h*w, ...]``."""
    sizes = (image_grid_thw[:, 0] * image_grid_thw[:, 1] * image_grid_thw[:, 2]).tolist()
    return list(torch.split(x, [int(s) for s in sizes], 0))
