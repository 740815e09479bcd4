"""YaRN helpers of DeepSeek-V3 under the reference's import path (models/deepseek/rope_util.py); the engine's rotary tables come from
``modules/rope.py`` (``compute_inv_freq`` handles the ``yarn`` scaling dict, mscale included)."""
from __future__ import annotations

import math

import torch

from ...modules.rope import RotaryEmbedding


def yarn_find_correction_dim(num_rotations: float, dim: int, base: float = 10000.0, max_position_embeddings: int = 2048) -> float:
    return (dim * math.log(max_position_embeddings / (num_rotations * 2 * math.pi))) / (2 * math.log(base))


def yarn_find_correction_range(low_rot: float, high_rot: float, dim: int, base: float = 10000.0, max_position_embeddings: int = 2048):
    low = math.floor(yarn_find_correction_dim(low_rot, dim, base, max_position_embeddings))
    high = math.ceil(yarn_find_correction_dim(high_rot, dim, base, max_position_embeddings))
    return max(low, 0), min(high, dim - 1)


def yarn_get_mscale(scale: float = 1.0, mscale: float = 1.0) -> float:
    return 1.0 if scale <= 1 else 0.1 * mscale * math.log(scale) + 1.0


def yarn_linear_ramp_mask(lo: float, hi: float, dim: int) -> torch.Tensor:
    if lo == hi:
        hi += 0.001
    return ((torch.arange(dim, dtype=torch.float32) - lo) / (hi - lo)).clamp(0, 1)


class DeepseekV3YarnRotaryEmbedding(RotaryEmbedding):
    def __init__(self, dim, max_position_embeddings=2048, base=10000.0, scaling_factor=1.0, original_max_position_embeddings=4096,
                 beta_fast=32, beta_slow=1, mscale=1, mscale_all_dim=0, device=None):
        super().__init__(dim, max_position_embeddings, base,
                         dict(rope_type="yarn", factor=scaling_factor, original_max_position_embeddings=original_max_position_embeddings,
                              beta_fast=beta_fast, beta_slow=beta_slow, mscale=mscale, mscale_all_dim=mscale_all_dim), device)


DeepseekV3RotaryEmbedding = RotaryEmbedding
