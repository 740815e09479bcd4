"""Adaptive layer norms of the MMDiT blocks (reference models/diffusers/normalization.py)."""
from __future__ import annotations

import torch
import torch.nn as nn


class AdaLayerNormZero(nn.Module):
    """LayerNorm(no affine) modulated by 6 chunks of ``Linear(silu(emb))``: shift/scale/gate for attention and MLP."""

    def __init__(self, dim, n_chunks, dtype, device):
        super().__init__()
        self.linear = nn.Linear(dim, n_chunks * dim, dtype=dtype, device=device)
        self.n = n_chunks
        self.dim = dim

    def forward(self, x, emb):
        mods = self.linear(nn.functional.silu(emb)).unsqueeze(1).chunk(self.n, -1)
        xn = nn.functional.layer_norm(x, (self.dim,), eps=1e-6)
        return xn * (1 + mods[1]) + mods[0], mods[2:]


class AdaLayerNormContinuous(nn.Module):
    def __init__(self, dim, cond_dim, dtype, device):
        super().__init__()
        self.linear = nn.Linear(cond_dim, 2 * dim, dtype=dtype, device=device)
        self.dim = dim

    def forward(self, x, emb):
        scale, shift = self.linear(nn.functional.silu(emb)).unsqueeze(1).chunk(2, -1)
        return nn.functional.layer_norm(x, (self.dim,), eps=1e-6) * (1 + scale) + shift
