"""Timestep / rotary embeddings for diffusion transformers (reference models/diffusers/embeddings.py)."""
from __future__ import annotations

import math

import torch
import torch.nn as nn


def timestep_sinusoid(t: torch.Tensor, dim: int = 256, max_period: float = 10000.0, flip_sin_to_cos: bool = True,
                      downscale_freq_shift: float = 0.0) -> torch.Tensor:
    half = dim // 2
    exponent = -math.log(max_period) * torch.arange(half, dtype=torch.float32, device=t.device) / (half - downscale_freq_shift)
    ang = t.float()[:, None] * exponent.exp()[None]
    emb = torch.cat([ang.sin(), ang.cos()], -1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], -1)
    return emb


class TimestepEmbedding(nn.Module):
    def __init__(self, in_dim, dim, dtype, device):
        super().__init__()
        self.linear_1 = nn.Linear(in_dim, dim, dtype=dtype, device=device)
        self.linear_2 = nn.Linear(dim, dim, dtype=dtype, device=device)

    def forward(self, x):
        return self.linear_2(nn.functional.silu(self.linear_1(x.to(self.linear_1.weight.dtype))))


class CombinedTimestepGuidanceTextProjEmbeddings(nn.Module):
    def __init__(self, dim, pooled_dim, guidance: bool, dtype, device):
        super().__init__()
        self.timestep_embedder = TimestepEmbedding(256, dim, dtype, device)
        self.guidance_embedder = TimestepEmbedding(256, dim, dtype, device) if guidance else None
        self.text_embedder = TimestepEmbedding(pooled_dim, dim, dtype, device)

    def forward(self, timestep, guidance, pooled):
        e = self.timestep_embedder(timestep_sinusoid(timestep))
        if self.guidance_embedder is not None and guidance is not None:
            e = e + self.guidance_embedder(timestep_sinusoid(guidance))
        return e + self.text_embedder(pooled)


def rope_nd(ids: torch.Tensor, axes_dim, theta: float = 10000.0):
    """ids [N, n_axes] -> (cos, sin) [N, sum(axes_dim)/2] for interleaved (pairwise) rotation."""
    cos, sin = [], []
    for i, d in enumerate(axes_dim):
        freqs = 1.0 / (theta ** (torch.arange(0, d, 2, dtype=torch.float64, device=ids.device) / d))
        ang = ids[:, i].double()[:, None] * freqs[None]
        cos.append(ang.cos().float())
        sin.append(ang.sin().float())
    return torch.cat(cos, -1), torch.cat(sin, -1)
