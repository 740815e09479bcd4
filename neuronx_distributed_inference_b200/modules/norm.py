"""Normalisation layers.  reference: modules/custom_calls.py:8-45 (CustomRMSNorm), Gemma offset
norm (models/gemma3), DBRX LayerNorm-without-bias."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops


class RMSNorm(nn.Module):
    def __init__(self, hidden_size: int, eps: float = 1e-6, dtype=torch.float32, offset: float = 0.0,
                 device=None):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(hidden_size, dtype=dtype, device=device), requires_grad=False)
        self.variance_epsilon = eps
        self.offset = offset

    def forward(self, x, residual=None):
        return ops.rmsnorm(x, self.weight, self.variance_epsilon, self.offset, residual)


class L2Norm(nn.Module):
    """Weight-less RMS norm (Llama-4 post-RoPE QK norm)."""

    def __init__(self, eps: float = 1e-6):
        super().__init__()
        self.eps = eps

    def forward(self, x):
        return ops.ref.rmsnorm(x, None, self.eps)


class LayerNorm(nn.LayerNorm):
    def __init__(self, hidden_size, eps=1e-5, bias=True, dtype=torch.float32, device=None):
        super().__init__(hidden_size, eps=eps, bias=bias, dtype=dtype, device=device)
        for p in self.parameters():
            p.requires_grad_(False)

    def forward(self, x):
        return nn.functional.layer_norm(x.float(), self.normalized_shape, self.weight.float(),
                                        None if self.bias is None else self.bias.float(), self.eps).to(x.dtype)


class IdentityNorm(nn.Module):
    """Placeholder with the RMSNorm interface (``weight is None`` => the fused projections skip the norm).  EAGLE draft
    models have no input norm on their first layer and no final norm (reference modeling_llama.py:892-899,1155)."""
    weight = None
    variance_epsilon = 1e-6
    offset = 0.0

    def forward(self, x, residual=None):
        return x if residual is None else x + residual
