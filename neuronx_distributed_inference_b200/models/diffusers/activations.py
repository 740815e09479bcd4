"""Activation modules of the diffusion stack (reference models/diffusers/activations.py:28-98): the GELU feed-forward input projection
(column-parallel linear + GELU, optionally the tanh approximation), an fp32 SiLU for the timestep embedders, and the name -> module
lookup diffusers configs use."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from ...parallel.layers import ColumnParallelLinear


class NeuronGELU(nn.Module):
    def __init__(self, dim_in: int, dim_out: int, approximate: str = "none", bias: bool = True, dtype=torch.float32, device=None):
        super().__init__()
        self.proj = ColumnParallelLinear(dim_in, dim_out, bias=bias, gather_output=False, dtype=dtype, device=device)
        self.approximate = approximate

    def gelu(self, gate: torch.Tensor) -> torch.Tensor:
        return F.gelu(gate, approximate=self.approximate)

    def forward(self, hidden_states):
        return self.gelu(self.proj(hidden_states))


class FP32SiLU(nn.Module):
    """SiLU evaluated in fp32 whatever the input dtype (timestep / guidance embedders are precision sensitive)."""

    def forward(self, inputs: torch.Tensor) -> torch.Tensor:
        return F.silu(inputs.float()).to(inputs.dtype)


_ACTIVATIONS = {"swish": nn.SiLU, "silu": nn.SiLU, "mish": nn.Mish, "gelu": nn.GELU, "relu": nn.ReLU, "silu_fp32": FP32SiLU}


def get_activation(act_fn: str) -> nn.Module:
    try:
        return _ACTIVATIONS[act_fn.lower()]()
    except KeyError:
        raise ValueError(f"Unsupported activation function: {act_fn}") from None
