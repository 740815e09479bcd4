"""Gated MLP (SwiGLU / GeGLU) with fused gate+up projection.

reference: ``NeuronLlamaMLP`` (models/llama/modeling_llama.py:300-737).  There the fused kernel
(K5) does residual-add + RMSNorm + gate/up + act*mul + down, followed by a separate all-reduce;
here:   act = W_gate_up(rmsnorm(h))  [norm prologue, SwiGLU epilogue, one kernel]
        h  += W_down(act) [+all-reduce fused]                         [one kernel]
State-dict key ``gate_up_proj.weight`` = [gate; up] stacked on dim 0 (unsharded), sharded with
``partition_stride=2`` so every rank holds [gate_r; up_r].
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from .. import ops
from ..parallel.layers import ColumnParallelLinear, RowParallelLinear
from ..parallel.state import Group

_GLU_ACT = {"silu": "silu_mul", "swish": "silu_mul", "gelu": "gelu_mul", "gelu_pytorch_tanh": "gelu_tanh_mul",
            "gelu_new": "gelu_tanh_mul", "gelu_tanh": "gelu_tanh_mul"}
_PLAIN_ACT = {"silu": "silu", "gelu": "gelu", "gelu_pytorch_tanh": "gelu_tanh", "gelu_new": "gelu_tanh",
              "relu": "relu", "relu2": "relu2"}


class GatedMLP(nn.Module):
    def __init__(self, hidden_size: int, intermediate_size: int, hidden_act: str = "silu", dtype=torch.float32,
                 bias: bool = False, tp_group: Optional[Group] = None, device=None,
                 sequence_parallel_enabled: bool = False, reduce_dtype=None):
        super().__init__()
        self.act = _GLU_ACT[hidden_act]
        self.gate_up_proj = ColumnParallelLinear(hidden_size, 2 * intermediate_size, bias=bias, gather_output=False,
                                                 dtype=dtype, device=device, tensor_model_parallel_group=tp_group,
                                                 stride=2, sequence_parallel_enabled=sequence_parallel_enabled)
        self.down_proj = RowParallelLinear(intermediate_size, hidden_size, bias=bias, input_is_parallel=True,
                                           dtype=dtype, device=device, tensor_model_parallel_group=tp_group,
                                           sequence_parallel_enabled=sequence_parallel_enabled,
                                           reduce_dtype=reduce_dtype)

    def forward(self, x, norm_weight=None, norm_eps=1e-6, norm_offset=0.0, residual=None, lora=None, adapter_ids=None):
        if lora is not None and adapter_ids is not None and (lora.has("gate_up_proj") or lora.has("down_proj")):
            # LoRA deltas enter before the activation / before the reduction: take the un-fused route
            xn = ops.rmsnorm(x, norm_weight, norm_eps, norm_offset) if norm_weight is not None else x
            gu = self.gate_up_proj(xn) + lora("gate_up_proj", xn, adapter_ids)
            h = ops.activation(gu, self.act)
            y = self.down_proj(h, residual)
            d = lora("down_proj", h, adapter_ids)
            if not isinstance(d, int):
                from ..parallel import mappings
                y = y + mappings.all_reduce(d, self.down_proj.tensor_parallel_group)
            return y
        h = self.gate_up_proj(x, norm_weight=norm_weight, norm_eps=norm_eps, norm_offset=norm_offset, act=self.act)
        return self.down_proj(h, residual)


class PlainMLP(nn.Module):
    """fc1 -> act -> fc2 (Whisper / CLIP / T5 / vision towers)."""

    def __init__(self, hidden_size, intermediate_size, hidden_act="gelu", dtype=torch.float32, bias=True,
                 tp_group=None, device=None):
        super().__init__()
        self.act = _PLAIN_ACT[hidden_act]
        self.fc1 = ColumnParallelLinear(hidden_size, intermediate_size, bias=bias, gather_output=False, dtype=dtype,
                                        device=device, tensor_model_parallel_group=tp_group)
        self.fc2 = RowParallelLinear(intermediate_size, hidden_size, bias=bias, dtype=dtype, device=device,
                                     tensor_model_parallel_group=tp_group)

    def forward(self, x, residual=None):
        return self.fc2(ops.activation(self.fc1(x), self.act), residual)
