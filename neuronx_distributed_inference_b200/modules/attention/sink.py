"""Learned attention sinks (GPT-OSS; reference modules/attention/sink.py): one learned logit per attention head that joins the softmax
denominator without contributing a value.  ``AttentionBase(learned_sinks=True)`` owns the parameter directly; this module is the
reference's stand-alone spelling of the same thing — one fp32 scalar per LOCAL q head, sharded along the head plan of the TP group."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from ...parallel.state import Group, get_tensor_model_parallel_group


class LearnedSink(nn.Module):
    def __init__(self, learned_sinks_size: int, num_attention_heads: int, torch_dtype=torch.float32, tensor_model_parallel_size: Optional[int] = None,
                 tp_group: Optional[Group] = None, device=None):
        super().__init__()
        assert learned_sinks_size == 1, f"Learned sinks only supports learned_sinks_size == 1 ({learned_sinks_size})"
        g = tp_group or get_tensor_model_parallel_group()
        tp = tensor_model_parallel_size or g.size
        assert num_attention_heads % tp == 0
        self.sink = nn.Parameter(torch.zeros(num_attention_heads // tp, dtype=torch_dtype, device=device), requires_grad=False)
        self.sink.partition_dim, self.sink.tp_group = 0, g

    def get_sink(self) -> torch.Tensor:
        return self.sink

    def forward(self, scores: torch.Tensor) -> torch.Tensor:
        """scores [B, H_local, T, S] (already scaled / masked) -> softmax probabilities over the S real keys with the sink in the
        denominator (rows no longer sum to one)."""
        s = self.sink.float().view(1, -1, 1, 1).expand(scores.shape[0], -1, scores.shape[2], 1)
        return torch.softmax(torch.cat([scores.float(), s], -1), -1)[..., :-1]
