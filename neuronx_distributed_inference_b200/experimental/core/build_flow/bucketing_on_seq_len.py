"""Build flow with several sequence-length buckets (reference experimental/core/build_flow/bucketing_on_seq_len.py:11-89): entries
``prefill_<L>`` (tokens and mask ``[B, L]``) and ``decode_<L>`` (one token against an ``L``-wide mask)."""
from __future__ import annotations

from typing import List

import torch


def build_for_bucketing_on_seq_len(model: torch.nn.Module, world_size: int = -1, batch_size: int = 1, prefill_buckets: List[int] = (1024,),
                                   decode_buckets: List[int] = (1024,), cuda_graphs: bool = False):
    from ..functions import BuiltModel
    built = BuiltModel(model, cuda_graphs)
    ones = lambda *s: torch.ones(s, dtype=torch.int32)      # noqa: E731
    for L in sorted(set(prefill_buckets)):
        built.trace(dict(tokens=ones(batch_size, L), last_pos=torch.zeros(batch_size, dtype=torch.int32), attention_mask=ones(batch_size, L)),
                    tag=f"prefill_{L}")
    for L in sorted(set(decode_buckets)):
        built.trace(dict(tokens=ones(batch_size, 1), last_pos=torch.zeros(batch_size, dtype=torch.int32), attention_mask=ones(batch_size, L)),
                    tag=f"decode_{L}")
    return built
