"""Bucketing processor (reference experimental/core/processor/bucketing_processor.py + build_flow/bucketing_on_seq_len.py): pick the
smallest sequence bucket that fits and right-pad ``(tokens, attention_mask)`` to it."""
from __future__ import annotations

from typing import List, Tuple

import torch


class BucketingProcessor:
    def __init__(self, buckets: List[int], pad_token_id: int = 0):
        self.buckets = sorted(buckets)
        self.pad = pad_token_id

    def select(self, n: int) -> int:
        for b in self.buckets:
            if b >= n:
                return b
        raise ValueError(f"sequence of {n} tokens exceeds the largest bucket {self.buckets[-1]}")

    def __call__(self, tokens: torch.Tensor, attention_mask: torch.Tensor = None) -> Tuple[torch.Tensor, torch.Tensor, int]:
        B, T = tokens.shape
        b = self.select(T)
        if attention_mask is None:
            attention_mask = torch.ones_like(tokens)
        if b > T:
            tokens = torch.cat([tokens, tokens.new_full((B, b - T), self.pad)], 1)
            attention_mask = torch.cat([attention_mask, attention_mask.new_zeros(B, b - T)], 1)
        return tokens, attention_mask, b
