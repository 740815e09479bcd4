"""Bucketing processor (reference experimental/core/processor/bucketing_processor.py:16-214): route a request to the smallest built
entry that fits and right-pad ``(input_tokens, attention_mask)`` to its shapes.

Two uses of one class: ``BucketingProcessor(model, pad_token_id)`` wraps a built model (entries from ``model.reserved_example_inputs``),
``BucketingProcessor([64, 128], pad_token_id)`` is the stand-alone padder returning ``(tokens, mask, bucket)``."""
from __future__ import annotations

import bisect
from typing import Dict, List, Tuple

import torch

from ..pad import pad_to_shape


def collect_buckets(reserved_example_inputs: Dict[str, tuple]):
    """tags ``prefill*`` / ``decode*`` -> ({kind: {kv_len: example inputs}}, {kind: sorted kv_lens}); the third example input is the
    attention mask ``[batch, kv_len]``."""
    buckets = {"prefill": {}, "decode": {}}
    for tag, inputs in reserved_example_inputs.items():
        kind = "prefill" if tag.startswith("prefill") else "decode" if tag.startswith("decode") else None
        if kind is None:
            raise ValueError(f"model tag {tag!r} must start with 'prefill' or 'decode'")
        buckets[kind][int(inputs[2].shape[1])] = inputs
    return buckets, {k: sorted(v) for k, v in buckets.items()}


def get_buckets_by_model_type(tokens: torch.Tensor, buckets: dict, bucket_table: dict):
    kind = "prefill" if tokens.shape[1] > 1 else "decode"
    return buckets[kind], bucket_table[kind]


def select_smallest_bucket(bucket_choices: dict, bucket_table: List[int], cur_len: int):
    i = bisect.bisect_left(bucket_table, cur_len)
    if i == len(bucket_table):
        raise ValueError(f"no bucket holds {cur_len} positions; available: {bucket_table}")
    return bucket_choices[bucket_table[i]]


class BucketingProcessor(torch.nn.Module):
    def __init__(self, model, pad_token_id: int = 0):
        super().__init__()
        self.pad_token_id = pad_token_id
        if isinstance(model, (list, tuple)):
            self.model, self.plain_buckets = None, sorted(int(b) for b in model)
        else:
            self.model, self.plain_buckets = model, None
            self.buckets, self.bucket_table = collect_buckets(model.reserved_example_inputs)

    # ---- stand-alone padder ------------------------------------------------------------------------------------------
    def select(self, n: int) -> int:
        i = bisect.bisect_left(self.plain_buckets, n)
        if i == len(self.plain_buckets):
            raise ValueError(f"sequence of {n} tokens exceeds the largest bucket {self.plain_buckets[-1]}")
        return self.plain_buckets[i]

    def _pad_plain(self, tokens, attention_mask) -> Tuple[torch.Tensor, torch.Tensor, int]:
        B, T = tokens.shape
        b = self.select(T)
        if attention_mask is None:
            attention_mask = torch.ones_like(tokens)
        return (pad_to_shape(tokens, (B, b), value=self.pad_token_id), pad_to_shape(attention_mask, (B, b), value=0), b)

    # ---- model wrapper ------------------------------------------------------------------------------------------------
    def pre_process(self, input_tokens, last_pos, attention_mask):
        """Prefill: tokens and mask padded to the smallest prefill entry >= the prompt width.  Decode: the mask padded to the smallest
        decode entry that holds ``max(last_pos) + 1`` positions."""
        choices, table = get_buckets_by_model_type(input_tokens, self.buckets, self.bucket_table)
        need = input_tokens.shape[1] if input_tokens.shape[1] > 1 else int(last_pos.max()) + 1
        ex_tokens, _, ex_mask = select_smallest_bucket(choices, table, max(need, 1))
        B = input_tokens.shape[0]
        if B > ex_tokens.shape[0]:
            raise ValueError(f"batch {B} exceeds the built batch {ex_tokens.shape[0]}")
        tokens = pad_to_shape(input_tokens, ex_tokens.shape, value=self.pad_token_id)
        mask = attention_mask[:, : ex_mask.shape[1]] if attention_mask.shape[1] > ex_mask.shape[1] else attention_mask
        mask = pad_to_shape(mask, ex_mask.shape, value=0)
        last = pad_to_shape(last_pos, (ex_tokens.shape[0],), value=0)
        return tokens, last, mask

    def forward(self, input_tokens, last_pos=None, attention_mask=None, **kwargs):
        if self.model is None:
            return self._pad_plain(input_tokens, attention_mask if attention_mask is not None else last_pos)
        B = input_tokens.shape[0]
        tokens, last, mask = self.pre_process(input_tokens, last_pos, attention_mask)
        out = self.model(tokens, last, mask)
        return out[:B]

    def reset(self):
        if self.model is not None and hasattr(self.model, "reset"):
            self.model.reset()
