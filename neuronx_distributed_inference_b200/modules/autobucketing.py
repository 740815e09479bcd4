"""Bucket ladders + bucket selection.

Behavioural parity with reference modules/autobucketing.py:8-341 and the selection rules of
models/model_wrapper.py:831-921.  On B200 a bucket is a CUDA-graph key (and the launch-geometry
hint for split-KV decode attention), not a separately compiled program: kernels read true
lengths from device memory, so padding to the bucket costs nothing but the padded tokens.
"""
from __future__ import annotations

from bisect import bisect_left
from math import log2
from typing import List, Sequence, Union

BUCKET_SELECTION_STRATEGIES = {"max", "first_fit", "second_fit"}


def generate_buckets(min_length: int, max_length: int) -> List[int]:
    """Powers of two from min_length up to (excluding ~) max_length, then max_length itself."""
    if min_length >= max_length:
        return [max_length]
    lo = int(log2(min_length))
    hi = round(log2(max_length))
    return [2 ** i for i in range(lo, hi)] + [max_length]


def generate_2d_buckets_for_prefix_caching(min_active, max_active, min_prefix, max_prefix, is_context_encode=False):
    """(active tokens x prefix length) grid; prefix 0 = "no cached prefix" for prefill."""
    act = generate_buckets(min_active, max_active)
    pre = generate_buckets(min_prefix, max_prefix)
    if is_context_encode:
        pre = [0] + pre
    return [[a, p] for a in act for p in pre]


def _2d_from_lists(act, pre, is_context_encode=False):
    pre = ([0] if is_context_encode else []) + list(pre)
    return [[a, p] for a in act for p in pre]


def generate_buckets_for_chunked_prefill_cte(config):
    """[chunk tokens, max blocks] pairs (reference autobucketing.py:65-146, simplified ladder)."""
    nc = config.neuron_config
    q_tile = nc.chunked_prefill_config.kernel_q_tile_size
    max_tok = nc.max_context_length
    toks = [t for t in generate_buckets(q_tile, max_tok)]
    blocks = generate_buckets(max(1, 128 // max(nc.pa_block_size, 1)), max(1, nc.seq_len // nc.pa_block_size))
    return [[t, b] for t in toks for b in blocks]


def generate_buckets_for_cte(config) -> list:
    nc = config.neuron_config
    if nc.is_chunked_prefill:
        return generate_buckets_for_chunked_prefill_cte(config)
    mcl = nc.max_context_length
    if not nc.enable_bucketing:
        return generate_2d_buckets_for_prefix_caching(mcl, mcl, mcl, mcl, True) if nc.is_prefix_caching \
            else generate_buckets(mcl, mcl)
    if nc.context_encoding_buckets is not None:
        if nc.is_prefix_caching:
            return _2d_from_lists(nc.context_encoding_buckets, nc.prefix_buckets or [mcl], True)
        return list(nc.context_encoding_buckets)
    if nc.is_prefix_caching:
        return generate_2d_buckets_for_prefix_caching(min(512, mcl), mcl, min(512, mcl), mcl, True)
    return generate_buckets(min(128, mcl), mcl)


def generate_2d_buckets_for_batch_bucketing(nc, seq_buckets):
    batches = set(nc.token_generation_batches) | {nc.tkg_batch_size}
    return [[b, s] for b in sorted(batches, reverse=True) for s in seq_buckets]


def _tkg_like(config, allow_batch=True):
    nc = config.neuron_config
    ml = nc.max_length
    if not nc.enable_bucketing:
        b = generate_2d_buckets_for_prefix_caching(1, 1, ml, ml) if nc.is_prefix_caching else generate_buckets(ml, ml)
    elif nc.token_generation_buckets is not None:
        b = [[1, i] for i in nc.token_generation_buckets] if nc.is_prefix_caching else list(nc.token_generation_buckets)
    elif nc.is_prefix_caching:
        b = generate_2d_buckets_for_prefix_caching(1, 1, min(256, ml), ml)
    else:
        b = generate_buckets(min(128, ml), ml)
    if allow_batch and nc.token_generation_batches is not None:
        if nc.is_prefix_caching:
            raise NotImplementedError("batch bucketing with prefix caching")
        return generate_2d_buckets_for_batch_bucketing(nc, b)
    return b


def generate_buckets_for_tkg(config):
    return _tkg_like(config, True)


def generate_buckets_for_fused_spec(config):
    return _tkg_like(config, False)


def generate_buckets_for_speculation(config):
    nc = config.neuron_config
    if not nc.enable_bucketing:
        return generate_buckets(nc.max_length, nc.max_length)
    if nc.token_generation_buckets is not None:
        return list(nc.token_generation_buckets)
    return generate_buckets(min(128, nc.max_length), nc.max_length)


def select_bucket(buckets: Sequence[int], length: int, speculation_length: int = 0,
                  strategy: str = "first_fit", allow_truncation: bool = False) -> int:
    """Index of the target bucket.  first_fit: first bucket with ``length + spec < bucket``
    (strict, as reference model_wrapper.py:903-921), falling back to the largest bucket when
    ``length`` equals it; second_fit: one above first fit (async look-ahead,
    async_execution.py:172-187); max: the largest."""
    assert strategy in BUCKET_SELECTION_STRATEGIES
    n = len(buckets)
    if strategy == "max":
        return n - 1
    need = length + speculation_length
    first = None
    for i, b in enumerate(buckets):
        if need < b:
            first = i
            break
    if first is None:
        if need == buckets[-1] or length <= buckets[-1] or allow_truncation:
            first = n - 1
        else:
            raise ValueError(f"input length {length} exceeds the largest bucket {buckets[-1]}")
    if strategy == "second_fit":
        first = min(first + 1, n - 1)
    return first


def select_prefill_bucket(buckets: Sequence[int], length: int, allow_truncation: bool = False) -> int:
    """Prefill pads *to* the bucket: smallest bucket >= length."""
    i = bisect_left(list(buckets), length)
    if i == len(buckets):
        if allow_truncation:
            return len(buckets) - 1
        raise ValueError(f"Inputs supplied ({length}) are longer than the largest bucket ({buckets[-1]}); "
                         "set allow_input_truncation to truncate")
    return i


def select_2d_bucket(buckets: Sequence[Sequence[int]], active: int, prefix: int) -> int:
    """Prefix caching: smallest (active, prefix) bucket covering the request; prefix bucket 0 only
    when there is no cached prefix (reference model_wrapper.py:923-1045)."""
    best, best_cost = None, None
    for i, (a, p) in enumerate(buckets):
        if a < active or p < prefix or (prefix > 0 and p == 0):
            continue
        cost = (a, p) if prefix else (p, a)
        if best is None or cost < best_cost:
            best, best_cost = i, cost
    if best is None:
        raise ValueError(f"no 2D bucket fits active={active} prefix={prefix}")
    return best
