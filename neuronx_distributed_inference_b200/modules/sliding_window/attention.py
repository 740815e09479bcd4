"""Sliding-window flash attention entry point (reference modules/sliding_window/attention.py:235-477 ``flash_fwd``): the window
is a parameter of the engine's flash kernels (prefill and decode), not a separate kernel."""
from ... import ops


def flash_fwd(q, k, v, window_size: int, softmax_scale=None, causal: bool = True):
    """q [B,T,Hq,D], k/v [B,T,Hkv,D] -> [B,T,Hq,D]; key j visible iff j <= i and j > i - window_size."""
    return ops.attention_prefill(q, k, v, softmax_scale if softmax_scale is not None else q.shape[-1] ** -0.5, causal, window_size)
