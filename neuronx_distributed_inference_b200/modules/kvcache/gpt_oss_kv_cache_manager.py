"""Dual (global + rolling sliding-window) KV cache for models that mix full-attention and sliding-window layers
(GPT-OSS, Gemma-2/3, Cohere2, EXAONE-4, all-SWA Mistral): global layers keep ``max_len`` slots per line, sliding layers keep only
``window`` slots and are written modulo the window.  For GPT-OSS-120B at 128k context this is ~half of the KV bytes
(18 of 36 layers shrink from 131072 to 128 slots).

reference: modules/kvcache/gpt_oss_kv_cache_manager.py:30-396 (two ParameterLists, ``[B, H, W, D]`` for sliding layers,
``get_kv_by_layer_id`` picking by layer kind) and the "sliding-window modulo" update of kv_cache_manager.py:588-614.

B200 design: two contiguous allocations (stable pointers under CUDA graphs), one per layer kind.  Keys are stored post-RoPE, so
attention over a rolling line does not care about slot order; a decode step therefore runs the ordinary flash-decode kernel on the
``[.., W, D]`` view with horizon ``min(pos, W - 1)`` and no window (every resident slot is inside the window by construction).
Only the LAST ``W`` tokens of a prompt are written at prefill (earlier ones would be overwritten anyway, and duplicate slots inside
one scatter would be order-dependent).  Not combinable with features that read a positional prefix back from the cache (prefix
caching, chunked prefill, windowed context encoding) or that write several speculative tokens per step.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from .kv_cache_manager import KVCacheManager


class HybridKVCacheManager(nn.Module):
    def __init__(self, layer_windows: Sequence[Optional[int]], num_kv_heads: int, head_dim: int, max_len: int, num_lines: int,
                 dtype=torch.bfloat16, device=None, quant_config=None, **kw):
        super().__init__()
        self.layer_windows = [int(w) if w and int(w) < max_len else None for w in layer_windows]
        wins = {w for w in self.layer_windows if w}
        if len(wins) > 1:
            raise NotImplementedError(f"one sliding window per model, got {sorted(wins)}")
        self.window = wins.pop() if wins else None
        self.num_layers, self.num_kv_heads, self.head_dim = len(self.layer_windows), num_kv_heads, head_dim
        self.max_len, self.num_lines = max_len, num_lines
        g_layers = [i for i, w in enumerate(self.layer_windows) if not w]
        s_layers = [i for i, w in enumerate(self.layer_windows) if w]
        self._slot = {}
        for j, i in enumerate(g_layers):
            self._slot[i] = ("full", j)
        for j, i in enumerate(s_layers):
            self._slot[i] = ("window", j)
        mk = lambda n, S: KVCacheManager(n, num_kv_heads, head_dim, S, num_lines, dtype, device, quant_config=quant_config, **kw)  # noqa: E731
        self.full = mk(len(g_layers), max_len) if g_layers else None
        self.windowed = mk(len(s_layers), self.window) if s_layers else None
        ref = self.full if self.full is not None else self.windowed
        self.garbage, self.store_dtype, self.k_scale, self.v_scale = ref.garbage, ref.store_dtype, ref.k_scale, ref.v_scale
        self.quant_config = quant_config

    # ---- layer-kind helpers --------------------------------------------------------------------------------------------------
    def _mgr(self, layer: int) -> Tuple[KVCacheManager, int]:
        kind, j = self._slot[layer]
        return (self.full if kind == "full" else self.windowed), j

    def rolling_window(self, layer: int) -> Optional[int]:
        """Window size if ``layer`` lives in the rolling cache, else None."""
        return self.layer_windows[layer]

    @staticmethod
    def rolling_positions(window: int, write_positions: torch.Tensor, position_ids: torch.Tensor, is_prefill: bool):
        """(slots to write, attention horizon) for a rolling line.  Negative write positions stay negative (skipped).
        Prefill keeps only the last ``window`` tokens of every row."""
        wp = write_positions
        if is_prefill:
            last = wp.max(dim=1, keepdim=True).values
            keep = (wp >= 0) & (wp > last - window)
        else:
            if wp.shape[1] != 1:
                raise NotImplementedError("rolling sliding-window caches take one new token per decode step")
            keep = wp >= 0
        slots = torch.where(keep, wp % window, torch.full_like(wp, -1))
        return slots, position_ids.clamp(max=window - 1)

    # ---- KVCacheManager surface ------------------------------------------------------------------------------------------------
    @property
    def past_key_values(self) -> List[torch.Tensor]:
        out = []
        for i in range(self.num_layers):
            out += list(self.get_kv_by_layer_id(i))
        return out

    def get_kv_by_layer_id(self, idx: int):
        m, j = self._mgr(idx)
        return m.get_kv_by_layer_id(j)

    def get_cache(self, seq_len: Optional[int] = None):
        return [self.get_kv_by_layer_id(i) for i in range(self.num_layers)]

    def reset(self):
        for m in (self.full, self.windowed):
            if m is not None:
                m.reset()

    def lines_for(self, seq_ids):
        return (self.full if self.full is not None else self.windowed).lines_for(seq_ids)

    def update(self, layer: int, k_new, v_new, seq_ids, positions, lines=None):
        """``positions`` are cache SLOTS: callers convert through :meth:`rolling_positions` for rolling layers."""
        m, j = self._mgr(layer)
        m.update(j, k_new, v_new, seq_ids, positions, lines)

    def move(self, seq_ids, src, dst):
        raise NotImplementedError("token-tree KV compaction is not defined on a rolling cache")

    def bytes(self) -> int:
        return sum(m.bytes() for m in (self.full, self.windowed) if m is not None)


GptOssKVCacheManager = HybridKVCacheManager        # the reference's name for the dual sliding / global manager
