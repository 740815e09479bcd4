"""KV-cache managers.

* :class:`KVCacheManager` — contiguous ``[lines(+garbage), H_kv/tp, S, D]`` per layer, exposed as the
  flat list ``[K0,V0,K1,V1,...]`` (reference modules/kvcache/kv_cache_manager.py:106-692).  Continuous
  batching selects lines by ``seq_ids``; masked lines (``-1`` with ``apply_seq_ids_mask``) are redirected
  to a garbage line (reference uses a 128-slot garbage zone, kv_cache_manager.py:25-26,230-231).
* :class:`BlockKVCacheManager` — paged ``[num_blocks+1, block_size, H, D]`` with ``slot_mapping`` /
  ``block_table`` (reference block_kv_cache_manager.py:11-431); block ``num_blocks`` is the reserved
  scratch block that ``-1`` slots are redirected to by the reference; our kernels skip ``-1`` instead.
* :class:`DataParallelKVCacheManager` — per-DP-rank seq_id range remap (data_parallel_kv_cache_manager.py:8-39).

B200 design: the cache is ONE allocation ``[layers, 2, lines, H, S, D]`` (views per layer) so a CUDA
graph sees stable pointers and ``reset()`` is one memset.  Updates are in place (no input/output aliasing
machinery as in model_wrapper.py:1548-1627).  fp8 KV (``kv_cache_quant``) stores e4m3 with static scales.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.nn as nn

from ... import ops


class KVCacheManager(nn.Module):
    def __init__(self, num_layers: int, num_kv_heads: int, head_dim: int, max_len: int, num_lines: int,
                 dtype=torch.bfloat16, device=None, garbage_line: bool = True, quant_config=None,
                 sliding_window: Optional[int] = None, v_head_dim: Optional[int] = None):
        super().__init__()
        self.num_layers, self.num_kv_heads, self.head_dim = num_layers, num_kv_heads, head_dim
        self.max_len, self.num_lines = max_len, num_lines
        self.garbage = 1 if garbage_line else 0
        self.quant_config = quant_config
        self.store_dtype = dtype
        self.k_scale = self.v_scale = None
        if quant_config is not None:
            from ...config import to_torch_dtype
            self.store_dtype = to_torch_dtype(quant_config.dtype)
            if quant_config.scale_mode != "direct_cast":
                self.k_scale = float(quant_config.k_scale)
                self.v_scale = float(quant_config.v_scale)
        self.v_head_dim = v_head_dim or head_dim
        L = num_lines + self.garbage
        if self.v_head_dim == head_dim:
            self.register_buffer("cache", torch.zeros(num_layers, 2, L, num_kv_heads, max_len, head_dim,
                                                      dtype=self.store_dtype, device=device), persistent=False)
            self._k = [self.cache[i, 0] for i in range(num_layers)]
            self._v = [self.cache[i, 1] for i in range(num_layers)]
        else:  # MLA-style asymmetric caches
            self.register_buffer("cache_k", torch.zeros(num_layers, L, num_kv_heads, max_len, head_dim,
                                                        dtype=self.store_dtype, device=device), persistent=False)
            self.register_buffer("cache_v", torch.zeros(num_layers, L, num_kv_heads, max_len, self.v_head_dim,
                                                        dtype=self.store_dtype, device=device), persistent=False)
            self._k = [self.cache_k[i] for i in range(num_layers)]
            self._v = [self.cache_v[i] for i in range(num_layers)]

    # reference-compatible view -------------------------------------------------------------
    @property
    def past_key_values(self) -> List[torch.Tensor]:
        out = []
        for i in range(self.num_layers):
            out += [self._k[i], self._v[i]]
        return out

    def get_kv_by_layer_id(self, idx: int) -> Tuple[torch.Tensor, torch.Tensor]:
        return self._k[idx], self._v[idx]

    def get_cache(self, seq_len: Optional[int] = None):
        return [self.get_kv_by_layer_id(i) for i in range(self.num_layers)]

    def reset(self):
        for b in self.buffers():
            b.zero_()

    def lines_for(self, seq_ids: torch.Tensor) -> torch.Tensor:
        """Map seq_ids to cache lines; negative / out-of-range ids go to the garbage line (or are
        skipped by the kernels when there is none)."""
        if self.garbage:
            bad = (seq_ids < 0) | (seq_ids >= self.num_lines)
            return torch.where(bad, torch.full_like(seq_ids, self.num_lines), seq_ids)
        return seq_ids

    def update(self, layer: int, k_new, v_new, seq_ids, positions, lines=None):
        """k_new/v_new [B,T,H,D]; positions [B,T] (negative => skip)."""
        k, v = self._k[layer], self._v[layer]
        if self.store_dtype != k_new.dtype:
            if self.k_scale is not None:
                k_new, v_new = k_new.float() / self.k_scale, v_new.float() / self.v_scale
            fi = torch.finfo(self.store_dtype)
            k_new = k_new.float().clamp(fi.min, fi.max).to(self.store_dtype)
            v_new = v_new.float().clamp(fi.min, fi.max).to(self.store_dtype)
            self._append_torch(k, v, k_new, v_new, self.lines_for(seq_ids) if lines is None else lines, positions)
        elif (k_new.is_cuda and k_new.dtype in (torch.bfloat16, torch.float16)
              and (self.head_dim * k_new.element_size()) % 16 == 0):
            ops.kv_append(k, v, k_new, v_new, self.lines_for(seq_ids) if lines is None else lines, positions)
        else:
            self._append_torch(k, v, k_new, v_new, self.lines_for(seq_ids) if lines is None else lines, positions)

    def _append_torch(self, k, v, k_new, v_new, lines, positions):
        """PyTorch cache write without data-dependent shapes (CUDA-graph capturable): skipped entries (masked line or
        padding position) are redirected to slot 0 of the garbage line instead of being filtered out."""
        if not self.garbage:
            return ops.ref.kv_append(k, v, k_new, v_new, lines, positions)
        B, T = positions.shape
        L, H, S, D = k.shape
        line = lines.view(B, 1).expand(B, T).reshape(-1).long()
        pos = positions.reshape(-1).long()
        ok = (line >= 0) & (line < L) & (pos >= 0) & (pos < S)
        line = torch.where(ok, line, torch.full_like(line, L - 1))
        pos = torch.where(ok, pos, torch.zeros_like(pos))
        k[line, :, pos] = k_new.reshape(B * T, H, D).to(k.dtype)
        v[line, :, pos] = v_new.reshape(B * T, H, v_new.shape[-1]).to(v.dtype)

    def move(self, seq_ids: torch.Tensor, src: torch.Tensor, dst: torch.Tensor):
        """Compact accepted tree nodes: for every layer copy slot ``src[b,j]`` to ``dst[b,j]`` of line ``seq_ids[b]``
        (negative src/dst => no-op, realised as a self-copy of slot 0 on the garbage line).  One gather + one scatter over
        the whole ``[layers, 2, ...]`` allocation; static shapes, capture-safe.  Role of the reference's accepted-index KV
        gather/scatter for Medusa / token trees (kv_cache_manager.py ``accepted_indices`` / ``current_length``)."""
        B, n = src.shape
        line = self.lines_for(seq_ids).view(B, 1).expand(B, n).long()
        ok = (src >= 0) & (dst >= 0)
        Lg = self.num_lines + self.garbage - 1
        line = torch.where(ok, line, torch.full_like(line, Lg)).reshape(-1)
        s = torch.where(ok, src, torch.zeros_like(src)).long().reshape(-1)
        d = torch.where(ok, dst, torch.zeros_like(dst)).long().reshape(-1)
        for c in ([self.cache] if hasattr(self, "cache") else [self.cache_k.unsqueeze(1), self.cache_v.unsqueeze(1)]):
            vals = c[:, :, line, :, s]              # [B*n, layers, 2, H, D] (advanced indices first)
            c[:, :, line, :, d] = vals

    def bytes(self) -> int:
        return sum(b.numel() * b.element_size() for b in self.buffers())


class DataParallelKVCacheManager(KVCacheManager):
    """Attention-DP decode: this DP rank holds lines for seq_ids in [dp_rank*n, (dp_rank+1)*n)."""

    def __init__(self, *a, dp_rank: int = 0, dp_size: int = 1, **kw):
        super().__init__(*a, **kw)
        self.dp_rank, self.dp_size = dp_rank, dp_size

    def lines_for(self, seq_ids):
        local = seq_ids - self.dp_rank * self.num_lines
        return super().lines_for(torch.where((local < 0) | (local >= self.num_lines),
                                             torch.full_like(local, -1), local))


class BlockKVCacheManager(nn.Module):
    """Paged cache.  ``slot_mapping[b,t] = block_id*block_size + offset`` (-1 = skip)."""

    def __init__(self, num_layers: int, num_kv_heads: int, head_dim: int, num_blocks: int, block_size: int,
                 dtype=torch.bfloat16, device=None):
        super().__init__()
        self.num_layers, self.num_kv_heads, self.head_dim = num_layers, num_kv_heads, head_dim
        self.num_blocks, self.block_size = num_blocks, block_size
        self.register_buffer("cache", torch.zeros(num_layers, 2, num_blocks + 1, block_size, num_kv_heads, head_dim,
                                                  dtype=dtype, device=device), persistent=False)

    @property
    def past_key_values(self):
        out = []
        for i in range(self.num_layers):
            out += [self.cache[i, 0], self.cache[i, 1]]
        return out

    def get_kv_by_layer_id(self, idx):
        return self.cache[idx, 0], self.cache[idx, 1]

    def reset(self):
        self.cache.zero_()

    def update(self, layer, k_new, v_new, slot_mapping):
        ops.paged_kv_append(self.cache[layer, 0], self.cache[layer, 1], k_new, v_new, slot_mapping)

    def bytes(self):
        return self.cache.numel() * self.cache.element_size()


def generate_tokengen_slot_mapping(position_ids: torch.Tensor, slot_mapping: torch.Tensor,
                                   block_table: torch.Tensor, block_size: int) -> torch.Tensor:
    """Slot for the token at ``position_ids`` from the block table (on-device regeneration used by
    async decode; reference block_kv_cache_manager.py:376-431)."""
    blk = torch.gather(block_table.long(), 1, (position_ids.long() // block_size).clamp(0, block_table.shape[1] - 1))
    return (blk * block_size + position_ids.long() % block_size).to(slot_mapping.dtype)


def generate_fusedspec_slot_mapping(position_ids, slot_mapping, block_table, block_size, k: int):
    pos = position_ids.long() + torch.arange(k, device=position_ids.device).view(1, k)
    return generate_tokengen_slot_mapping(pos, slot_mapping, block_table, block_size)


def get_active_block_table(block_table: torch.Tensor, context_lens: torch.Tensor, block_size: int) -> torch.Tensor:
    """Compact the per-sequence block tables into one list of blocks in use, padded with 0
    (vLLM hook; reference kvcache/utils.py:153-206)."""
    n = (context_lens.long() + block_size - 1) // block_size
    rows = [block_table[i, : int(n[i])] for i in range(block_table.shape[0])]
    return torch.cat(rows) if rows else block_table.new_zeros(0)
