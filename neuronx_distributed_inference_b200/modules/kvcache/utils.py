"""Cache helpers under the reference's names (modules/kvcache/utils.py).  There they are XLA ``DynamicUpdateSlice`` custom calls, an
NKI indirect-DMA writer and the index arithmetic that stitches "cached prefix + new chunk" together for chunked prefill; here the
cache is a plain device buffer updated in place, so they are thin, in-place torch / CUDA-kernel equivalents."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch

from ... import ops
from .kv_cache_manager import get_active_block_table  # noqa: F401


def write_kv_cache_at_batch(k_cache, v_cache, k_new, v_new, seq_ids, positions):
    """k_new/v_new [B,T,H,D] -> cache[L,H,S,D] at (seq_ids[b], positions[b,t]); negative line / position = skip
    (reference ``write_kv_cache_at_batch_kernel`` :17-67, OOB-skip semantics)."""
    return ops.kv_append(k_cache, v_cache, k_new, v_new, seq_ids, positions)


def fill_prefix(cache: torch.Tensor, prefix_cache: torch.Tensor) -> torch.Tensor:
    """Write ``prefix_cache`` into the leading corner of ``cache`` (every dimension starts at 0) — reference :70-85."""
    cache[tuple(slice(0, s) for s in prefix_cache.shape)] = prefix_cache.to(cache.dtype)
    return cache


def dynamic_update_slice(tensor: torch.Tensor, update: torch.Tensor, start_indices: Sequence) -> torch.Tensor:
    """XLA DynamicUpdateSlice semantics, in place: ``tensor[s0:s0+u0, s1:s1+u1, ...] = update`` with the start indices clamped so
    the update fits (reference :87-133)."""
    assert len(start_indices) == tensor.dim(), "one start index per dimension"
    idx = []
    for s, u, n in zip(start_indices, update.shape, tensor.shape):
        s = max(0, min(int(s), n - u))
        idx.append(slice(s, s + u))
    tensor[tuple(idx)] = update.to(tensor.dtype)
    return tensor


def update_cache_const_indices(cache: torch.Tensor, updates: torch.Tensor, sequence_ids: torch.Tensor) -> torch.Tensor:
    """Prefill write of a whole bucket: ``cache[sequence_ids[b], :, :T] = updates[b]`` for ``updates [B, H, T, D]``
    (reference :136-150; out-of-range lines are skipped)."""
    B, H, T, D = updates.shape
    lines = sequence_ids.long().view(-1)
    ok = (lines >= 0) & (lines < cache.shape[0])
    cache[lines[ok], :, :T] = updates[ok].to(cache.dtype)
    return cache


def contexted_kv_indexing(q_lens: torch.Tensor, k_lens: torch.Tensor, block_table: torch.Tensor, block_size: int
                          ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Index plan for chunked prefill: sequence ``b`` contributes ``k_lens[b] - q_lens[b]`` tokens that already sit in the paged cache
    followed by its ``q_lens[b]`` new tokens; sequences are packed back to back (reference :312-505, dynamic flavour).
    -> (cache_slots [n_cached] flat slot ids into the paged cache, dst_cached [n_cached], dst_new [n_new]) where ``dst_*`` are positions
    in the packed ``[sum(k_lens)]`` axis; new tokens are taken in their packed ``[sum(q_lens)]`` order."""
    cache_slots, dst_c, dst_n = [], [], []
    off = 0
    for b in range(q_lens.shape[0]):
        q, k = int(q_lens[b]), int(k_lens[b])
        c = k - q
        pos = torch.arange(c, dtype=torch.long)
        blocks = block_table[b, pos // block_size].long()
        cache_slots.append(blocks * block_size + pos % block_size)
        dst_c.append(off + pos)
        dst_n.append(off + c + torch.arange(q, dtype=torch.long))
        off += k
    return torch.cat(cache_slots), torch.cat(dst_c), torch.cat(dst_n)


def contexted_kv(cache: torch.Tensor, current: torch.Tensor, cache_slots: torch.Tensor, dst_cached: torch.Tensor,
                 dst_new: torch.Tensor) -> torch.Tensor:
    """Combine the paged cache ``[num_blocks, block_size, H, D]`` and the new tokens ``current [n_new, H, D]`` into one packed
    ``[sum(k_lens), H, D]`` K (or V) following :func:`contexted_kv_indexing` (reference :209-255)."""
    nb, bs, H, D = cache.shape
    total = int(dst_cached.numel() + dst_new.numel())
    out = current.new_zeros(total, H, D)
    dev = current.device
    out[dst_cached.to(dev)] = cache.view(nb * bs, H, D)[cache_slots.to(dev)].to(current.dtype)
    out[dst_new.to(dev)] = current
    return out


def get_layer_to_kv_cache_size_mapping_for_mixed_attn(local_cache_size: int, global_cache_size: int, is_layer_locals: List[bool]) -> List[int]:
    """Per-layer cache length for models that mix sliding-window (local) and full (global) attention (reference :507-517)."""
    if local_cache_size is None or global_cache_size is None:
        raise ValueError("both cache sizes are required")
    return [local_cache_size if loc else global_cache_size for loc in is_layer_locals]


def get_kv_shapes(max_len: int, bsz: int, num_kv_heads_per_rank: int, head_dim: int, k_cache_transposed: bool = False,
                  is_kv_cache_tiled: bool = False) -> Tuple[Tuple[int, ...], Tuple[int, ...]]:
    """(K shape, V shape) of one layer's cache (reference :519-).  The 128-tiled and transposed-K layouts are Trainium DMA
    optimisations; they are reported for parity, the B200 kernels read the plain ``[B, H, S, D]`` layout."""
    v = (bsz, num_kv_heads_per_rank, max_len, head_dim)
    if is_kv_cache_tiled:
        v = (bsz, num_kv_heads_per_rank, max_len // 128, 128, head_dim)
    k = (bsz, num_kv_heads_per_rank, head_dim, max_len) if k_cache_transposed else v
    return k, v
