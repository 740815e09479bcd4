"""Cache write helpers under the reference's names (modules/kvcache/utils.py: ``write_kv_cache_at_batch_kernel`` is the NKI
indirect-DMA writer; here it is the ``kv_append`` CUDA kernel with the same out-of-bounds-skip semantics)."""
from ... import ops
from .kv_cache_manager import get_active_block_table  # noqa: F401


def write_kv_cache_at_batch(k_cache, v_cache, k_new, v_new, seq_ids, positions):
    """k_new/v_new [B,T,H,D] -> cache[L,H,S,D] at (seq_ids[b], positions[b,t]); negative line / position = skip."""
    return ops.kv_append(k_cache, v_cache, k_new, v_new, seq_ids, positions)


dynamic_update_slice = write_kv_cache_at_batch
