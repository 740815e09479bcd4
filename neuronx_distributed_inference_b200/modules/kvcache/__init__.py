"""KV-cache package — reference layout modules/kvcache/{kv_cache_manager,block_kv_cache_manager,data_parallel_kv_cache_manager,
gpt_oss_kv_cache_manager,multimodal_kv_cache_manager,utils}.py."""
from .gpt_oss_kv_cache_manager import GptOssKVCacheManager, HybridKVCacheManager  # noqa: F401
from .kv_cache_manager import (BlockKVCacheManager, DataParallelKVCacheManager, KVCacheManager,  # noqa: F401
                               generate_fusedspec_slot_mapping, generate_tokengen_slot_mapping, get_active_block_table)
from .multimodal_kv_cache_manager import MultimodalKVCacheManager, VisionKVStore  # noqa: F401
