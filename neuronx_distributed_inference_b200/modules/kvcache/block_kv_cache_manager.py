from .kv_cache_manager import (BlockKVCacheManager, generate_fusedspec_slot_mapping, generate_tokengen_slot_mapping,  # noqa: F401
                               get_active_block_table)
