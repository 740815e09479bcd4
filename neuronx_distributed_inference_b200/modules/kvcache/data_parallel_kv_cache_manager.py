from .kv_cache_manager import DataParallelKVCacheManager  # noqa: F401
