from .kv_cache import MultiLayerPagedKVCache4Bit  # noqa: F401
