from .kv_cache import HFCacheAdapter, MultiLayerPagedKVCache4Bit  # noqa: F401
from .modeling_llama import (FlatQuantLlamaAttention, FlatQuantLlamaConfig, FlatQuantLlamaDecoderLayer,  # noqa: F401
                             FlatQuantLlamaForCausalLM, FlatQuantLlamaMLP, convert_checkpoint_keys)
from .modeling_deepseek import FlatQuantDeepSeekBlock  # noqa: F401
