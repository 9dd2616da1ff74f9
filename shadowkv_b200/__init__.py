"""B200-native ShadowKV sparse-KV decode step.

Host-side mirror of the reference's interface for this path:

* ``shadowkv_b200.shadowkv``   -- the 12 extension entry points (reference kernels/main.cu:45-80)
* ``shadowkv_b200.tensor_op``  -- the three wrappers of reference models/tensor_op.py:154-238
* ``shadowkv_b200.kv_cache``   -- ``ShadowKVCache_CPU`` / ``ShadowKVCache`` (reference models/kv_cache.py)
* ``shadowkv_b200.attention``  -- ``flash_attn_with_kvcache`` replacement (reference models/base.py:341)

All compute goes through ``lib/libshadowkv_b200.so`` (hand-written sm_100a CUDA behind a C ABI,
``include/shadowkv_b200.h``).  There is no CPU or PyTorch fallback: importing ``_lib`` raises if the
library has not been built.
"""
__version__ = "0.1.0"
