"""Loads the reference's OWN Python for the hot path -- models/kv_cache.py (ShadowKVCache_CPU) and
models/tensor_op.py, copied verbatim into the git-ignored oracle/_ref/models/ by oracle/build_ref.sh -- on top of a
chosen native module:

* ``backend="b200"``  : ``kernels.shadowkv`` -> ``shadowkv_b200.shadowkv`` (the drop-in shim over the C ABI) and
  ``flash_attn_with_kvcache`` -> ``shadowkv_b200.attention``: INTEGRATION.md level 1, as an executed claim.
* ``backend="reference"`` : the unmodified reference extension built into oracle/_ref/kernels + the flash-attn wheel:
  the reference's own CUDA path, the checker.

Test infrastructure only (it executes oracle/_ref); the product never imports this.  ``minference`` (prefill-only,
absent from the image) is stubbed; ``flashinfer`` is the image's.  Both reference modules need a CUDA device at import
(models/tensor_op.py:57 builds a CUDA tensor at module level), so this is only usable from ``-m gpu`` tests."""
from __future__ import annotations

import importlib
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
REF_PY = os.path.join(REF_DIR, "models", "kv_cache.py")
REF_SO_DIR = os.path.join(REF_DIR, "kernels")


def available(backend: str) -> bool:
    if not os.path.exists(REF_PY):
        return False
    if backend == "reference":
        return any(f.startswith("shadowkv.") and f.endswith(".so") for f in os.listdir(REF_SO_DIR)) \
            if os.path.isdir(REF_SO_DIR) else False
    return True


def _native(backend: str):
    if backend == "b200":
        from shadowkv_b200 import shadowkv as native
        from shadowkv_b200.attention import flash_attn_with_kvcache
        return native, flash_attn_with_kvcache
    if REF_SO_DIR not in sys.path:
        sys.path.insert(0, REF_SO_DIR)
    import torch  # noqa: F401  (the extension links against libtorch)
    native = importlib.import_module("shadowkv")            # TORCH_EXTENSION_NAME=shadowkv (oracle/build_ref.sh)
    from flash_attn import flash_attn_with_kvcache          # what reference models/base.py:38 imports
    return native, flash_attn_with_kvcache


def load(backend: str):
    """Returns (kv_cache module, tensor_op module, flash_attn_with_kvcache) of the reference, bound to `backend`.
    Every call imports fresh module objects, so two backends can live side by side in one process."""
    native, attn = _native(backend)
    if "minference" not in sys.modules:
        stub = types.ModuleType("minference")
        for name in ("vertical_slash_sparse_attention", "block_sparse_attention", "streaming_forward"):
            setattr(stub, name, None)                       # prefill-only kernels (tensor_op.py:30), never called here
        sys.modules["minference"] = stub
    saved = {k: sys.modules.get(k) for k in ("kernels", "kernels.shadowkv", "models", "models.tensor_op",
                                              "models.kv_cache")}
    try:
        kernels = types.ModuleType("kernels")
        kernels.shadowkv = native
        kernels.__path__ = []
        sys.modules["kernels"] = kernels
        sys.modules["kernels.shadowkv"] = native
        pkg = types.ModuleType("models")
        pkg.__path__ = [os.path.join(REF_DIR, "models")]
        sys.modules["models"] = pkg
        sys.modules.pop("models.tensor_op", None)
        sys.modules.pop("models.kv_cache", None)
        tensor_op = importlib.import_module("models.tensor_op")
        kv_cache = importlib.import_module("models.kv_cache")
        assert kv_cache.shadowkv is native and tensor_op.shadowkv is native
        return kv_cache, tensor_op, attn
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def decode_layer_like_base_py(cache, attn, layer_idx, query_states, key_states, value_states, cos_sin_cache):
    """The decode branch of reference models/base.py:320-341, call for call (q / k arrive post-RoPE, which is what
    `apply_rotary_pos_emb` at :321 leaves)."""
    import torch
    cache.update_kv_cache(key_states, value_states, layer_idx)                                        # :322
    position_ids = cache.get_retrieval_position_ids(layer_idx=layer_idx, query_states=query_states)   # :324
    curr_stream = torch.cuda.current_stream()                                                          # :326
    get_value_stream = cache.copy_stream
    with torch.cuda.stream(get_value_stream):                                                          # :329-332
        get_value_stream.wait_stream(curr_stream)
        v = cache.get_value_cache(layer_idx, position_ids)
    k = cache.get_key_cache(layer_idx=layer_idx, position_ids=position_ids, rope_func=None,
                            cos_sin_cache=cos_sin_cache)                                               # :335
    curr_stream.wait_stream(get_value_stream)                                                          # :338
    return attn(q=query_states.transpose(1, 2), k_cache=k.transpose(1, 2), v_cache=v.transpose(1, 2),
                causal=True)                                                                           # :341
