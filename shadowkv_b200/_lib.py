"""ctypes binding of include/shadowkv_b200.h.  Fails loudly when the library is missing."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SKV_LIB selects another build of the same library (the phase-timeline build of `make trace`, for tuning)
LIB_PATH = os.environ.get("SKV_LIB") or os.path.join(_HERE, "lib", "libshadowkv_b200.so")


class SkvError(RuntimeError):
    pass


class Geom(C.Structure):
    """struct skv_geom (include/shadowkv_b200.h)."""
    _fields_ = [(n, C.c_int32) for n in (
        "batch", "kv_heads", "q_groups", "head_dim", "chunk", "select", "n_landmarks", "rank", "seq_len",
        "buf_len", "sparse_start", "rope_glm", "cos_sin_rows", "table_mode", "flags")] + [("row_flags", C.c_void_p), ("wait_flags", C.c_void_p),
                                                                           ("wait_epoch_dev", C.c_void_p), ("wait_count", C.c_int32),
                                                                           ("reserved", C.c_int32), ("select_hint", C.c_void_p)]


GEOM_DYNAMIC_LANDMARKS = 1   # SKV_GEOM_DYNAMIC_LANDMARKS
GEOM_WAIT_FLAGS_ONLY = 2     # SKV_GEOM_WAIT_FLAGS_ONLY


class Append(C.Structure):
    """struct skv_append (include/shadowkv_b200.h)."""
    _fields_ = [("k_new", C.c_void_p), ("v_new", C.c_void_p), ("k_buf", C.c_void_p), ("v_buf", C.c_void_p),
                ("pos_dev", C.c_void_p), ("pos", C.c_int32), ("reserved", C.c_int32)]


class VStore(C.Structure):
    """struct skv_vstore (include/shadowkv_b200.h)."""
    _fields_ = [("tags", C.c_void_p), ("data", C.c_void_p), ("pcie_count", C.c_void_p), ("ways", C.c_int32),
                ("reserved", C.c_int32)]


MAX_PEERS = 8


class Exchange(C.Structure):
    """struct skv_exchange (include/shadowkv_b200.h)."""
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("peer_out", C.c_void_p * MAX_PEERS),
                ("peer_flags", C.c_void_p * MAX_PEERS), ("rows_done", C.c_void_p), ("epoch_dev", C.c_void_p)]


class DenseEpilogue(C.Structure):
    """struct skv_dense_epilogue (include/shadowkv_b200.h)."""
    _fields_ = [("mode", C.c_int32), ("q_heads", C.c_int32), ("kv_heads", C.c_int32), ("rope_mode", C.c_int32),
                ("cos_sin_rows", C.c_int32), ("pos", C.c_int32), ("static_weights", C.c_int32), ("reserved", C.c_int32),
                ("pos_dev", C.c_void_p), ("bias", C.c_void_p),
                ("resid", C.c_void_p), ("cos_sin", C.c_void_p), ("out", C.c_void_p), ("k_out", C.c_void_p),
                ("v_out", C.c_void_p), ("head0", C.c_int32), ("head_count", C.c_int32), ("kv_row", C.c_int32),
                ("kv_rows", C.c_int32), ("kv_stride_b", C.c_int64), ("kv_stride_h", C.c_int64),
                ("prefetch_next", C.c_void_p), ("prefetch_bytes", C.c_int64)]


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `make` (or `python -c 'import __graft_entry__ as g; g.build()'`)."
        " shadowkv_b200 has no CPU/PyTorch fallback.")

lib = C.CDLL(LIB_PATH)

_p, _i, _l, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float

_SIGS = {
    "skv_gather_copy": [_p, _p, _p] + [_i] * 5 + [_p],
    "skv_gather_copy_d2d_with_offsets": [_p, _p, _p] + [_i] * 6 + [_p],
    "skv_reorder_keys_and_compute_offsets": [_p, _p, _p, _p] + [_i] * 3 + [_p],
    "skv_gather_copy_with_offsets": [_p] * 6 + [_i] * 7 + [_p],
    "skv_apply_rotary_pos_emb": [_p] * 5 + [_i] * 14 + [_p],
    "skv_apply_rotary_pos_emb_new": [_p] * 4 + [_i] * 13 + [_p],
    "skv_apply_rotary_pos_emb_new_v2": [_p] * 4 + [_i] * 14 + [_p],
    "skv_apply_rotary_pos_emb_push_cache": [_p] * 5 + [_i] * 19 + [_p],
    "skv_apply_rotary_pos_emb_push_cache_opt": [_p] * 5 + [_i] * 19 + [_p],
    "skv_apply_rotary_pos_emb_push_cache_opt_glm": [_p] * 5 + [_i] * 19 + [_p],
    "skv_batch_gather_gemm": [_p] * 6 + [_i] * 8 + [_p, _p],
    "skv_batch_gemm_softmax": [_p] * 6 + [_i] * 4 + [_f, _f, _p],
    "skv_select_chunks": [C.POINTER(Geom)] + [_p] * 9 + [C.POINTER(Append), _p],
    "skv_score_landmarks": [C.POINTER(Geom)] + [_p] * 3 + [C.POINTER(Append), _p],
    "skv_select_from_scores": [C.POINTER(Geom)] + [_p] * 8,
    "skv_compact_hits": [C.POINTER(Geom)] + [_p] * 5,
    "skv_fetch_misses": [C.POINTER(Geom), _p, _l, _p, _p, _p, _p, C.POINTER(VStore), _p],
    "skv_vstore_fill": [C.POINTER(Geom), _p, _p, _p, _p, C.POINTER(VStore), _p],
    "skv_rebuild_keys": [C.POINTER(Geom)] + [_p] * 7 + [_i, _p],
    "skv_chunk_stats": [_p, _i, _i, _i, _l, _l, _l, _p, _p, _p],
    "skv_select_outliers": [_p, _i, _i, _i, _p, _p, _p],
    "skv_gather_rows": [_p, _p, _p, _i, _i, _i, _p],
    "skv_gather_chunks": [_p, _l, _p, _l, _l, _p, _i, _i, _p],
    "skv_append_kv": [C.POINTER(Geom)] + [_p] * 4 + [_i, _p, _p],
    "skv_attention_decode": [_p] * 4 + [_i] * 5 + [_l] * 3 + [_f, _i, _p, _p, _p],
    "skv_attention_decode_ex": [_p] * 4 + [_i] * 5 + [_l] * 3 + [_f, _i, _p, _p, _i, _p],
    "skv_fetch_and_rebuild": [C.POINTER(Geom), _p, _l, _p, _p, _p, _p, _p, _p, _p, _p, C.POINTER(VStore), _p],
    "skv_fetch_rebuild_attend": [C.POINTER(Geom), _p, _l, _p, _p, _p, _p, _p, _p, _p, _p, C.POINTER(VStore), _p, _p, _i, _p,
                                 _f, _p, C.POINTER(Exchange), _p],
    "skv_wait_flags": [_p, _i, _p, _p],
    "skv_dense_linear": [_p, _p, _i, _i, _i, C.POINTER(DenseEpilogue), _p, _p],
    "skv_rmsnorm": [_p, _p, _p, _p, _f, _p, _p, _i, _i, _p],
    "skv_argmax": [_p, _p, _i, _i, _p],
    "skv_dense_finish_rmsnorm": [_p, _i, _i, _i, _p, _p, _p, _f, _p, _p],
    "skv_l2_prefetch": [_p, _l, _p],
}
# every symbol include/shadowkv_b200.h declares (tests check the library exports all of them)
EXPORTS = sorted(list(_SIGS) + ["skv_version", "skv_last_error", "skv_active_config", "skv_select_workspace_bytes",
                                "skv_attention_workspace_bytes", "skv_attend_workspace_bytes", "skv_dense_workspace_bytes",
                                "skv_dense_partial_bytes"])

for _name, _args in _SIGS.items():
    _fn = getattr(lib, _name)
    _fn.argtypes = _args
    _fn.restype = C.c_int
lib.skv_version.restype = C.c_char_p
lib.skv_last_error.restype = C.c_char_p
lib.skv_active_config.restype = C.c_char_p
lib.skv_select_workspace_bytes.argtypes = [C.POINTER(Geom)]
lib.skv_select_workspace_bytes.restype = C.c_int64
lib.skv_attention_workspace_bytes.argtypes = [_i] * 5
lib.skv_attention_workspace_bytes.restype = C.c_int64
lib.skv_attend_workspace_bytes.argtypes = [_i] * 3
lib.skv_attend_workspace_bytes.restype = C.c_int64
lib.skv_dense_workspace_bytes.argtypes = []
lib.skv_dense_workspace_bytes.restype = C.c_int64
lib.skv_dense_partial_bytes.argtypes = [_i, _i, _i]
lib.skv_dense_partial_bytes.restype = C.c_int64


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib.skv_last_error().decode()
        if rc == -1:
            raise ValueError(f"{what}: unsupported argument: {msg}")
        if rc == -2:
            raise ValueError(f"{what}: bad argument: {msg}")
        raise SkvError(f"{what}: rc={rc}: {msg}")


def version() -> str:
    return lib.skv_version().decode()


def active_config() -> str:
    """Implementations active in this process (product kernels or an SKV_* selected cross-check / A-B variant)."""
    return lib.skv_active_config().decode()
