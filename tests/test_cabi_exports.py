"""The C-ABI library loads on a CPU-only host and exports every symbol include/shadowkv_b200.h declares.
No compute calls here (no GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "shadowkv_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(skv_[a-z0-9_]+)\s*\(", text)))


def test_header_lists_the_reference_entry_points():
    names = _declared()
    ref = ["gather_copy", "gather_copy_d2d_with_offsets", "reorder_keys_and_compute_offsets",
           "gather_copy_with_offsets", "apply_rotary_pos_emb", "apply_rotary_pos_emb_new",
           "apply_rotary_pos_emb_new_v2", "apply_rotary_pos_emb_push_cache", "apply_rotary_pos_emb_push_cache_opt",
           "apply_rotary_pos_emb_push_cache_opt_glm", "batch_gather_gemm", "batch_gemm_softmax"]   # main.cu:45-80
    for r in ref:
        assert "skv_" + r in names


def test_library_exports_every_declared_symbol():
    from shadowkv_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(_lib.EXPORTS) == declared
    assert "sm_100a" in _lib.version()


def test_shim_has_the_reference_names():
    from shadowkv_b200 import shadowkv
    for n in shadowkv.__all__:
        assert callable(getattr(shadowkv, n))
    assert len(shadowkv.__all__) == 12


def test_argument_errors_do_not_need_a_gpu():
    # null pointers / unsupported sizes are rejected before any launch
    from shadowkv_b200 import _lib
    rc = _lib.lib.skv_reorder_keys_and_compute_offsets(0, 0, 0, 0, 1, 1, 256, 0)
    assert rc == -2 and b"null" in _lib.lib.skv_last_error()
    import pytest
    with pytest.raises(ValueError):
        _lib.check(rc, "reorder")


def test_fused_entry_points_reject_bad_arguments_without_a_gpu():
    import ctypes as C
    import pytest
    from shadowkv_b200 import _lib
    g = _lib.Geom(batch=1, kv_heads=8, q_groups=4, head_dim=128, chunk=8, select=256, n_landmarks=15560, rank=160,
                  seq_len=124928, buf_len=2592, sparse_start=448)
    lib = _lib.lib
    assert lib.skv_select_chunks(C.byref(g), 0, 0, 0, 0, 0, 0, 0, 0, 0, None, 0) == -2
    bad = _lib.Geom(batch=1, kv_heads=8, q_groups=4, head_dim=64, chunk=8, select=256, n_landmarks=15560, rank=160,
                    seq_len=124928, buf_len=2592, sparse_start=448)
    assert lib.skv_score_landmarks(C.byref(bad), 1, 1, 1, None, 0) == -1          # head_dim 64 unsupported
    with pytest.raises(ValueError):
        _lib.check(-1, "score")
    bad_s = _lib.Geom(batch=1, kv_heads=8, q_groups=4, head_dim=128, chunk=8, select=200, n_landmarks=15560, rank=160,
                      seq_len=124928, buf_len=2592, sparse_start=448)
    assert lib.skv_compact_hits(C.byref(bad_s), 1, 1, 1, 1, 0) == -1                # select must be 128/256/512/1024
    assert lib.skv_gather_copy(1, 1, 1, 1, 8, 100, 100, 200, 0) == -1               # map_size 200 (reference silently no-ops)
    assert lib.skv_rebuild_keys(C.byref(g), 0, 0, 0, 0, 0, 0, 0, 0, 0) == -2
    assert lib.skv_attention_decode(1, 1, 1, 1, 1, 8, 4, 64, 100, 8, 8, 8, 0.1, 0, 1, 0, 0) == -1   # head_dim 64
    assert lib.skv_batch_gemm_softmax(1, 1, 1, 1, 1, 1, 8, 4, 100, 64, 0.1, 0.0, 0) == -1          # k must be 128
    vs = _lib.VStore(tags=1, data=1, pcie_count=0, ways=100)
    assert lib.skv_fetch_misses(C.byref(g), 1, 0, 1, 1, 1, 1, C.byref(vs), 0) == -2  # ways not a power of two
    assert lib.skv_select_workspace_bytes(C.byref(g)) > 8 * 4 * 15560 * 2
