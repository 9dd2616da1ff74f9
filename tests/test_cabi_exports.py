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
    # the single-launch V fetch + K rebuild and the attention variant with launch flags
    assert lib.skv_fetch_and_rebuild(C.byref(g), 1, 0, 1, 1, 1, 1, 1, 1, 1, 0, None, 0) == -2      # null offsets
    assert lib.skv_fetch_and_rebuild(C.byref(bad_s), 1, 0, 1, 1, 1, 1, 1, 1, 1, 1, None, 0) == -1  # select 200
    assert lib.skv_fetch_and_rebuild(C.byref(g), 1, 0, 1, 1, 1, 1, 1, 1, 1, 1, C.byref(vs), 0) == -2
    bad_r = _lib.Geom(batch=1, kv_heads=8, q_groups=4, head_dim=128, chunk=8, select=256, n_landmarks=15560, rank=164,
                      seq_len=124928, buf_len=2592, sparse_start=448)
    assert lib.skv_fetch_and_rebuild(C.byref(bad_r), 1, 0, 1, 1, 1, 1, 1, 1, 1, 1, None, 0) == -1  # rank % 8
    assert lib.skv_attention_decode_ex(1, 1, 1, 1, 1, 8, 4, 64, 100, 8, 8, 8, 0.1, 0, 1, 0, 1, 0) == -1


def test_dense_entry_points_reject_bad_arguments_without_a_gpu():
    """Part 4 of the header (dense remainder of the decode layer): argument checks run before any launch."""
    import ctypes as C
    from shadowkv_b200 import _lib
    lib = _lib.lib
    epi = _lib.DenseEpilogue(mode=1, out=8)
    one = 16                                               # any non-null "pointer": nothing is dereferenced
    assert lib.skv_dense_linear(0, one, 1, 4096, 4096, C.byref(epi), one, 0) == -2          # null x
    assert lib.skv_dense_linear(one, one, 65, 4096, 4096, C.byref(epi), one, 0) == -1       # batch > 64
    assert lib.skv_dense_linear(one, one, 1, 100, 4096, C.byref(epi), one, 0) == -1         # in_features % 64
    assert b"in_features" in lib.skv_last_error()
    bad = _lib.DenseEpilogue(mode=9, out=8)
    assert lib.skv_dense_linear(one, one, 1, 4096, 4096, C.byref(bad), one, 0) == -2        # unknown epilogue
    qkv = _lib.DenseEpilogue(mode=3, out=8, q_heads=32, kv_heads=8, k_out=8, v_out=8)
    assert lib.skv_dense_linear(one, one, 1, 4096, 4096, C.byref(qkv), one, 0) == -2        # out_features != 48 heads
    silu = _lib.DenseEpilogue(mode=2, out=8)
    assert lib.skv_dense_linear(one, one, 1, 4096, 200, C.byref(silu), one, 0) == -1        # 2F with F % 64 != 0
    assert lib.skv_rmsnorm(0, 0, 0, one, 1e-5, 0, one, 1, 4096, 0) == -2                    # neither x nor tok
    assert lib.skv_rmsnorm(one, 0, 0, one, 1e-5, 0, one, 1, 4100, 0) == -1                  # hidden % 8
    assert lib.skv_argmax(0, one, 1, 10, 0) == -2
    assert lib.skv_l2_prefetch(one + 1, 4096, 0) == -2                                      # not 16-byte aligned
    assert lib.skv_dense_finish_rmsnorm(one, 1, 4096, 4100, 0, one, one, 1e-5, one, 0) == -1
    assert lib.skv_dense_workspace_bytes() > 4096 * 4
    # plane buffer: 4096 x 4096 splits into 32 tiles x 64 slabs over 148 CTAs -> at most 6 CTAs touch a tile
    assert lib.skv_dense_partial_bytes(1, 4096, 4096) == 6 * 4096 * 4
    assert lib.skv_dense_partial_bytes(2, 64, 128) == 2 * 2 * 128 * 4


def test_header_is_plain_c_and_a_c_caller_links(tmp_path):
    """include/shadowkv_b200.h is the boundary a cgo / JNI / ctypes binding would consume: it has to compile as C99
    (no C++ in the signatures) and a C program has to link against the library and get the documented error codes."""
    import shutil
    import subprocess
    import pytest
    from shadowkv_b200 import _lib
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "caller.c"
    src.write_text(
        '#include "shadowkv_b200.h"\n#include <string.h>\n'
        "int main(void) {\n"
        '  if (!strstr(skv_version(), "sm_100a")) return 2;\n'
        "  if (skv_reorder_keys_and_compute_offsets(0, 0, 0, 0, 1, 1, 256, 0) != SKV_ERR_ARGUMENT) return 3;\n"
        '  if (!strstr(skv_last_error(), "null")) return 4;\n'
        "  if (skv_gather_copy((const skv_bf16 *)16, (skv_bf16 *)16, (const int64_t *)16, 1, 8, 100, 100, 200, 0)"
        " != SKV_ERR_UNSUPPORTED) return 5;\n"
        "  return 0;\n}\n")
    libdir = os.path.dirname(_lib.LIB_PATH)
    exe = tmp_path / "caller"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           str(src), "-o", str(exe), "-L", libdir, "-lshadowkv_b200", f"-Wl,-rpath,{libdir}"])
    assert subprocess.call([str(exe)]) == 0


def test_fused_attention_entry_point_checks_its_exchange_without_a_gpu():
    """skv_fetch_rebuild_attend validates the kv-head split's exchange before it touches CUDA: every peer needs a flag
    pointer and an epoch; an output pointer may only be NULL in the one-GPU form (world == 1: the row's flag word only)."""
    import ctypes as C
    from shadowkv_b200 import _lib
    lib = _lib.lib
    g = _lib.Geom(batch=1, kv_heads=8, q_groups=4, head_dim=128, chunk=8, select=256, n_landmarks=15560, rank=160,
                  seq_len=124928, buf_len=2592, sparse_start=448, table_mode=1)

    def call(ex, geom=g):
        return lib.skv_fetch_rebuild_attend(C.byref(geom), 1, 0, 1, 1, 1, 1, 1, 1, 1, 1, None, 1, 1, 2497, 0, 0.088, 1,
                                            C.byref(ex) if ex is not None else None, 0)

    two = _lib.Exchange(world=2, rank=0)
    two.peer_out[0], two.peer_flags[0], two.peer_flags[1], two.epoch_dev = 1, 1, 1, 1
    assert call(two) == -2 and b"null peer pointer" in lib.skv_last_error()        # peer 1 has no output buffer
    one = _lib.Exchange(world=1, rank=0)
    one.epoch_dev = 1
    assert call(one) == -2 and b"null peer pointer" in lib.skv_last_error()        # the one-GPU form still needs its flag word
    one.peer_flags[0] = 1
    one.epoch_dev = None
    assert call(one) == -2 and b"epoch_dev" in lib.skv_last_error()
    nine = _lib.Exchange(world=9, rank=0)
    nine.epoch_dev = 1
    assert call(nine) == -2
    ref_table = _lib.Geom(batch=1, kv_heads=8, q_groups=4, head_dim=128, chunk=8, select=256, n_landmarks=15560, rank=160,
                          seq_len=124928, buf_len=2592, sparse_start=448, table_mode=0)
    assert call(None, ref_table) == -1                                             # needs the in-place chunk table
