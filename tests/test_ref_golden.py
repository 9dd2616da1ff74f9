"""Fixtures captured from the UNMODIFIED reference extension on a B200 (tests/golden/ref_golden.npz, made by
tests/golden/make_ref_golden.py).  The CPU tests pin the oracle to them; the GPU tests pin the CUDA path."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import shadowkv_oracle as O
from shadowkv_b200 import synth
from tests.golden import inputs as I
from tests.golden.make_ref_golden import cases
from tests.tolerances import assert_rope_rows_close, assert_ulp_bf16

PATH = os.path.join(os.path.dirname(__file__), "golden", "ref_golden.npz")
needs_golden = pytest.mark.skipif(not os.path.exists(PATH), reason="ref_golden.npz not generated yet")


def _g():
    return np.load(PATH)


def _bf(a):
    return torch.from_numpy(a.copy()).view(torch.bfloat16)


# ------------------------------------------------------------------ oracle vs reference (CPU) ----
@needs_golden
def test_oracle_reorder_equals_reference():
    g, c = _g(), cases()["reorder"]
    nk, no, nc = O.reorder_keys_and_compute_offsets(c["cached"], c["cur"])
    np.testing.assert_array_equal(nk, g["reorder_new"])
    np.testing.assert_array_equal(no, g["reorder_off"])
    np.testing.assert_array_equal(nc, g["reorder_cnt"])


def _gather_inputs():
    B, H, nchunks, ss, S = 2, 4, 2048, 64, 256
    Lbuf = ss + S * 8 + 32
    host = I.bf16_tensor((B * H, nchunks, 1024), seed=3)
    kbuf = I.bf16_tensor((B * H, Lbuf, 128), seed=4)
    vbuf = I.bf16_tensor((B * H, Lbuf, 128), seed=5)
    return B, H, nchunks, ss, S, Lbuf, host, kbuf, vbuf


@needs_golden
def test_oracle_cached_gather_equals_reference():
    g = _g()
    B, H, nchunks, ss, S, Lbuf, host, kbuf, vbuf = _gather_inputs()
    nk, off, cnt = O.reorder_keys_and_compute_offsets(g["gather_cached"], g["gather_cur"])
    np.testing.assert_array_equal(nk, g["gather_new"])
    np.testing.assert_array_equal(off, g["gather_off"])
    np.testing.assert_array_equal(cnt, g["gather_cnt"])
    ks = kbuf[:, ss:ss + S * 8].clone(); O.gather_copy_d2d(ks, off, cnt); kbuf[:, ss:ss + S * 8] = ks
    vs = vbuf[:, ss:ss + S * 8].clone(); O.gather_copy_with_offsets(host, vs, off, cnt); vbuf[:, ss:ss + S * 8] = vs
    np.testing.assert_array_equal(I.bits(kbuf[:, ::8]), g["gather_k_rows"])
    np.testing.assert_array_equal(I.bits(vbuf[:, ::8]), g["gather_v_rows"])
    np.testing.assert_array_equal(kbuf.view(torch.int16).long().sum(dim=(1, 2)).numpy(), g["gather_k_sum"])
    np.testing.assert_array_equal(vbuf.view(torch.int16).long().sum(dim=(1, 2)).numpy(), g["gather_v_sum"])
    assert (g["gather_signals"] == 0).all()


def _rope_inputs():
    Br, Hr, Sr = 2, 2, 4
    x = I.bf16_tensor((Br, Hr, Sr * 8, 128), seed=6, exp_base=0x7D)
    ids = torch.from_numpy(I.distinct_ids(Br * Hr, Sr, 500, seed=7)).view(Br, Hr, Sr)
    rcnt = [0, 1, 3, 2]
    tok = (ids.unsqueeze(-1) * 8 + torch.arange(8)).view(Br, Hr, Sr * 8)
    return Br, Hr, Sr, x, ids, rcnt, tok


@needs_golden
def test_oracle_rope_equals_reference_bit_exact():
    """Pins the bf16 operation order (one rounded product folded into a bf16 FMA) seen in the reference SASS."""
    g = _g()
    Br, Hr, Sr, x, ids, rcnt, tok = _rope_inputs()
    cs = synth.cos_sin_table(4096 + 1024)
    np.testing.assert_array_equal(I.bits(O.rope_bf16(x, cs[tok], fused=True)), g["rope_new"])
    assert (I.bits(O.rope_bf16(x, cs[tok], fused=False)) != g["rope_new"]).any()     # the unfused order is NOT it
    for name, glm in (("neox", False), ("glm", True)):
        cs = synth.cos_sin_table(4096 + 1024, glm=glm)
        rot = O.rope_bf16(x, cs[tok], fused=True)
        want = torch.zeros(Br, Hr, 16 + Sr * 8, 128, dtype=torch.bfloat16)
        for r in range(Br * Hr):
            b, h = divmod(r, Hr)
            want[b, h, 16 + rcnt[r] * 8:] = rot[b, h, rcnt[r] * 8:]
        np.testing.assert_array_equal(I.bits(want), g[f"rope_push_{name}"])


def _score_inputs():
    bc, m, n = 4, 4, 2048
    A = I.bf16_tensor((bc, m, 128), seed=8, exp_base=0x7D)
    Bm = I.bf16_tensor((bc, n, 128), seed=9, exp_base=0x7D)
    return bc, m, n, A, Bm


@needs_golden
def test_oracle_scores_equal_reference_to_bf16_ulp():
    g = _g()
    bc, m, n, A, Bm = _score_inputs()
    d, s = O.landmark_scores(A.view(bc, m, 1, 128), Bm.view(bc, 1, n, 128))
    assert_ulp_bf16(d.view(bc, m, n), _bf(g["score_D"]), ulps=1, frac_exact=0.97, what="D vs reference")
    assert_ulp_bf16(s.view(bc, m, n), _bf(g["score_softmax"]), ulps=2, frac_exact=0.90, what="Softmax vs reference")


def _rebuild_inputs():
    Bg, Hg, Lg, R, Sg = 1, 2, 4096, 160, 128
    U = I.bf16_tensor((Bg, Lg, R), seed=10, exp_base=0x79)
    SV = I.bf16_tensor((Bg, Hg, 128, R), seed=11, exp_base=0x7D)
    gid = torch.from_numpy(I.distinct_ids(Bg * Hg, Sg, Lg // 8, seed=12)).view(Bg, Hg, Sg)
    return Bg, Hg, Lg, R, Sg, U, SV, gid, [0, 45]


@needs_golden
def test_oracle_rebuild_equals_reference_to_bf16_ulp():
    g = _g()
    Bg, Hg, Lg, R, Sg, U, SV, gid, gcnt = _rebuild_inputs()
    cs = synth.cos_sin_table(Lg + 1024)
    kpre, tok = O.rebuild_keys_pre_rope(U, SV, gid)
    ref_pre = _bf(g["rebuild_pre_rows"])
    ref_rope = _bf(g["rebuild_rope_rows"])
    rot = O.rope_bf16(kpre, cs[tok], fused=True)
    for h in range(Hg):
        first_tile = (gcnt[h] * 8 // 128) * 128
        assert_ulp_bf16(kpre[0, h, first_tile:][::4] if first_tile % 4 == 0 else None, ref_pre[0, h, first_tile // 4:],
                        ulps=1, frac_exact=0.97, what="pre-RoPE K vs reference")
        c = gcnt[h] * 8
        want = torch.zeros(32 + Sg * 8, 128, dtype=torch.bfloat16)
        want[32 + c:] = rot[0, h, c:]
        assert_rope_rows_close(want[::4], ref_rope[0, h], frac_exact=0.95, what="RoPE'd K vs reference")


# ------------------------------------------------------------------ CUDA path vs reference (GPU) ----
@needs_golden
@pytest.mark.gpu
def test_cuda_integer_path_equals_reference(cuda):
    from shadowkv_b200 import shadowkv
    g = _g()
    B, H, nchunks, ss, S, Lbuf, host, kbuf, vbuf = _gather_inputs()
    d_cached = torch.from_numpy(g["gather_cached"].copy()).to(cuda)
    d_cur = torch.from_numpy(g["gather_cur"].copy()).to(cuda)
    off = torch.zeros(B * H * S, dtype=torch.int32, device=cuda)
    cnt = torch.zeros(B * H, dtype=torch.int32, device=cuda)
    shadowkv.reorder_keys_and_compute_offsets(d_cached, d_cur, off, cnt, B, H, S)
    dk, dv = kbuf.to(cuda), vbuf.to(cuda)
    shadowkv.gather_copy_d2d_with_offsets(dk, off, cnt, B, H, S * 1024, ss * 128, Lbuf * 128, S)
    shadowkv.gather_copy_with_offsets(host.pin_memory(), dv, None, off, cnt, None, B, H, nchunks * 1024, S * 1024,
                                      ss * 128, Lbuf * 128, S)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(d_cached.cpu().numpy(), g["gather_new"])
    np.testing.assert_array_equal(off.cpu().numpy().reshape(B * H, S), g["gather_off"])
    np.testing.assert_array_equal(cnt.cpu().numpy(), g["gather_cnt"])
    np.testing.assert_array_equal(I.bits(dk[:, ::8]), g["gather_k_rows"])
    np.testing.assert_array_equal(I.bits(dv[:, ::8]), g["gather_v_rows"])
    np.testing.assert_array_equal(dk.view(torch.int16).long().sum(dim=(1, 2)).cpu().numpy(), g["gather_k_sum"])
    np.testing.assert_array_equal(dv.view(torch.int16).long().sum(dim=(1, 2)).cpu().numpy(), g["gather_v_sum"])


@needs_golden
@pytest.mark.gpu
def test_cuda_rope_and_scores_and_rebuild_equal_reference(cuda):
    from shadowkv_b200 import shadowkv
    from shadowkv_b200.tensor_op import (apply_rotary_pos_emb_cuda, apply_rotary_pos_emb_cuda_push_cache,
                                         batch_gather_gemm_rotary_pos_emb_cuda)
    g = _g()
    Br, Hr, Sr, x, ids, rcnt, tok = _rope_inputs()
    cs = synth.cos_sin_table(4096 + 1024).to(cuda)
    np.testing.assert_array_equal(I.bits(apply_rotary_pos_emb_cuda(x.to(cuda), cs, tok.to(cuda))), g["rope_new"])
    for name, glm in (("neox", False), ("glm", True)):
        cs = synth.cos_sin_table(4096 + 1024, glm=glm).to(cuda)
        cache = torch.zeros(Br, Hr, 16 + Sr * 8, 128, dtype=torch.bfloat16, device=cuda)
        apply_rotary_pos_emb_cuda_push_cache(x.to(cuda), cs, ids.to(torch.int32).to(cuda), 8, cache, 16, 16 + Sr * 8,
                                             torch.tensor(rcnt, dtype=torch.int32, device=cuda))
        np.testing.assert_array_equal(I.bits(cache), g[f"rope_push_{name}"])
    bc, m, n, A, Bm = _score_inputs()
    D = torch.zeros(bc, m, n, dtype=torch.bfloat16, device=cuda)
    Sm = torch.zeros_like(D)
    T = (n + 255) // 256
    shadowkv.batch_gemm_softmax(A.to(cuda), Bm.to(cuda), D, torch.zeros(bc, m, T, device=cuda),
                                torch.zeros(bc, m, T, device=cuda), Sm, bc, m, n, 128, 1 / math.sqrt(128), 0)
    assert_ulp_bf16(D, _bf(g["score_D"]), ulps=1, frac_exact=0.97, what="D vs reference")
    assert_ulp_bf16(Sm, _bf(g["score_softmax"]), ulps=2, frac_exact=0.90, what="Softmax vs reference")
    Bg, Hg, Lg, R, Sg, U, SV, gid, gcnt = _rebuild_inputs()
    cs = synth.cos_sin_table(Lg + 1024).to(cuda)
    scratch = torch.zeros(Bg, Hg, Sg * 8, 128, dtype=torch.bfloat16, device=cuda)
    kc = torch.zeros(Bg, Hg, 32 + Sg * 8, 128, dtype=torch.bfloat16, device=cuda)
    batch_gather_gemm_rotary_pos_emb_cuda(U.to(cuda), SV.to(cuda), cs, gid.to(cuda), scratch, 8, kc, 32, 32 + Sg * 8,
                                          torch.tensor(gcnt, dtype=torch.int32, device=cuda))
    torch.cuda.synchronize()
    assert_ulp_bf16(scratch[:, :, ::4], _bf(g["rebuild_pre_rows"]), ulps=1, frac_exact=0.97, what="pre-RoPE K vs reference")
    assert_rope_rows_close(kc[:, :, ::4], _bf(g["rebuild_rope_rows"]), frac_exact=0.95, what="RoPE'd K vs reference")


@needs_golden
@pytest.mark.gpu
def test_cuda_product_scoring_and_selection_kernels_equal_reference(cuda):
    """The PRODUCT kernels (score_tcgen05_kernel: TMA ring + tcgen05; select_kernel), not the SIMT drop-in entry point,
    against the reference's batch_gemm_softmax outputs: the scaled bf16 dot products D within 1 bf16 ulp, and the
    max-over-heads bf16 softmax the selection ranks within 2 ulp of the max over the reference's Softmax rows."""
    import ctypes as C
    from shadowkv_b200 import _lib
    g = _g()
    bc, m, n, A, Bm = _score_inputs()
    S = 256
    geom = _lib.Geom(batch=bc, kv_heads=1, q_groups=m, head_dim=128, chunk=8, select=S, n_landmarks=n, rank=160,
                     seq_len=n * 8, buf_len=S * 8 + 600, sparse_start=448, table_mode=1)
    ws = torch.zeros(_lib.lib.skv_select_workspace_bytes(C.byref(geom)), dtype=torch.uint8, device=cuda)
    d_q, d_lm = A.to(cuda).contiguous(), Bm.to(cuda).contiguous()
    idx = torch.arange(n, dtype=torch.int64, device=cuda).expand(bc, 1, n).contiguous()
    res = torch.full((bc, 1, S), -1, dtype=torch.int64, device=cuda)
    off = torch.zeros(bc * S, dtype=torch.int32, device=cuda)
    cnt = torch.zeros(bc, dtype=torch.int32, device=cuda)
    sel = torch.zeros(bc, 1, S, dtype=torch.int64, device=cuda)
    score = torch.zeros(bc, 1, n, dtype=torch.bfloat16, device=cuda)
    _lib.check(_lib.lib.skv_select_chunks(C.byref(geom), d_q.data_ptr(), d_lm.data_ptr(), idx.data_ptr(), res.data_ptr(),
                                          off.data_ptr(), cnt.data_ptr(), sel.data_ptr(), score.data_ptr(), ws.data_ptr(),
                                          None, torch.cuda.current_stream().cuda_stream), "select_chunks")
    torch.cuda.synchronize()
    ldn = (n + 7) // 8 * 8
    D = ws[:bc * m * ldn * 2].view(torch.bfloat16).view(bc, m, ldn)[:, :, :n]      # the workspace starts with the scores
    assert_ulp_bf16(D, _bf(g["score_D"]), ulps=1, frac_exact=0.97, what="product scoring kernel D vs reference")
    want_score = _bf(g["score_softmax"]).float().max(dim=1).values.to(torch.bfloat16)
    assert_ulp_bf16(score.view(bc, n), want_score, ulps=2, frac_exact=0.90, what="selection key vs max_g reference Softmax")
    # and the selection itself: top-S of the REFERENCE's softmax, modulo chunks within 2 ulp of its S-th value
    key = want_score.view(torch.int16).int()
    for b in range(bc):
        thr = int(key[b].sort(descending=True).values[S - 1])
        mine = set(sel[b, 0].tolist())
        theirs = set(torch.sort(want_score[b].float(), descending=True, stable=True).indices[:S].tolist())
        assert len(mine ^ theirs) // 2 <= 3
        for c in mine ^ theirs:
            assert abs(int(key[b, c]) - thr) <= 2
