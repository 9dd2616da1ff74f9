"""K rebuild (tcgen05 path and its SIMT cross-check) against the oracle; drop-in batch_gather_gemm too."""
import ctypes as C

import pytest
import torch

from oracle import shadowkv_oracle as O
from shadowkv_b200 import synth
from tests.tolerances import assert_close_bf16, assert_rope_rows_close, assert_ulp_bf16, cosine

pytestmark = pytest.mark.gpu


def _setup(B, H, L, R, S, seed, glm):
    g = torch.Generator().manual_seed(seed)
    U = (torch.randn(B, L, R, generator=g) / R ** 0.5).to(torch.bfloat16)
    SV = torch.randn(B, H, 128, R, generator=g).to(torch.bfloat16)
    ids = torch.stack([torch.randperm(L // 8, generator=g)[:S] for _ in range(B * H)]).view(B, H, S)
    cs = synth.cos_sin_table(L + 1024, glm=glm)
    return U, SV, ids, cs


@pytest.mark.parametrize("impl", [0, 1])
@pytest.mark.parametrize("glm,R,S", [(False, 160, 256), (True, 160, 128), (False, 64, 128), (False, 256, 128),
                                     (False, 32, 128), (False, 160, 1024)])
def test_rebuild_keys(cuda, impl, glm, R, S):
    from shadowkv_b200 import _lib
    B, H, L = 2, 2, 8192
    U, SV, ids, cs = _setup(B, H, L, R, S, seed=R + S, glm=glm)
    cnts = torch.tensor([0, S, 37 % S, S - 1][:B * H], dtype=torch.int32)
    ss, Lbuf = 448, 448 + S * 8 + 96
    g = torch.Generator().manual_seed(9)
    kbuf = torch.randn(B, H, Lbuf, 128, generator=g).to(torch.bfloat16)
    want = kbuf.clone()
    rot = O.rebuild_keys(U, SV, ids, cs)
    for r in range(B * H):
        b, h = divmod(r, H)
        c = int(cnts[r]) * 8
        want[b, h, ss + c:ss + S * 8] = rot[b, h, c:]
    geom = _lib.Geom(batch=B, kv_heads=H, q_groups=4, head_dim=128, chunk=8, select=S, n_landmarks=L // 8, rank=R,
                     seq_len=L, buf_len=Lbuf, sparse_start=ss, rope_glm=int(glm))
    d = [t.to(cuda) for t in (U, SV, cs, ids, cnts, kbuf)]
    _lib.check(_lib.lib.skv_rebuild_keys(C.byref(geom), *[t.data_ptr() for t in d[:5]], 0, d[5].data_ptr(), impl,
                                         torch.cuda.current_stream().cuda_stream), "rebuild")
    torch.cuda.synchronize()
    got = d[5].cpu()
    # rows outside the missed slots are untouched, bit for bit
    mask = torch.zeros(B, H, Lbuf, dtype=torch.bool)
    for r in range(B * H):
        b, h = divmod(r, H)
        mask[b, h, ss + int(cnts[r]) * 8:ss + S * 8] = True
    assert torch.equal(got[~mask], kbuf[~mask])
    # the bound: <= 1 bf16 ulp of the pre-RoPE value, rotated (tests/tolerances.py); no fixed absolute cap, see there.
    # north_star's cosine bar is kept on top of it
    assert_rope_rows_close(got[mask], want[mask], glm=glm, frac_exact=0.90, what="rebuilt K")
    assert cosine(got[mask], want[mask]) > 0.9999


def test_tcgen05_equals_simt(cuda):
    from shadowkv_b200 import _lib
    B, H, L, R, S = 1, 8, 124928, 160, 256           # config-2 shapes
    U, SV, ids, cs = _setup(B, H, L, R, S, seed=5, glm=False)
    cnts = torch.randint(0, S, (B * H,), dtype=torch.int32)
    geom = _lib.Geom(batch=B, kv_heads=H, q_groups=4, head_dim=128, chunk=8, select=S, n_landmarks=15560, rank=R,
                     seq_len=L, buf_len=2592, sparse_start=448)
    outs = []
    for impl in (0, 1):
        kbuf = torch.zeros(B, H, 2592, 128, dtype=torch.bfloat16, device=cuda)
        d = [t.to(cuda) for t in (U, SV, cs, ids, cnts)]
        _lib.check(_lib.lib.skv_rebuild_keys(C.byref(geom), *[t.data_ptr() for t in d], 0, kbuf.data_ptr(), impl,
                                             torch.cuda.current_stream().cuda_stream), "rebuild")
        torch.cuda.synchronize()
        outs.append(kbuf.cpu())
    assert_rope_rows_close(outs[0], outs[1], frac_exact=0.95, what="tcgen05 vs SIMT")
    assert float(outs[0].float().abs().max()) > 1.0


def test_batch_gather_gemm_wrapper(cuda):
    """Drop-in path of kv_cache.py:1157-1168: batch_gather_gemm into a bf16 scratch, then RoPE + push."""
    from shadowkv_b200.tensor_op import batch_gather_gemm_rotary_pos_emb_cuda
    B, H, L, R, S = 2, 4, 4096, 160, 128
    U, SV, ids, cs = _setup(B, H, L, R, S, seed=11, glm=False)
    cnts = torch.tensor([0, 3, 16, 17, 100, 128, 64, 1], dtype=torch.int32)
    ss, Lbuf = 64, 64 + S * 8 + 32
    cache = torch.zeros(B, H, Lbuf, 128, dtype=torch.bfloat16)
    want = cache.clone()
    rot = O.rebuild_keys(U, SV, ids, cs)
    for r in range(B * H):
        b, h = divmod(r, H)
        c = int(cnts[r]) * 8
        want[b, h, ss + c:ss + S * 8] = rot[b, h, c:]
    out = torch.zeros(B, H, S * 8, 128, dtype=torch.bfloat16, device=cuda)
    d_cache = cache.to(cuda)
    batch_gather_gemm_rotary_pos_emb_cuda(U.to(cuda), SV.to(cuda), cs.to(cuda), ids.to(cuda), out, 8, d_cache, ss,
                                          ss + S * 8, cnts.to(cuda))
    torch.cuda.synchronize()
    assert_rope_rows_close(d_cache.cpu(), want, frac_exact=0.90, what="gather-gemm + rope push")
    # pre-RoPE scratch: rows of tiles at or above the first missed tile are written (reference skips whole
    # 128-row tiles of hits, gemm_universal_batch_gather_indices.h:727-738)
    kpre, _ = O.rebuild_keys_pre_rope(U, SV, ids)
    for r in range(B * H):
        b, h = divmod(r, H)
        first = (int(cnts[r]) * 8 // 128) * 128
        if first < S * 8:
            assert_ulp_bf16(out[b, h, first:].cpu(), kpre[b, h, first:], ulps=1, frac_exact=0.95, what="pre-RoPE K")
        if first:
            assert float(out[b, h, :first].float().abs().max()) == 0.0
