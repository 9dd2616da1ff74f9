"""Prefill-side state build (ShadowKVCache_CPU.get_svd / prefill_kv_cache, reference
models/kv_cache.py:666-737, 788-980) on the CUDA kernels of csrc/prefill.cu against the oracle's restatement."""
import ctypes as C

import pytest
import torch

from oracle import shadowkv_oracle as O
from shadowkv_b200 import synth
from tests.tolerances import assert_close_bf16, assert_ulp_bf16

pytestmark = pytest.mark.gpu


def _prompt(B, H, L, seed, structured=True):
    """Keys with chunk structure (so cosine similarities spread out) plus a few wild chunks (outliers)."""
    g = torch.Generator().manual_seed(seed)
    n = L // 8
    centre = torch.randn(B, H, n, 1, 128, generator=g)
    noise = torch.randn(B, H, n, 8, 128, generator=g)
    scale = torch.rand(B, H, n, 1, 1, generator=g) * 1.5 + 0.05          # per-chunk spread -> varied min-cos
    k = (centre + noise * scale).reshape(B, H, n * 8, 128)
    tail = torch.randn(B, H, L - n * 8, 128, generator=g)
    k = torch.cat([k, tail], dim=2).to(torch.bfloat16)
    v = torch.randn(B, H, L, 128, generator=g).to(torch.bfloat16)
    return k, v


@pytest.mark.parametrize("B,H,L", [(2, 2, 4096), (1, 8, 8200), (1, 4, 33000)])
def test_chunk_stats_and_outliers(cuda, B, H, L):
    from shadowkv_b200 import _lib
    k, _ = _prompt(B, H, L, seed=L)
    chunks = L // 8 - 4
    chunks -= chunks % 8
    dk = k.to(cuda)
    means = torch.empty(B, H, chunks, 128, dtype=torch.bfloat16, device=cuda)
    mincos = torch.empty(B, H, chunks, dtype=torch.bfloat16, device=cuda)
    _lib.check(_lib.lib.skv_chunk_stats(dk.data_ptr(), B, H, chunks, dk.stride(0), dk.stride(1), dk.stride(2),
                                        means.data_ptr(), mincos.data_ptr(), 0), "chunk_stats")
    torch.cuda.synchronize()
    want_m, want_c = O.chunk_stats(k[:, :, :chunks * 8].view(B, H, chunks, 8, 128))
    assert torch.equal(means.cpu(), want_m), "chunk means must be bit-exact (fp32 sum of 8 bf16 values)"
    assert_ulp_bf16(mincos, want_c, ulps=1, frac_exact=0.98, what="min cosine similarity per chunk")
    n_out = 48
    out = torch.empty(B, H, n_out, dtype=torch.int64, device=cuda)
    rest = torch.empty(B, H, chunks - n_out, dtype=torch.int64, device=cuda)
    _lib.check(_lib.lib.skv_select_outliers(mincos.data_ptr(), B * H, chunks, n_out, out.data_ptr(), rest.data_ptr(), 0),
               "select_outliers")
    torch.cuda.synchronize()
    # exact rule on the kernel's own values: n_out smallest, ties to the lowest chunk index, both lists ascending
    w_out, w_rest = O.select_outliers(mincos.cpu(), n_out)
    assert torch.equal(out.cpu(), w_out) and torch.equal(rest.cpu(), w_rest)
    # against the oracle's values: differences only among near-ties of the n_out-th smallest value
    o_out, _ = O.select_outliers(want_c, n_out)
    for b in range(B):
        for h in range(H):
            a, bb = set(out[b, h].tolist()), set(o_out[b, h].tolist())
            kth = torch.sort(want_c[b, h].float()).values[n_out - 1]
            for c in a ^ bb:
                assert abs(float(want_c[b, h, c]) - float(kth)) <= 2 ** -6


@pytest.mark.parametrize("model,B,L,budget", [("llama-3.1-8b", 2, 8192, 1024), ("yi-9b-200k", 1, 16400, 2048)])
def test_prefill_kv_cache_matches_oracle(cuda, model, B, L, budget):
    from shadowkv_b200.kv_cache import ShadowKVCache_CPU
    shape = synth.MODELS[model]
    H, Hq, layers = shape.num_key_value_heads, shape.num_attention_heads, 2
    cache = ShadowKVCache_CPU(shape.config(layers), batch_size=B, max_length=L, device=str(cuda), sparse_budget=budget)
    states = []
    for l in range(layers):
        k, v = _prompt(B, H, L, seed=100 + l)
        g = torch.Generator().manual_seed(7 + l)
        last_q = torch.randn(B, Hq, 128, generator=g).to(torch.bfloat16)
        k_pre = torch.randn(B, L, H * 128, generator=g).to(torch.bfloat16)           # pre-RoPE keys for the SVD
        cache.get_svd(k_pre.to(cuda), l)
        outl = torch.zeros(B, H, cache.outlier_chunk, dtype=torch.int64, device=cuda)
        cache.prefill_kv_cache(v.to(cuda), l, k.to(cuda), last_q.to(cuda), outliers_out=outl)
        torch.cuda.synchronize()
        states.append((k, v, last_q, k_pre, outl.cpu()))
    assert cache.kv_offset == L and cache.prefilled_batch == B
    for l, (k, v, last_q, k_pre, outl) in enumerate(states):
        own = O.prefill_state(k, v, last_q, budget=budget)
        for b in range(B):
            for h in range(H):                                   # outliers: equal to the oracle's modulo near-ties
                a, bb = set(outl[b, h].tolist()), set(own["own_outliers"][b, h].tolist())
                assert len(a ^ bb) <= 4, f"layer {l}: outlier sets differ in {len(a ^ bb)} chunks"
        sel = cache.position_ids[l].cpu()
        want = O.prefill_state(k, v, last_q, budget=budget, outliers_override=outl, selected_override=sel)
        assert (want["sparse_start"], want["sparse_end"]) == (cache.sparse_start, cache.sparse_end)
        assert torch.equal(cache.k_landmark_idx[l], want["rest_idx"])
        assert torch.equal(cache.k_landmark[l], want["landmarks"]), "landmarks = bf16 chunk means, bit-exact"
        nch = L // 8
        assert torch.equal(cache.v_cache_cpu[l][:, :, :nch], want["v_store"]), "pinned V store bytes"
        assert torch.equal(cache.k_cache_buffer[l].cpu(), want["k_buf"]), "K buffer (local | outliers | first selection)"
        assert torch.equal(cache.v_cache_buffer[l].cpu(), want["v_buf"]), "V buffer"
        for b in range(B):
            for h in range(H):                                   # first selection: the oracle's modulo top-k ties
                assert len(set(sel[b, h].tolist()) ^ set(want["own_selected"][b, h].tolist())) <= max(8, cache.select_sets // 16)
                assert torch.equal(sel[b, h], sel[b, h].sort().values) and len(set(sel[b, h].tolist())) == cache.select_sets
        # low-rank factors: the PRODUCT matches the reference's truncated SVD (the factors are sign-ambiguous)
        Uo, SVo = O.low_rank_factors(k_pre, cache.rank, H)
        rec = torch.einsum("blr,bhdr->blhd", cache.U[l].float(), cache.SV[l].float())
        rec_o = torch.einsum("blr,bhdr->blhd", Uo.float(), SVo.float())
        rel = float((rec - rec_o).norm() / rec_o.norm())
        assert rel < 2e-2, f"rank-{cache.rank} reconstruction differs from the reference SVD's by {rel:.3g}"
    # the freshly built state decodes
    cache.H2D()
    cs = synth.cos_sin_table(L + 1024, theta=shape.rope_theta, glm=shape.glm_rope).to(cuda)
    q, kn, vn = synth.queries(shape, B, seed=5)
    out = cache.decode_layer(0, q.to(cuda), kn.to(cuda), vn.to(cuda), cs)
    torch.cuda.synchronize()
    L_att = cache.sparse_end + 1
    ref = O.attention(q, cache.k_cache_buffer[0].cpu()[:, :, :L_att], cache.v_cache_buffer[0].cpu()[:, :, :L_att])
    assert_close_bf16(out, ref, max_abs=8e-3, what="first decode step after prefill")
