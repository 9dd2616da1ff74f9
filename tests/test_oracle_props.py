"""Self-consistency of the oracle's floating-point restatement (CPU only, small sizes)."""
import torch

from oracle import shadowkv_oracle as O
from shadowkv_b200 import synth
from tests.tolerances import assert_close_bf16

SHAPE = synth.ModelShape("tiny", 1, 1024, 8, 2, 0, 0)      # 8 q heads, 2 kv heads, D=128


def _state(prompt_len=4096, budget=1024, seed=3, glm=False):
    cs = synth.cos_sin_table(prompt_len + 1024, 128, glm=glm)
    s = synth.layer_state(SHAPE, 1, prompt_len, seed=seed, rank=32, sparse_budget=budget, cos_sin=cs)
    lay = s["layout"]
    st = O.LayerState(landmarks=s["k_landmark"], landmark_idx=s["k_landmark_idx"], U=s["U"], SV=s["SV"],
                      v_host=s["v_chunks"], k_buf=s["k_buf"], v_buf=s["v_buf"], resident=s["resident"].clone())
    geo = O.Geometry(sparse_start=lay.sparse_start, sparse_end=lay.sparse_end)
    return s, st, geo, cs


def test_rope_bf16_close_to_float_rope():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 3, 16, 128, generator=g).to(torch.bfloat16)
    tok = torch.randint(0, 4096, (2, 3, 16), generator=g)
    for glm in (False, True):
        cs = synth.cos_sin_table(4096, 128, glm=glm)
        ref = O.rope_torch(x.float(), cs.float(), tok)
        for fused in (False, True):
            got = O.rope_bf16(x, cs[tok], fused=fused)
            assert_close_bf16(got, ref, max_abs=4e-2, what=f"rope glm={glm} fused={fused}")
    # rotation preserves the norm of every rotated pair (up to bf16 rounding)
    cs = synth.cos_sin_table(4096, 128)
    got = O.rope_bf16(x, cs[tok], fused=True).float()
    n0 = (x.float()[..., :64] ** 2 + x.float()[..., 64:] ** 2)
    n1 = (got[..., :64] ** 2 + got[..., 64:] ** 2)
    assert float(((n0 - n1).abs() / (n0 + 1e-3)).max()) < 0.05


def test_softmax_rows_sum_to_one_and_scores_round_trip():
    s, st, geo, cs = _state()
    q, _, _ = synth.queries(SHAPE, 1, seed=1)
    gemm_o, soft = O.landmark_scores(q, st.landmarks)
    assert gemm_o.dtype == torch.bfloat16 and soft.shape == gemm_o.shape
    assert float((soft.float().sum(-1) - 1).abs().max()) < 2e-2
    sel, top, score = O.select_chunks(soft, st.landmark_idx, 128)
    assert sel.shape == (1, 2, 128)
    for h in range(2):
        assert O.selection_equal_modulo_ties(score[0, h], top[0, h], 128)
        assert len(set(sel[0, h].tolist())) == 128


def test_cached_step_equals_uncached_step():
    """Appendix B of the survey: the chunk cache only permutes slots; attention output and the selected
    set must equal the cache-free pure-torch path (reference kv_cache.py:421-470)."""
    s, st, geo, cs = _state()
    lay = s["layout"]
    k_buf2, v_buf2 = s["k_buf"].clone(), s["v_buf"].clone()
    v_full = s["v_chunks"].view(1, 2, -1, 128)
    SV_rd = s["SV"].transpose(-1, -2).contiguous()
    base = None
    for step in range(3):
        q, kn, vn = synth.queries(SHAPE, 1, seed=10 + step, base=base, noise=0.3)
        base = q if base is None else base
        geo.gen_offset = step
        out2, sel2 = O.decode_step_uncached(s["k_landmark"], s["k_landmark_idx"], s["U"], SV_rd, v_full, k_buf2,
                                            v_buf2, geo, q, kn, vn,
                                            lambda x, tok: O.rope_bf16(x.to(torch.bfloat16), cs[tok]))
        out, info = O.decode_step(st, geo, q, kn, vn, cs, selected_override=sel2)
        for h in range(2):     # the oracle's own selection = the torch path's, up to near-ties at the boundary
            a, b = set(info["own_selected"][0, h].tolist()), set(sel2[0, h].tolist())
            idx_of = {int(c): i for i, c in enumerate(s["k_landmark_idx"][0, h].tolist())}
            sc = info["score"][0, h].float()
            kth = torch.sort(sc, descending=True).values[lay.select - 1]
            # the torch path rounds bf16(q.k) and then bf16(./sqrt(128)); the kernel path rounds once
            # (kv_cache.py:425-430 vs batch_gemm_softmax.cu:260-281), so boundary scores may move 2 bf16 ulp
            assert len(a ^ b) <= 8
            assert all(abs(float(sc[idx_of[c]]) - float(kth)) <= float(kth) * 2 ** -6 for c in a ^ b)
            assert set(st.resident[0, h].tolist()) == b
        assert_close_bf16(out, out2, what=f"step {step}")
        if step:
            assert int(info["cnts"].min()) > 0          # correlated queries -> cache hits
    # V rows in the sparse region are exactly the store's chunks named by the chunk table
    vs = st.v_buf[0, :, lay.sparse_start:lay.sparse_end].reshape(2, lay.select, -1)
    for h in range(2):
        assert torch.equal(vs[h], s["v_chunks"][0, h, st.resident[0, h]])


def test_attention_matches_sdpa():
    g = torch.Generator().manual_seed(5)
    q = torch.randn(2, 8, 1, 128, generator=g).to(torch.bfloat16)
    K = torch.randn(2, 2, 300, 128, generator=g).to(torch.bfloat16)
    V = torch.randn(2, 2, 300, 128, generator=g).to(torch.bfloat16)
    got = O.attention(q, K, V)
    Kr = K.float().repeat_interleave(4, dim=1)
    Vr = V.float().repeat_interleave(4, dim=1)
    ref = torch.nn.functional.scaled_dot_product_attention(q.float(), Kr, Vr).transpose(1, 2)
    assert_close_bf16(got, ref, what="attention")


def test_chunk_stats_chain_equals_torch_cosine_similarity():
    """The oracle spells torch's bf16 cosine_similarity out with explicit roundings (so that it does not depend on
    the installed torch); on this image the two must agree bit for bit, and the mean must equal torch.mean."""
    g = torch.Generator().manual_seed(3)
    k = (torch.randn(2, 3, 96, 1, 128, generator=g) + torch.randn(2, 3, 96, 8, 128, generator=g) *
         torch.rand(2, 3, 96, 1, 1, generator=g) * 2).to(torch.bfloat16)
    means, mincos = O.chunk_stats(k)
    lm = k.mean(dim=-2)
    assert torch.equal(lm, means)
    cos = torch.nn.functional.cosine_similarity(lm.unsqueeze(3).expand(-1, -1, -1, 8, -1), k, dim=-1)
    assert torch.equal(cos.min(dim=-1).values, mincos)
    out, rest = O.select_outliers(mincos, 24)
    ref = mincos.float().topk(24, largest=False).indices
    for b in range(2):
        for h in range(3):
            kth = mincos[b, h].float().sort().values[23]
            assert all(float(mincos[b, h, c]) <= float(kth) for c in out[b, h].tolist())
            assert set(ref[b, h].tolist()) ^ set(out[b, h].tolist()) <= set((mincos[b, h].float() == kth).nonzero().flatten().tolist())
            assert sorted(out[b, h].tolist() + rest[b, h].tolist()) == list(range(96))
