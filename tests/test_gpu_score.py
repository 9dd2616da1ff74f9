"""Landmark scoring + chunk selection against the oracle.

Scores are bf16-rounded fp32 dot products: a different fp32 summation order can move a value to the
neighbouring bf16, so D / Softmax are compared to 1 bf16 ulp (and must be mostly bit-identical), and the
selection is checked bit-exactly against the top-S rule applied to the kernel's OWN scores (ties: lowest
landmark index), plus near-tie agreement with the oracle's scores."""
import math

import numpy as np
import pytest
import torch

from oracle import shadowkv_oracle as O
from shadowkv_b200 import synth
from tests.tolerances import MAX_SELECTION_DIFF, assert_ulp_bf16

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("G,N", [(4, 4044), (8, 2048), (1, 1000), (16, 520), (4, 15560)])
def test_batch_gemm_softmax(cuda, G, N):
    from shadowkv_b200 import shadowkv
    B, H = 2, 3
    g = torch.Generator().manual_seed(N + G)
    q = torch.randn(B, H * G, 1, 128, generator=g).to(torch.bfloat16)
    lm = torch.randn(B, H, N, 128, generator=g).to(torch.bfloat16)
    gemm_o = torch.zeros(B, H, G, N, dtype=torch.bfloat16, device=cuda)
    soft_o = torch.zeros_like(gemm_o)
    T = (N + 255) // 256
    norm = torch.zeros(B * H, G, T, device=cuda)
    ssum = torch.zeros(B * H, G, T, device=cuda)
    shadowkv.batch_gemm_softmax(q.to(cuda), lm.to(cuda), gemm_o, norm, ssum, soft_o, B * H, G, N, 128,
                                1 / math.sqrt(128), 0)
    torch.cuda.synchronize()
    want_d, want_s = O.landmark_scores(q, lm)
    assert_ulp_bf16(gemm_o, want_d, ulps=1, frac_exact=0.97, what="D")
    assert_ulp_bf16(soft_o, want_s, ulps=2, frac_exact=0.90, what="Softmax")
    assert float((soft_o.float().sum(-1) - 1).abs().max()) < 2e-2


def _select(cuda, shape, N, S, seed, glm=False):
    from shadowkv_b200 import _lib
    import ctypes as C
    B = 2
    H, G = shape.num_key_value_heads, shape.num_attention_heads // shape.num_key_value_heads
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(B, H * G, 1, 128, generator=g).to(torch.bfloat16)
    lm = torch.randn(B, H, N, 128, generator=g).to(torch.bfloat16)
    idx = torch.stack([torch.randperm(N + 48, generator=g)[:N].sort().values for _ in range(B * H)]).view(B, H, N)
    resident = torch.stack([idx.view(B * H, N)[r][torch.randperm(N, generator=g)[:S]] for r in range(B * H)]).view(B, H, S)
    geom = _lib.Geom(batch=B, kv_heads=H, q_groups=G, head_dim=128, chunk=8, select=S, n_landmarks=N, rank=160,
                     seq_len=N * 8, buf_len=S * 8 + 600, sparse_start=448)
    ws = torch.empty(_lib.lib.skv_select_workspace_bytes(C.byref(geom)), dtype=torch.uint8, device=cuda)
    d_res = resident.to(cuda)
    off = torch.zeros(B * H * S, dtype=torch.int32, device=cuda)
    cnt = torch.zeros(B * H, dtype=torch.int32, device=cuda)
    sel = torch.zeros(B, H, S, dtype=torch.int64, device=cuda)
    score = torch.zeros(B, H, N, dtype=torch.bfloat16, device=cuda)
    d_q, d_lm, d_idx = q.to(cuda), lm.to(cuda), idx.to(cuda)     # named: a temporary would be freed before the launch
    _lib.check(_lib.lib.skv_select_chunks(C.byref(geom), d_q.data_ptr(), d_lm.data_ptr(),
                                          d_idx.data_ptr(), d_res.data_ptr(), off.data_ptr(), cnt.data_ptr(),
                                          sel.data_ptr(), score.data_ptr(), ws.data_ptr(), None,
                                          torch.cuda.current_stream().cuda_stream), "select")
    torch.cuda.synchronize()
    return q, lm, idx, resident, d_res.cpu(), off.cpu().view(B * H, S), cnt.cpu(), sel.cpu(), score.cpu()


@pytest.mark.parametrize("model,N,S", [("llama-3.1-8b", 4044, 256), ("glm-4-9b-1m", 2040, 128),
                                       ("llama-3.1-8b", 15560, 256), ("yi-9b-200k", 3000, 512),
                                       ("llama-3.1-8b", 1024, 1024),
                                       # more than 2048 landmarks per CTA of the selection cluster (several rounds):
                                       ("llama-3-8b-1m", 65480, 256),      # cfg3: 512K context
                                       ("yi-9b-200k", 25544, 256),         # cfg5: 200K context, G = 8
                                       ("glm-4-9b-1m", 16328, 256),        # cfg4: 128K context, G = 8
                                       ("llama-3.1-8b", 20004, 1024), ("llama-3.1-8b", 16388, 512)])
def test_select_chunks(cuda, model, N, S):
    shape = synth.MODELS[model]
    q, lm, idx, resident, new_res, off, cnt, sel, score = _select(cuda, shape, N, S, seed=N + S)
    B, H = sel.shape[:2]
    _, soft = O.landmark_scores(q, lm)
    want_sel, want_top, want_score = O.select_chunks(soft, idx, S)
    assert_ulp_bf16(score, want_score, ulps=2, frac_exact=0.90, what="max-over-groups softmax")
    pos_of = [{int(c): i for i, c in enumerate(idx.view(B * H, N)[r].tolist())} for r in range(B * H)]
    worst = 0
    for r in range(B * H):
        b, h = divmod(r, H)
        got_top = torch.tensor([pos_of[r][int(c)] for c in sel[b, h].tolist()])
        # 1. bit-exact rule on the kernel's own scores: top-S, ties to the lowest index, emitted ascending
        own = score[b, h].float()
        rule = torch.sort(own, descending=True, stable=True).indices[:S].sort().values
        assert torch.equal(got_top, rule), f"row {r}: selection is not top-S/lowest-index of the kernel's scores"
        # 2. against the oracle's scores: differences only among near-ties of the S-th value
        a, bset = set(got_top.tolist()), set(want_top[b, h].tolist())
        kth = float(torch.sort(want_score[b, h].float(), descending=True).values[S - 1])
        worst = max(worst, len(a ^ bset) // 2)
        assert len(a ^ bset) // 2 <= MAX_SELECTION_DIFF, f"row {r}: {len(a ^ bset) // 2} chunks differ"
        for i in a ^ bset:          # ... and only where the oracle's score is within 1 bf16 ulp of its S-th value
            assert abs(float(want_score[b, h, i]) - kth) <= kth * 2 ** -7
        # 3. chunk-table update is bit-exact given the kernel's selection
        nk, no, nc = O.reorder_keys_and_compute_offsets(resident.view(B * H, S)[r:r + 1].numpy(),
                                                        sel.view(B * H, S)[r:r + 1].numpy())
        np.testing.assert_array_equal(new_res.view(B * H, S)[r].numpy(), nk[0])
        np.testing.assert_array_equal(off[r].numpy(), no[0])
        assert int(cnt[r]) == int(nc[0])
    print(f"\n[{model} N={N} S={S}] chunks selected differently from the oracle (1-ulp score differences at the S-th "
          f"value): worst row {worst} of {S}")


@pytest.mark.parametrize("model,N,S", [("llama-3.1-8b", 15560, 256), ("glm-4-9b-1m", 16328, 256),
                                       ("llama-3-8b-1m", 65480, 256), ("llama-3.1-8b", 4044, 512)])
def test_select_threshold_hint_never_changes_the_result(cuda, model, N, S):
    """skv_geom.select_hint is a speed hint: a window of single-key histogram bins around the previous threshold settles
    the top-S threshold in one exchange when the new threshold lies inside it; outside it the two-level search runs.  The
    selection, the chunk table and the hit counts must be the same for ANY hint content, and the word must come back as
    the threshold key of this selection."""
    from shadowkv_b200 import _lib
    import ctypes as C
    shape = synth.MODELS[model]
    B = 2
    H, G = shape.num_key_value_heads, shape.num_attention_heads // shape.num_key_value_heads
    g = torch.Generator().manual_seed(N * 3 + S)
    q = torch.randn(B, H * G, 1, 128, generator=g).to(torch.bfloat16).to(cuda)
    lm = torch.randn(B, H, N, 128, generator=g).to(torch.bfloat16).to(cuda)
    idx = torch.stack([torch.randperm(N + 48, generator=g)[:N].sort().values for _ in range(B * H)]).view(B, H, N).to(cuda)
    resident = torch.stack([idx.view(B * H, N)[r][torch.randperm(N, generator=g)[:S].to(cuda)] for r in range(B * H)]).view(B, H, S)
    geom = _lib.Geom(batch=B, kv_heads=H, q_groups=G, head_dim=128, chunk=8, select=S, n_landmarks=N, rank=160,
                     seq_len=N * 8, buf_len=S * 8 + 600, sparse_start=448)
    ws = torch.empty(_lib.lib.skv_select_workspace_bytes(C.byref(geom)), dtype=torch.uint8, device=cuda)

    def run(hint):
        res = resident.clone()
        off = torch.zeros(B * H * S, dtype=torch.int32, device=cuda)
        cnt = torch.zeros(B * H, dtype=torch.int32, device=cuda)
        sel = torch.zeros(B, H, S, dtype=torch.int64, device=cuda)
        score = torch.zeros(B, H, N, dtype=torch.bfloat16, device=cuda)
        geom.select_hint = hint.data_ptr() if hint is not None else None
        _lib.check(_lib.lib.skv_select_chunks(C.byref(geom), q.data_ptr(), lm.data_ptr(), idx.data_ptr(), res.data_ptr(),
                                              off.data_ptr(), cnt.data_ptr(), sel.data_ptr(), score.data_ptr(), ws.data_ptr(),
                                              None, torch.cuda.current_stream().cuda_stream), "select")
        torch.cuda.synchronize()
        return res.cpu(), off.cpu(), cnt.cpu(), sel.cpu(), score.cpu()

    base = run(None)
    # the threshold key of every row, from the kernel's own scores: the S-th largest key (bf16 bits are monotone for p >= 0)
    keys = base[4].view(torch.int16).int().view(B * H, N)
    thr = torch.sort(keys, dim=1, descending=True).values[:, S - 1].to(torch.int32)
    assert int(thr.min()) > 0
    cases = {"zero (first token)": torch.zeros_like(thr), "exact": thr.clone(), "a few steps off": thr + torch.arange(B * H, dtype=torch.int32) * 7 - 20,
             "window edges": thr + torch.tensor([126, -127] * (B * H // 2), dtype=torch.int32),
             "just outside": thr + torch.tensor([127, -128] * (B * H // 2), dtype=torch.int32),
             "far above": thr + 3000, "far below": (thr - 3000).clamp(min=1), "huge": torch.full_like(thr, 65535)}
    for name, h in cases.items():
        hint = h.to(cuda)
        got = run(hint)
        for a, b, what in zip(got[:4], base[:4], ("resident", "offsets", "cnts", "selected")):
            assert torch.equal(a, b), f"hint '{name}': {what} differs from the run without a hint"
        assert torch.equal(hint.cpu(), thr), f"hint '{name}': the word must come back as this selection's threshold key"


def test_select_massive_ties(cuda):
    """All landmarks identical -> every score ties; the rule must pick the S lowest indices."""
    from shadowkv_b200 import _lib
    import ctypes as C
    B, H, G, N, S = 1, 2, 4, 3000, 256
    q = torch.randn(B, H * G, 1, 128).to(torch.bfloat16)
    lm = torch.ones(B, H, N, 128, dtype=torch.bfloat16) * 0.01
    idx = torch.arange(N).expand(B, H, N).contiguous() * 3
    geom = _lib.Geom(batch=B, kv_heads=H, q_groups=G, head_dim=128, chunk=8, select=S, n_landmarks=N, rank=160,
                     seq_len=N * 8, buf_len=4000, sparse_start=448)
    ws = torch.empty(_lib.lib.skv_select_workspace_bytes(C.byref(geom)), dtype=torch.uint8, device=cuda)
    res = torch.full((B, H, S), -1, dtype=torch.int64, device=cuda)
    off = torch.zeros(B * H * S, dtype=torch.int32, device=cuda)
    cnt = torch.zeros(B * H, dtype=torch.int32, device=cuda)
    d_q, d_lm, d_idx = q.to(cuda), lm.to(cuda), idx.to(cuda)
    _lib.check(_lib.lib.skv_select_chunks(C.byref(geom), d_q.data_ptr(), d_lm.data_ptr(),
                                          d_idx.data_ptr(), res.data_ptr(), off.data_ptr(), cnt.data_ptr(), 0, 0,
                                          ws.data_ptr(), None, torch.cuda.current_stream().cuda_stream), "select")
    torch.cuda.synchronize()
    assert (cnt.cpu() == 0).all()
    assert torch.equal(res.cpu(), (torch.arange(S) * 3).expand(B, H, S))


def test_select_underflowed_probabilities_and_ragged_tail(cuda):
    """One landmark takes all the probability mass; every other softmax value underflows to bf16 zero, so the
    threshold key is 0 and N is not a multiple of 8 (the last 16-byte group of the last cluster slice is ragged):
    the padding must never be selected -- the answer is that landmark plus the S-1 lowest indices."""
    from shadowkv_b200 import _lib
    import ctypes as C
    B, H, G, N, S = 1, 2, 4, 1004, 128
    q = torch.full((B, H * G, 1, 128), 4.0, dtype=torch.bfloat16)
    lm = torch.full((B, H, N, 128), -4.0, dtype=torch.bfloat16)
    lm[:, :, 777] = 4.0
    idx = (torch.arange(N) * 2 + 5).expand(B, H, N).contiguous()
    geom = _lib.Geom(batch=B, kv_heads=H, q_groups=G, head_dim=128, chunk=8, select=S, n_landmarks=N, rank=160,
                     seq_len=N * 16 + 64, buf_len=2000, sparse_start=448)
    ws = torch.empty(_lib.lib.skv_select_workspace_bytes(C.byref(geom)), dtype=torch.uint8, device=cuda)
    res = torch.full((B, H, S), -1, dtype=torch.int64, device=cuda)
    off = torch.zeros(B * H * S, dtype=torch.int32, device=cuda)
    cnt = torch.zeros(B * H, dtype=torch.int32, device=cuda)
    sel = torch.full((B, H, S), -7, dtype=torch.int64, device=cuda)
    score = torch.zeros(B, H, N, dtype=torch.bfloat16, device=cuda)
    d_q, d_lm, d_idx = q.to(cuda), lm.to(cuda), idx.to(cuda)
    _lib.check(_lib.lib.skv_select_chunks(C.byref(geom), d_q.data_ptr(), d_lm.data_ptr(), d_idx.data_ptr(), res.data_ptr(),
                                          off.data_ptr(), cnt.data_ptr(), sel.data_ptr(), score.data_ptr(), ws.data_ptr(),
                                          None, torch.cuda.current_stream().cuda_stream), "select")
    torch.cuda.synchronize()
    sc = score.float().cpu()
    assert float(sc[0, 0, 777]) == 1.0 and float(sc[0, 0].sum()) == 1.0        # everything else underflowed
    want = torch.cat([torch.arange(127), torch.tensor([777])]) * 2 + 5
    for h in range(H):
        assert torch.equal(sel[0, h].cpu(), want), "underflow ties must resolve to the lowest landmark indices"
        assert torch.equal(res[0, h].cpu(), want)


@pytest.mark.parametrize("model,B,N,S", [("llama-3.1-8b", 1, 15560, 256),      # 8 rows: 16-CTA clusters (cfg2)
                                         ("llama-3.1-8b", 2, 15560, 256),      # 16 rows: 8-CTA clusters
                                         ("glm-4-9b-1m", 1, 16328, 256),       # 4 rows x G = 8
                                         ("llama-3-8b-1m", 4, 65480, 256),     # cfg3: 64 KiB of scores per CTA
                                         ("yi-9b-200k", 3, 25544, 128),
                                         ("llama-3.1-8b", 1, 1000, 128)])      # fewer tiles than CTAs, ragged N
def test_fused_score_select_experiment_agrees_with_the_two_launches(cuda, monkeypatch, model, B, N, S):
    """The one-launch experiment (SKV_FUSED_SELECT=1, score_select.cu; measured slower, not the product path) against the
    product path skv_score_landmarks + skv_select_from_scores: same bf16 scores; the softmax sums are accumulated in a
    different order (per CTA, then over the cluster), so a key may differ by 1 bf16 ulp and the selection only at the
    threshold value."""
    from shadowkv_b200 import _lib
    import ctypes as C
    shape = synth.MODELS[model]
    H, G = shape.num_key_value_heads, shape.num_attention_heads // shape.num_key_value_heads
    g = torch.Generator().manual_seed(N + S + B)
    q = torch.randn(B, H * G, 1, 128, generator=g).to(torch.bfloat16).to(cuda)
    lm = torch.randn(B, H, N, 128, generator=g).to(torch.bfloat16).to(cuda)
    idx = torch.stack([torch.randperm(N + 48, generator=g)[:N].sort().values for _ in range(B * H)]).view(B, H, N)
    resident = torch.stack([idx.view(B * H, N)[r][torch.randperm(N, generator=g)[:S]] for r in range(B * H)]).view(B, H, S)
    geom = _lib.Geom(batch=B, kv_heads=H, q_groups=G, head_dim=128, chunk=8, select=S, n_landmarks=N, rank=160,
                     seq_len=N * 8, buf_len=S * 8 + 600, sparse_start=448, table_mode=1)
    ws = torch.empty(_lib.lib.skv_select_workspace_bytes(C.byref(geom)), dtype=torch.uint8, device=cuda)
    d_idx = idx.to(cuda)
    st = torch.cuda.current_stream().cuda_stream
    outs = []
    for fused in ("1", "0"):
        monkeypatch.setenv("SKV_FUSED_SELECT", fused)
        res = resident.to(cuda).clone()
        off = torch.zeros(B * H * S, dtype=torch.int32, device=cuda)
        cnt = torch.zeros(B * H, dtype=torch.int32, device=cuda)
        sel = torch.zeros(B, H, S, dtype=torch.int64, device=cuda)
        score = torch.zeros(B, H, N, dtype=torch.bfloat16, device=cuda)
        _lib.check(_lib.lib.skv_select_chunks(C.byref(geom), q.data_ptr(), lm.data_ptr(), d_idx.data_ptr(), res.data_ptr(),
                                              off.data_ptr(), cnt.data_ptr(), sel.data_ptr(), score.data_ptr(),
                                              ws.data_ptr(), None, st), "select_chunks")
        torch.cuda.synchronize()
        outs.append((res.cpu(), off.cpu(), cnt.cpu(), sel.cpu(), score.cpu()))
    (r1, o1, c1, s1, k1), (r2, o2, c2, s2, k2) = outs
    kd = (k1.view(torch.int16).int() - k2.view(torch.int16).int()).abs()
    assert int(kd.max()) <= 1 and float((kd == 0).float().mean()) >= 0.999, "selection keys differ by more than the summation order explains"
    for r in range(B * H):
        a, b = set(s1.view(B * H, S)[r].tolist()), set(s2.view(B * H, S)[r].tolist())
        assert len(a) == S and len(a ^ b) <= 4, f"row {r}: {len(a ^ b)} chunks selected differently"
    if torch.equal(s1, s2):
        assert torch.equal(c1, c2) and torch.equal(r1, r2), "chunk table differs"
        nm = (S - c1.long()).view(-1, 1)
        live = torch.arange(S).view(1, -1) < nm                   # offsets[i] is defined for i < number of misses
        assert torch.equal(o1.view(B * H, S)[live], o2.view(B * H, S)[live])


@pytest.mark.parametrize("model,B,N,S", [("llama-3-8b-1m", 2, 65480, 256), ("yi-9b-200k", 2, 25544, 128),
                                         ("llama-3.1-8b", 1, 15560, 256)])
def test_select_wide_cta_equals_base(cuda, monkeypatch, model, B, N, S):
    """select_kernel with 512 threads per CTA (long slices: contexts of 256K tokens and more) against the 256-thread
    instantiation on the same scores: same keys, same selection, same chunk table, bit for bit."""
    from shadowkv_b200 import _lib
    import ctypes as C
    shape = synth.MODELS[model]
    H, G = shape.num_key_value_heads, shape.num_attention_heads // shape.num_key_value_heads
    g = torch.Generator().manual_seed(N + S + B + 1)
    q = torch.randn(B, H * G, 1, 128, generator=g).to(torch.bfloat16).to(cuda)
    lm = torch.randn(B, H, N, 128, generator=g).to(torch.bfloat16).to(cuda)
    idx = torch.stack([torch.randperm(N + 48, generator=g)[:N].sort().values for _ in range(B * H)]).view(B, H, N)
    resident = torch.stack([idx.view(B * H, N)[r][torch.randperm(N, generator=g)[:S]] for r in range(B * H)]).view(B, H, S)
    geom = _lib.Geom(batch=B, kv_heads=H, q_groups=G, head_dim=128, chunk=8, select=S, n_landmarks=N, rank=160,
                     seq_len=N * 8, buf_len=S * 8 + 600, sparse_start=448, table_mode=1)
    ws = torch.empty(_lib.lib.skv_select_workspace_bytes(C.byref(geom)), dtype=torch.uint8, device=cuda)
    d_idx = idx.to(cuda)
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(_lib.lib.skv_score_landmarks(C.byref(geom), q.data_ptr(), lm.data_ptr(), ws.data_ptr(), None, st), "score")
    outs = []
    for wide_from in ("1", "1000000"):
        monkeypatch.setenv("SKV_SELECT_WIDE_FROM", wide_from)
        res = resident.to(cuda).clone()
        off = torch.zeros(B * H * S, dtype=torch.int32, device=cuda)
        cnt = torch.zeros(B * H, dtype=torch.int32, device=cuda)
        sel = torch.zeros(B, H, S, dtype=torch.int64, device=cuda)
        score = torch.zeros(B, H, N, dtype=torch.bfloat16, device=cuda)
        _lib.check(_lib.lib.skv_select_from_scores(C.byref(geom), d_idx.data_ptr(), res.data_ptr(), off.data_ptr(),
                                                   cnt.data_ptr(), sel.data_ptr(), score.data_ptr(), ws.data_ptr(), st),
                   "select_from_scores")
        torch.cuda.synchronize()
        outs.append((res.cpu(), off.cpu(), cnt.cpu(), sel.cpu(), score.cpu()))
    (r1, o1, c1, s1, k1), (r2, o2, c2, s2, k2) = outs
    assert torch.equal(k1.view(torch.int16), k2.view(torch.int16)), "selection keys differ"
    assert torch.equal(s1, s2) and torch.equal(c1, c2) and torch.equal(r1, r2), "selection / chunk table differs"
    live = torch.arange(S).view(1, -1) < (S - c1.long()).view(-1, 1)
    assert torch.equal(o1.view(B * H, S)[live], o2.view(B * H, S)[live])


def _fast_status(ws, B, H, G, N):
    """The per-row words select_fast_kernel leaves behind the partial sums of the select workspace (1 = row done there,
    <= 0 = left to the cluster kernel: -1 no usable hint, -2 candidate list overflow, -3 / -4 threshold below the window,
    -5 above it)."""
    al = lambda x: (x + 255) // 256 * 256   # noqa: E731
    ldn, T = (N + 7) // 8 * 8, (N + 127) // 128
    off = al(B * H * G * ldn * 2) + 2 * al(B * H * G * T * 4)
    return ws[off:off + B * H * 4].view(torch.int32).cpu()


@pytest.mark.parametrize("model,N,S", [("llama-3.1-8b", 15560, 256), ("glm-4-9b-1m", 16328, 256), ("llama-3.1-8b", 4044, 512),
                                       ("llama-3.1-8b", 1024, 128)])
def test_select_fast_path_equals_cluster_kernel(cuda, monkeypatch, model, N, S):
    """select_fast_kernel (one CTA per row: log-domain candidate filter below the hint window, exact keys only for the
    candidates, shared-memory histogram) must leave exactly what the cluster kernel leaves -- chunk table, offsets, hit
    counts, selected ids, hint word -- whenever it settles a row, and must leave the row alone otherwise."""
    from shadowkv_b200 import _lib
    import ctypes as C
    monkeypatch.setenv("SKV_SELECT_FAST", "1")       # an experiment, off by default (measured slower than the cluster kernel)
    shape = synth.MODELS[model]
    B = 2
    H, G = shape.num_key_value_heads, shape.num_attention_heads // shape.num_key_value_heads
    g = torch.Generator().manual_seed(N * 5 + S)
    q = torch.randn(B, H * G, 1, 128, generator=g).to(torch.bfloat16).to(cuda)
    lm = torch.randn(B, H, N, 128, generator=g).to(torch.bfloat16).to(cuda)
    idx = torch.stack([torch.randperm(N + 48, generator=g)[:N].sort().values for _ in range(B * H)]).view(B, H, N).to(cuda)
    resident = torch.stack([idx.view(B * H, N)[r][torch.randperm(N, generator=g)[:S].to(cuda)] for r in range(B * H)]).view(B, H, S)
    geom = _lib.Geom(batch=B, kv_heads=H, q_groups=G, head_dim=128, chunk=8, select=S, n_landmarks=N, rank=160,
                     seq_len=N * 8, buf_len=S * 8 + 600, sparse_start=448)
    ws = torch.empty(_lib.lib.skv_select_workspace_bytes(C.byref(geom)), dtype=torch.uint8, device=cuda)

    def run(hint, with_scores=False):
        res = resident.clone()
        off = torch.zeros(B * H * S, dtype=torch.int32, device=cuda)
        cnt = torch.zeros(B * H, dtype=torch.int32, device=cuda)
        sel = torch.zeros(B, H, S, dtype=torch.int64, device=cuda)
        score = torch.zeros(B, H, N, dtype=torch.bfloat16, device=cuda)
        geom.select_hint = hint.data_ptr() if hint is not None else None
        _lib.check(_lib.lib.skv_select_chunks(C.byref(geom), q.data_ptr(), lm.data_ptr(), idx.data_ptr(), res.data_ptr(),
                                              off.data_ptr(), cnt.data_ptr(), sel.data_ptr(),
                                              score.data_ptr() if with_scores else None, ws.data_ptr(),
                                              None, torch.cuda.current_stream().cuda_stream), "select")
        torch.cuda.synchronize()
        return res.cpu(), off.cpu(), cnt.cpu(), sel.cpu(), score.cpu()

    base = run(None, with_scores=True)               # the cluster kernel alone (no hint, score output requested)
    keys = base[4].view(torch.int16).int().view(B * H, N)
    thr = torch.sort(keys, dim=1, descending=True).values[:, S - 1].to(torch.int32)
    R = B * H
    cases = {"zero (first token)": (torch.zeros_like(thr), -1), "exact": (thr.clone(), 1),
             "a few steps off": (thr + torch.arange(R, dtype=torch.int32) * 7 - 20, 1),
             "window edges": (thr + torch.tensor([126, -127] * (R // 2), dtype=torch.int32), 1),
             "just outside": (thr + torch.tensor([127, -128] * (R // 2), dtype=torch.int32), None),
             "far above": (thr + 3000, None), "far below": ((thr - 3000).clamp(min=1), None)}
    took = {}
    for name, (h, want_status) in cases.items():
        hint = h.to(cuda)
        got = run(hint)
        status = _fast_status(ws, B, H, G, N)
        took[name] = status.tolist()
        for a, b, what in zip(got[:4], base[:4], ("resident", "offsets", "cnts", "selected")):
            assert torch.equal(a, b), f"hint '{name}' (fast status {status.tolist()}): {what} differs from the cluster kernel's"
        assert torch.equal(hint.cpu(), thr), f"hint '{name}': the word must come back as this selection's threshold key"
        if want_status == 1:
            assert bool(((status == 1) | (status == -2)).all()), f"hint '{name}': rows not settled by the fast path: {status.tolist()}"
        elif want_status is not None:
            assert bool((status == want_status).all()), f"hint '{name}': fast status {status.tolist()}"
        else:
            assert bool((status <= 0).all()), f"hint '{name}': the fast path must leave these rows alone: {status.tolist()}"
        # A/B switch: the cluster kernel with the same hint
        monkeypatch.setenv("SKV_SELECT_FAST", "0")
        ref = run(h.to(cuda))
        monkeypatch.setenv("SKV_SELECT_FAST", "1")
        for a, b, what in zip(got[:4], ref[:4], ("resident", "offsets", "cnts", "selected")):
            assert torch.equal(a, b), f"hint '{name}': {what} differs between SKV_SELECT_FAST=1 and 0"
    print(f"\n[{model} N={N} S={S}] fast-path status per case: {took}")
    assert any(v == 1 for v in took["exact"]), "the fast path never fired with an exact hint"


def test_select_fast_path_ties_take_the_lowest_indices(cuda, monkeypatch):
    """All landmarks identical -> every key ties at the threshold; with an exact hint the single-CTA path takes the S
    lowest landmark indices, like the cluster kernel."""
    from shadowkv_b200 import _lib
    import ctypes as C
    monkeypatch.setenv("SKV_SELECT_FAST", "1")
    B, H, G, N, S = 1, 2, 4, 1500, 256
    q = torch.randn(B, H * G, 1, 128).to(torch.bfloat16)
    lm = torch.ones(B, H, N, 128, dtype=torch.bfloat16) * 0.01
    idx = torch.arange(N).expand(B, H, N).contiguous() * 3
    geom = _lib.Geom(batch=B, kv_heads=H, q_groups=G, head_dim=128, chunk=8, select=S, n_landmarks=N, rank=160,
                     seq_len=N * 8, buf_len=4000, sparse_start=448)
    ws = torch.empty(_lib.lib.skv_select_workspace_bytes(C.byref(geom)), dtype=torch.uint8, device=cuda)
    hint = torch.zeros(B * H, dtype=torch.int32, device=cuda)
    geom.select_hint = hint.data_ptr()
    d_q, d_lm, d_idx = q.to(cuda), lm.to(cuda), idx.to(cuda)
    sels = []
    for rep in range(2):                             # first call: no hint (cluster kernel); second: its threshold as the hint
        res = torch.full((B, H, S), -1, dtype=torch.int64, device=cuda)
        off = torch.zeros(B * H * S, dtype=torch.int32, device=cuda)
        cnt = torch.zeros(B * H, dtype=torch.int32, device=cuda)
        sel = torch.zeros(B, H, S, dtype=torch.int64, device=cuda)
        _lib.check(_lib.lib.skv_select_chunks(C.byref(geom), d_q.data_ptr(), d_lm.data_ptr(), d_idx.data_ptr(),
                                              res.data_ptr(), off.data_ptr(), cnt.data_ptr(), sel.data_ptr(), None,
                                              ws.data_ptr(), None, torch.cuda.current_stream().cuda_stream), "select")
        torch.cuda.synchronize()
        sels.append(sel.cpu())
        status = _fast_status(ws, B, H, G, N)
        assert bool((status == (1 if rep else -1)).all()), status
    want = (torch.arange(S) * 3).expand(B, H, S)
    assert torch.equal(sels[0], want) and torch.equal(sels[1], want)


@pytest.mark.parametrize("model,N,S", [("llama-3.1-8b", 15560, 256), ("yi-9b-200k", 25544, 256), ("llama-3-8b-1m", 65480, 256),
                                       ("glm-4-9b-1m", 16328, 512), ("llama-3.1-8b", 2050, 128)])
def test_select_candidate_filter_equals_full_keys(cuda, monkeypatch, model, N, S):
    """The hinted selection forms exact keys only for the landmarks its log-domain filter cannot rule out
    (SKV_SELECT_FILTER, default on).  Chunk table, offsets, hit counts, selected ids and the hint word must equal the run
    that forms every key (SKV_SELECT_FILTER=0) and the run without a hint, for hints inside, at the edges of and outside
    the window, at every CTA width (256 / 512 / 1024 threads: N = 15560 / 25544 / 65480) and for ragged slices."""
    from shadowkv_b200 import _lib
    import ctypes as C
    shape = synth.MODELS[model]
    B = 2
    H, G = shape.num_key_value_heads, shape.num_attention_heads // shape.num_key_value_heads
    g = torch.Generator().manual_seed(N * 7 + S)
    q = torch.randn(B, H * G, 1, 128, generator=g).to(torch.bfloat16).to(cuda)
    lm = torch.randn(B, H, N, 128, generator=g).to(torch.bfloat16).to(cuda)
    idx = torch.stack([torch.randperm(N + 48, generator=g)[:N].sort().values for _ in range(B * H)]).view(B, H, N).to(cuda)
    resident = torch.stack([idx.view(B * H, N)[r][torch.randperm(N, generator=g)[:S].to(cuda)] for r in range(B * H)]).view(B, H, S)
    geom = _lib.Geom(batch=B, kv_heads=H, q_groups=G, head_dim=128, chunk=8, select=S, n_landmarks=N, rank=160,
                     seq_len=N * 8, buf_len=S * 8 + 600, sparse_start=448)
    ws = torch.empty(_lib.lib.skv_select_workspace_bytes(C.byref(geom)), dtype=torch.uint8, device=cuda)

    def run(hint, with_scores=False):
        res = resident.clone()
        off = torch.zeros(B * H * S, dtype=torch.int32, device=cuda)
        cnt = torch.zeros(B * H, dtype=torch.int32, device=cuda)
        sel = torch.zeros(B, H, S, dtype=torch.int64, device=cuda)
        score = torch.zeros(B, H, N, dtype=torch.bfloat16, device=cuda)
        geom.select_hint = hint.data_ptr() if hint is not None else None
        _lib.check(_lib.lib.skv_select_chunks(C.byref(geom), q.data_ptr(), lm.data_ptr(), idx.data_ptr(), res.data_ptr(),
                                              off.data_ptr(), cnt.data_ptr(), sel.data_ptr(),
                                              score.data_ptr() if with_scores else None, ws.data_ptr(),
                                              None, torch.cuda.current_stream().cuda_stream), "select")
        torch.cuda.synchronize()
        return res.cpu(), off.cpu(), cnt.cpu(), sel.cpu(), score.cpu()

    base = run(None, with_scores=True)
    keys = base[4].view(torch.int16).int().view(B * H, N)
    thr = torch.sort(keys, dim=1, descending=True).values[:, S - 1].to(torch.int32)
    R = B * H
    cases = {"exact": thr.clone(), "a few steps off": thr + torch.arange(R, dtype=torch.int32) * 5 - 12,
             "window edges": thr + torch.tensor([126, -127] * (R // 2), dtype=torch.int32),
             "just outside": thr + torch.tensor([127, -128] * (R // 2), dtype=torch.int32),
             "far above": thr + 3000, "far below": (thr - 3000).clamp(min=1)}
    for name, h in cases.items():
        outs = {}
        for flt in ("1", "0"):
            monkeypatch.setenv("SKV_SELECT_FILTER", flt)
            hint = h.to(cuda)
            outs[flt] = run(hint)
            assert torch.equal(hint.cpu(), thr), f"hint '{name}' filter={flt}: the word must come back as the threshold key"
        for what, a, b, c in zip(("resident", "offsets", "cnts", "selected"), outs["1"][:4], outs["0"][:4], base[:4]):
            assert torch.equal(a, c), f"hint '{name}': {what} with the candidate filter differs from the run without a hint"
            assert torch.equal(b, c), f"hint '{name}': {what} without the filter differs from the run without a hint"
