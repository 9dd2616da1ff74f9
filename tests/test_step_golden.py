"""Step-level parity against the REFERENCE's own CUDA path (north_star: "checked against the reference's own CUDA
path on identical synthetic inputs"): tests/golden/step_golden.npz holds, for whole decode steps of one layer at the
BASELINE.json shapes (Llama-3.1-8B 122K, GLM-4-9B 128K with interleaved RoPE and 4x8 heads, Yi-9B 200K, and a
batch-3 case), what the unmodified reference extension + flash_attn_with_kvcache produced on a B200
(tests/golden/make_step_golden.py, tests/refchain.py).

* CPU: the oracle replays the steps and must reproduce them -> the oracle is pinned at step level.
* GPU: ``ShadowKVCache_CPU.decode_layer`` (the product path: fused append + scoring, selection, single-launch V fetch +
  K rebuild, attention) and the CUDA-graph runner ``DecodeStepGraph`` replay the same steps.

Bars: resident chunk set identical modulo top-k ties (a differing chunk must sit within 2 bf16 ulp of the S-th score;
at most MAX_TIE_CHUNKS per row -- torch.topk leaves the order of equal scores unspecified, ours is lowest index), hit
counts equal while the sets are, V bytes bit-exact (checksum + slot-by-slot against the host store), K rows within the
RoPE pair-norm bound of 1 bf16 ulp pre-RoPE, attention output max-abs 2e-2 / cosine 0.999."""
import os

import numpy as np
import pytest
import torch

from oracle import shadowkv_oracle as O
from shadowkv_b200 import synth
from tests.golden import step_cases as SC
from tests.tolerances import assert_close_bf16, assert_rope_rows_close

PATH = os.path.join(os.path.dirname(__file__), "golden", "step_golden.npz")
needs_golden = pytest.mark.skipif(not os.path.exists(PATH), reason="step_golden.npz not generated yet")
MAX_TIE_CHUNKS = 2


def _bf(a):
    return torch.from_numpy(np.ascontiguousarray(a)).view(torch.bfloat16)


def _tie_check(score, landmark_idx, mine, theirs, S, what):
    """`mine` / `theirs`: sets of chunk ids of one (b,h) row.  Every chunk in the symmetric difference must have a score
    within 2 bf16 ulp of the S-th largest score (score: bf16 [N] max-over-heads softmax of the oracle)."""
    diff = mine ^ theirs
    if not diff:
        return 0
    key = score.view(torch.int16).int()                      # positive bf16: bit pattern is monotone
    thr = int(key.sort(descending=True).values[S - 1])
    pos = {int(c): i for i, c in enumerate(landmark_idx.tolist())}
    for c in diff:
        assert abs(int(key[pos[c]]) - thr) <= 2, f"{what}: chunk {c} differs but is not a near-tie at the S-th score"
    return len(diff) // 2


def _sample_k(k_buf, resident, golden_sorted, ss):
    """Rows of `k_buf` [B,H,Lbuf,D] for the chunks the fixture sampled (every 8th of the golden sorted set, row id % 8);
    returns (rows [B,H,n,D], present mask [B,H,n])."""
    B, H, S = resident.shape
    ids = torch.from_numpy(golden_sorted.astype(np.int64))[..., ::SC.SAMPLE_EVERY]
    eq = resident.unsqueeze(-1) == ids.unsqueeze(-2)          # [B,H,S,n]
    present = eq.any(dim=-2)
    slot = eq.float().argmax(dim=-2)
    rows = ss + slot * 8 + (ids % 8)
    return k_buf.gather(2, rows.unsqueeze(-1).expand(-1, -1, -1, k_buf.shape[-1])), present


def _check_step(name, t, g, shape, lay, *, resident, cnts, k_buf, v_buf, v_chunks, out, score, landmark_idx, sync_prev):
    """One step of one case against the fixture.  Returns (worst tie count, sets identical)."""
    B, H, S = resident.shape
    gold = g[f"{name}/resident_sorted/{t}"]
    worst = 0
    for b in range(B):
        for h in range(H):
            worst = max(worst, _tie_check(score[b, h], landmark_idx[b, h], set(resident[b, h].tolist()),
                                          set(gold[b, h].tolist()), S, f"{name} step {t} row {b},{h}"))
    assert worst <= MAX_TIE_CHUNKS, f"{name} step {t}: {worst} chunks differ from the reference's selection"
    same = worst == 0
    if same and sync_prev:
        np.testing.assert_array_equal(np.asarray(cnts).reshape(-1), g[f"{name}/cnts/{t}"])
    ss, se = lay.sparse_start, lay.sparse_end
    # V: every slot holds exactly its chunk of the store; with identical sets the reference's checksum follows
    vs = v_buf[:, :, ss:se].reshape(B, H, S, -1)
    for b in range(B):
        for h in range(H):
            assert torch.equal(vs[b, h], v_chunks[b, h, resident[b, h]]), f"{name} step {t}: V slot bytes"
    if same:
        np.testing.assert_array_equal(v_buf[:, :, ss:se].contiguous().view(torch.int16).long().sum(dim=(2, 3)).numpy(),
                                      g[f"{name}/v_sum/{t}"])
    rows, present = _sample_k(k_buf, resident, gold, ss)
    want = _bf(g[f"{name}/k_rows/{t}"])
    assert float(present.float().mean()) > 0.95
    assert_rope_rows_close(rows[present], want[present], glm=shape.glm_rope, frac_exact=0.90,
                           what=f"{name} step {t}: K rows vs the reference's gather-GEMM + RoPE kernels")
    err, cos = assert_close_bf16(out, _bf(g[f"{name}/attn/{t}"]), what=f"{name} step {t}: attention vs flash_attn_with_kvcache")
    return worst, same, err


# ---------------------------------------------------------------------------------------------------------------------
@needs_golden
@pytest.mark.parametrize("name", ["llama31_8b_122k_b1", "llama31_8b_32k_b3"])
def test_oracle_reproduces_reference_cuda_steps(name):
    """CPU: pins the oracle to whole steps of the reference's CUDA path."""
    g = np.load(PATH)
    shape, B, ctx, cs, st = SC.build(name)
    lay = st["layout"]
    ora = O.LayerState(landmarks=st["k_landmark"], landmark_idx=st["k_landmark_idx"], U=st["U"], SV=st["SV"],
                       v_host=st["v_chunks"], k_buf=st["k_buf"].clone(), v_buf=st["v_buf"].clone(),
                       resident=st["resident"].clone())
    geo = O.Geometry(sparse_start=lay.sparse_start, sparse_end=lay.sparse_end)
    for t, (q, k, v) in enumerate(SC.step_inputs(name, shape, B)):
        geo.gen_offset = t
        gold = torch.from_numpy(g[f"{name}/resident_sorted/{t}"].astype(np.int64))
        out, info = O.decode_step(ora, geo, q, k, v, cs, selected_override=gold)
        own = info["own_selected"]
        for b in range(B):
            for h in range(shape.num_key_value_heads):
                n = _tie_check(info["score"][b, h], st["k_landmark_idx"][b, h], set(own[b, h].tolist()),
                               set(gold[b, h].tolist()), lay.select, f"{name} step {t}")
                assert n <= MAX_TIE_CHUNKS
        _check_step(name, t, g, shape, lay, resident=ora.resident, cnts=info["cnts"], k_buf=ora.k_buf, v_buf=ora.v_buf,
                    v_chunks=st["v_chunks"], out=out, score=info["score"], landmark_idx=st["k_landmark_idx"],
                    sync_prev=True)


def _product_cache(cuda, shape, B, ctx, st):
    from shadowkv_b200.kv_cache import ShadowKVCache_CPU
    cache = ShadowKVCache_CPU(shape.config(1), batch_size=B, max_length=ctx, device=str(cuda),
                              sparse_budget=SC.BUDGET, rank=SC.RANK)
    synth.install(cache, [st])
    return cache


@needs_golden
@pytest.mark.gpu
@pytest.mark.parametrize("name", list(SC.CASES))
def test_decode_layer_and_graph_reproduce_reference_cuda_steps(cuda, name):
    from shadowkv_b200.engine import DecodeStepGraph
    g = np.load(PATH)
    shape, B, ctx, cs, st = SC.build(name)
    lay = st["layout"]
    d_cs = cs.to(cuda)
    steps = SC.step_inputs(name, shape, B)
    cache = _product_cache(cuda, shape, B, ctx, st)
    eager, sync, worst_all, errs = [], True, 0, []
    for t, (q, k, v) in enumerate(steps):
        out = cache.decode_layer(0, q.to(cuda), k.to(cuda), v.to(cuda), d_cs)
        torch.cuda.synchronize()
        # the near-tie check needs scores: the oracle's (CPU, <= 1 bf16 ulp from the kernel's)
        _, softmax_o = O.landmark_scores(q, st["k_landmark"])
        score = softmax_o.float().max(dim=-2).values.to(torch.bfloat16)      # [B,H,N]; max of bf16 values is exact
        worst, same, err = _check_step(name, t, g, shape, lay, resident=cache.position_ids[0].cpu(),
                                       cnts=cache.cnts.cpu().numpy(), k_buf=cache.k_cache_buffer[0].cpu(),
                                       v_buf=cache.v_cache_buffer[0].cpu(), v_chunks=st["v_chunks"], out=out.cpu(),
                                       score=score, landmark_idx=st["k_landmark_idx"], sync_prev=sync)
        sync = sync and same
        worst_all = max(worst_all, worst)
        errs.append(err)
        eager.append((out.clone(), cache.position_ids[0].clone(), cache.k_cache_buffer[0].clone(),
                      cache.v_cache_buffer[0].clone()))
    print(f"\n[{name}] chunks differing from the reference's selection (top-k ties): worst {worst_all} of {lay.select}; "
          f"attention max-abs vs flash-attn per step: {[round(e, 5) for e in errs]}")
    del cache
    # ---- the same steps through the captured CUDA graph: bit-identical to the eager calls -----------------------------
    cache = _product_cache(cuda, shape, B, ctx, st)
    run = DecodeStepGraph(cache, d_cs, host_io=False).capture(warmup=0)
    for t, (q, k, v) in enumerate(steps):
        out = run.step_device(q.to(cuda).unsqueeze(0), k.to(cuda).unsqueeze(0), v.to(cuda).unsqueeze(0))
        run.stream.synchronize()
        o, res, kb, vb = eager[t]
        assert torch.equal(out[0], o), f"{name} step {t}: graph replay differs from the eager step"
        assert torch.equal(cache.position_ids[0], res)
        assert torch.equal(cache.k_cache_buffer[0], kb) and torch.equal(cache.v_cache_buffer[0], vb)
