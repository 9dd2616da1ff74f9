"""skv_fetch_rebuild_attend: the V miss fetch, the tcgen05 K rebuild and the attention of a layer as ONE launch
(north_star: "(5) is fused with (2) so rebuilt keys never round-trip HBM") against the launches it replaces
(skv_fetch_and_rebuild + skv_attention_decode): same chunk table, same K / V bytes bit for bit, attention output within
the fp32-summation-order tolerance; and against the oracle through decode_layer.  Reference join: models/base.py:326-341."""
import numpy as np
import pytest
import torch

from oracle import shadowkv_oracle as O
from shadowkv_b200 import synth
from tests.tolerances import assert_close_bf16

pytestmark = pytest.mark.gpu


def _build(cuda, model, B, prompt_len, budget, layers, rank=160, cls=None, **kw):
    from shadowkv_b200.kv_cache import ShadowKVCache_CPU
    cls = cls or ShadowKVCache_CPU
    shape = synth.MODELS[model]
    uncached = cls is not ShadowKVCache_CPU
    cs = synth.cos_sin_table(prompt_len + 1024, 128, theta=shape.rope_theta, glm=shape.glm_rope)
    states = [synth.layer_state(shape, B, prompt_len, seed=2000 + l, rank=rank, sparse_budget=budget, cos_sin=cs,
                                uncached=uncached) for l in range(layers)]
    cache = cls(shape.config(layers), batch_size=B, max_length=prompt_len, device=str(cuda), sparse_budget=budget,
                rank=rank, **kw)
    synth.install(cache, states)
    return shape, cache, states, cs


@pytest.mark.parametrize("model,B,prompt_len,budget,ways,rank", [
    ("llama-3.1-8b", 2, 8192, 1024, 0, 160),          # S = 128, G = 4
    ("llama-3.1-8b", 1, 32768, 2048, 0, 160),         # S = 256
    ("glm-4-9b-1m", 3, 8192, 1024, 0, 160),           # G = 8, interleaved RoPE, batch 3
    ("yi-9b-200k", 2, 16384, 4096, 0, 160),           # S = 512, 4 kv x 8 q heads
    ("llama-3.1-8b", 1, 32768, 8192, 0, 160),         # S = 1024: 64 tile CTAs per row
    ("llama-3.1-8b", 2, 8232, 1024, 0, 160),          # ragged prompt: the sparse region starts at row 232, not a tile edge
    ("llama-3.1-8b", 2, 8192, 1024, 128, 160),        # HBM chunk store in front of the host store
    ("llama-3.1-8b", 2, 8192, 1024, 0, 64),           # rank 64: one slab
    ("llama-3.1-8b", 1, 8192, 1024, 0, 136),          # rank 136: half-valid last MMA step
])
def test_fused_launch_equals_fetch_rebuild_plus_attention(cuda, model, B, prompt_len, budget, ways, rank):
    shape, c1, _, cs = _build(cuda, model, B, prompt_len, budget, 2, rank=rank)
    _, c2, _, _ = _build(cuda, model, B, prompt_len, budget, 2, rank=rank)
    c1.fused_attend, c2.fused_attend = True, False
    assert c1.can_fuse_attend() and not c2.can_fuse_attend()
    if ways:
        c1.enable_hbm_store(ways)
        c2.enable_hbm_store(ways)
    d_cs = cs.to(cuda)
    base, worst = None, 0.0
    for step in range(6):
        for l in range(2):
            # step 3 repeats step 2's queries: every selected chunk is a hit, no tile CTA stays
            q, kn, vn = synth.queries(shape, B, seed=17 * (2 if step == 3 else step) + l, base=base, noise=0.4)
            base = q if base is None else base
            dq, dk, dv = q.to(cuda), kn.to(cuda), vn.to(cuda)
            o1 = c1.decode_layer(l, dq, dk, dv, d_cs)
            o2 = c2.decode_layer(l, dq, dk, dv, d_cs)
            torch.cuda.synchronize()
            assert torch.equal(c1.cnts, c2.cnts)
            assert torch.equal(c1.position_ids[l], c2.position_ids[l])
            assert torch.equal(c1.k_cache_buffer[l], c2.k_cache_buffer[l]), f"K bytes differ at step {step} layer {l}"
            assert torch.equal(c1.v_cache_buffer[l], c2.v_cache_buffer[l]), f"V bytes differ at step {step} layer {l}"
            err, _ = assert_close_bf16(o1, o2, max_abs=4e-3, what=f"fused vs separate attention, step {step} layer {l}")
            worst = max(worst, err)
            if step == 3:
                assert int(c1.cnts.min()) == budget // 8, "repeated queries must hit every slot"
    print(f"\n[{model} B={B} S={budget // 8} ways={ways} rank={rank}] fused vs separate attention: worst max-abs {worst:.2e}")


def test_fused_launch_all_misses_every_step(cuda):
    """The batch-1 GPU-resident class clears its chunk table every step: every slot is a miss, the split CTAs see only
    the local window, the outliers and the generated rows."""
    from shadowkv_b200.kv_cache import ShadowKVCache
    shape, c1, _, cs = _build(cuda, "llama-3.1-8b", 1, 16384, 2048, 1, cls=ShadowKVCache)
    _, c2, _, _ = _build(cuda, "llama-3.1-8b", 1, 16384, 2048, 1, cls=ShadowKVCache)
    c1.fused_attend, c2.fused_attend = True, False
    d_cs = cs.to(cuda)
    for step in range(3):
        q, kn, vn = synth.queries(shape, 1, seed=step)
        o1 = c1.decode_layer(0, q.to(cuda), kn.to(cuda), vn.to(cuda), d_cs)
        o2 = c2.decode_layer(0, q.to(cuda), kn.to(cuda), vn.to(cuda), d_cs)
        torch.cuda.synchronize()
        assert int(c1.cnts.sum()) == 0
        assert torch.equal(c1.k_cache_buffer[0], c2.k_cache_buffer[0]) and torch.equal(c1.v_cache_buffer[0], c2.v_cache_buffer[0])
        assert_close_bf16(o1, o2, max_abs=4e-3, what=f"all-miss step {step}")


@pytest.mark.parametrize("model,B", [("llama-3.1-8b", 2), ("glm-4-9b-1m", 1)])
def test_fused_launch_matches_oracle(cuda, model, B):
    prompt_len, budget, layers = 8192, 1024, 2
    shape, cache, states, cs = _build(cuda, model, B, prompt_len, budget, layers)
    assert cache.can_fuse_attend()
    lay = states[0]["layout"]
    ora = [O.LayerState(landmarks=s["k_landmark"], landmark_idx=s["k_landmark_idx"], U=s["U"], SV=s["SV"],
                        v_host=s["v_chunks"], k_buf=s["k_buf"].clone(), v_buf=s["v_buf"].clone(),
                        resident=s["resident"].clone()) for s in states]
    geo = O.Geometry(sparse_start=lay.sparse_start, sparse_end=lay.sparse_end)
    d_cs = cs.to(cuda)
    base = None
    H, S, ss = shape.num_key_value_heads, lay.select, lay.sparse_start
    for step in range(4):
        geo.gen_offset = step
        for l in range(layers):
            q, kn, vn = synth.queries(shape, B, seed=23 * step + l, base=base, noise=0.35)
            base = q if base is None else base
            out = cache.decode_layer(l, q.to(cuda), kn.to(cuda), vn.to(cuda), d_cs)
            torch.cuda.synchronize()
            res = cache.position_ids[l].cpu()
            want, info = O.decode_step(ora[l], geo, q, kn, vn, cs, selected_override=res)
            assert torch.equal(res.sort(-1).values, ora[l].resident.sort(-1).values)
            np.testing.assert_array_equal(cache.cnts.cpu().numpy(), info["cnts"])
            vb = cache.v_cache_buffer[l].cpu()
            for b in range(B):
                for h in range(H):
                    assert torch.equal(vb[b, h, ss:ss + S * 8].reshape(S, -1), states[l]["v_chunks"][b, h, res[b, h]])
            assert_close_bf16(out, want, what=f"attention vs oracle, step {step} layer {l}")
            # the oracle keeps the reference's slot order; carry the kernel's K rows over chunk by chunk
            kb = cache.k_cache_buffer[l].cpu()
            for b in range(B):
                for h in range(H):
                    pos = {int(c): i for i, c in enumerate(res[b, h].tolist())}
                    src = torch.tensor([pos[int(c)] for c in ora[l].resident[b, h].tolist()])
                    ora[l].k_buf[b, h, ss:ss + S * 8] = kb[b, h, ss:ss + S * 8].view(S, 8, -1)[src].reshape(S * 8, -1)


@pytest.mark.parametrize("model,B,prompt_len,budget", [("llama-3.1-8b", 1, 32768, 2048),     # one wave (296 CTAs)
                                                       ("llama-3.1-8b", 12, 8192, 2048),     # 96 rows: several waves
                                                       ("glm-4-9b-1m", 5, 8192, 1024)])
def test_row_flag_handover_equals_plain_dependent_launches(cuda, monkeypatch, model, B, prompt_len, budget):
    """skv_geom.row_flags: the CTAs of the fused fetch launch wait for their ROW's word from the selection kernel instead
    of for the completion of the selection grid.  Same kernels, same arithmetic: every output is bit-identical to the
    plain programmatic launches (SKV_ROW_FLAGS=0), over enough steps and rows that late-starting CTAs occur."""
    monkeypatch.setenv("SKV_ROW_FLAGS", "1")
    shape, c1, _, cs = _build(cuda, model, B, prompt_len, budget, 2)
    d_cs = cs.to(cuda)
    c1._geometry(d_cs)
    monkeypatch.setenv("SKV_ROW_FLAGS", "0")
    _, c2, _, _ = _build(cuda, model, B, prompt_len, budget, 2)
    c2._geometry(d_cs)
    assert c1._row_flags is not None and c2._row_flags is None
    base = None
    for step in range(8):
        for l in range(2):
            q, kn, vn = synth.queries(shape, B, seed=31 * step + l, base=base, noise=0.4)
            base = q if base is None else base
            dq, dk, dv = q.to(cuda), kn.to(cuda), vn.to(cuda)
            o1 = c1.decode_layer(l, dq, dk, dv, d_cs)
            o2 = c2.decode_layer(l, dq, dk, dv, d_cs)
        torch.cuda.synchronize()
        assert int(c1._row_flags.min()) == 1, "every row's word is raised once its table is written"
        for l in range(2):
            assert torch.equal(c1.position_ids[l], c2.position_ids[l]) and torch.equal(c1.cnts_all[l], c2.cnts_all[l])
            assert torch.equal(c1.k_cache_buffer[l], c2.k_cache_buffer[l]), f"K bytes differ at step {step} layer {l}"
            assert torch.equal(c1.v_cache_buffer[l], c2.v_cache_buffer[l]), f"V bytes differ at step {step} layer {l}"
        assert torch.equal(o1, o2), f"attention output differs at step {step}"
