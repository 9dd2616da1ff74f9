"""DecodeStepGraph (all layers of a token in one CUDA graph, device-side generation counters) against the
eager method-by-method path on identical state: same chunk table, same K/V buffer bytes, same attention
output (the split-K boundaries differ, so the output is compared within the bf16 tolerance)."""
import pytest
import torch

from shadowkv_b200 import synth
from tests.tolerances import assert_close_bf16

pytestmark = pytest.mark.gpu


def _cache(cuda, shape, B, prompt_len, budget, layers, cs):
    from shadowkv_b200.kv_cache import ShadowKVCache_CPU
    states = [synth.layer_state(shape, B, prompt_len, seed=1000 + l, sparse_budget=budget, cos_sin=cs)
              for l in range(layers)]
    cache = ShadowKVCache_CPU(shape.config(layers), batch_size=B, max_length=prompt_len, device=str(cuda),
                              sparse_budget=budget)
    return synth.install(cache, states)


@pytest.mark.parametrize("host_io", [False, True])
def test_graph_replay_equals_eager(cuda, host_io):
    from shadowkv_b200.engine import DecodeStepGraph
    shape = synth.MODELS["llama-3.1-8b"]
    B, prompt_len, budget, layers, tokens = 2, 8192, 1024, 3, 5
    cs = synth.cos_sin_table(prompt_len + 1024)
    d_cs = cs.to(cuda)
    c1 = _cache(cuda, shape, B, prompt_len, budget, layers, cs)
    c2 = _cache(cuda, shape, B, prompt_len, budget, layers, cs)
    Hq, H, D = shape.num_attention_heads, shape.num_key_value_heads, 128
    g = torch.Generator().manual_seed(3)
    base = torch.randn(layers, B, Hq, 1, D, generator=g)

    def tok_inputs():
        q = (0.95 * base + 0.3 * torch.randn(layers, B, Hq, 1, D, generator=g)).to(torch.bfloat16)
        k = torch.randn(layers, B, H, 1, D, generator=g).to(torch.bfloat16)
        v = torch.randn(layers, B, H, 1, D, generator=g).to(torch.bfloat16)
        return q, k, v

    ins = [tok_inputs() for _ in range(tokens)]
    run = DecodeStepGraph(c2, d_cs, host_io=host_io)
    # the capture warm-up consumes token 0 (a real decode step)
    if host_io:
        run.q_h.copy_(ins[0][0]); run.k_h.copy_(ins[0][1]); run.v_h.copy_(ins[0][2])
    else:
        run.q_d.copy_(ins[0][0]); run.k_d.copy_(ins[0][1]); run.v_d.copy_(ins[0][2])
    run.capture(warmup=1)
    assert c2.gen_offset == 1
    for t in range(tokens):
        q, k, v = ins[t]
        outs = [c1.decode_layer(l, q[l].to(cuda), k[l].to(cuda), v[l].to(cuda), d_cs) for l in range(layers)]
        torch.cuda.synchronize()
        if t > 0:
            if host_io and t % 2:            # pinned inputs are copied to the device as they are ...
                got = run.step_host(q.pin_memory(), k.pin_memory(), v.pin_memory()).to(cuda)
            elif host_io:                    # ... pageable ones go through the runner's pinned stage
                got = run.step_host(q, k, v).to(cuda)
            else:
                got = run.step_device(q.to(cuda), k.to(cuda), v.to(cuda))
        else:
            got = run.out_d
        torch.cuda.synchronize()
        assert c1.gen_offset == c2.gen_offset == t + 1
        assert torch.equal(c1.position_ids, c2.position_ids), f"chunk table differs at token {t}"
        assert torch.equal(c1.k_cache_buffer, c2.k_cache_buffer), f"K buffers differ at token {t}"
        assert torch.equal(c1.v_cache_buffer, c2.v_cache_buffer), f"V buffers differ at token {t}"
        for l in range(layers):
            assert_close_bf16(got[l], outs[l], max_abs=4e-3, what=f"token {t} layer {l}")
    assert int(run.hits_d.sum()) > 0


def test_rewind_reuses_generation_region(cuda):
    from shadowkv_b200.engine import DecodeStepGraph
    shape = synth.MODELS["llama-3.1-8b"]
    B, prompt_len, budget, layers = 1, 8192, 1024, 1
    cs = synth.cos_sin_table(prompt_len + 1024)
    c = _cache(cuda, shape, B, prompt_len, budget, layers, cs)
    run = DecodeStepGraph(c, cs.to(cuda), host_io=False).capture(warmup=1)
    cap = c.buf_len - c.sparse_end
    q = torch.randn(layers, B, 32, 1, 128, device=cuda).to(torch.bfloat16)
    k = torch.randn(layers, B, 8, 1, 128, device=cuda).to(torch.bfloat16)
    for _ in range(cap - 1):
        run.step_device(q, k, k)
    assert c.gen_offset == cap
    with pytest.raises(RuntimeError):
        run.replay()
    run.rewind_async(0)
    k2 = torch.full_like(k, 3.0)
    run.step_device(q, k2, k2)
    torch.cuda.synchronize()
    assert c.gen_offset == 1
    assert torch.equal(c.k_cache_buffer[0][:, :, c.sparse_end], k2[0][:, :, 0])


def test_peer_store_exchange_single_rank(cuda):
    """The kv-head split's exchange path with one rank (peer = self; no process group): the merging CTAs of the fused
    launch also store every row into the gathered buffer and the rank's last row raises the layer's flag with the token's
    epoch; skv_wait_flags is in the captured chain.  Same kernel path the N-GPU runs take over NVLink."""
    from shadowkv_b200.engine import DecodeStepGraph
    shape = synth.MODELS["llama-3.1-8b"]
    B, prompt_len, budget, layers = 2, 8192, 1024, 3
    cs = synth.cos_sin_table(prompt_len + 1024)
    c = _cache(cuda, shape, B, prompt_len, budget, layers, cs)
    run = DecodeStepGraph(c, cs.to(cuda), host_io=False, gather_heads=True).capture(warmup=1)
    assert run.xchg is not None and run.xchg.world == 1 and "P2P" in run.exchange
    g = torch.Generator().manual_seed(5)
    for t in range(4):
        q = torch.randn(layers, B, 32, 1, 128, generator=g).to(torch.bfloat16).to(cuda)
        k = torch.randn(layers, B, 8, 1, 128, generator=g).to(torch.bfloat16).to(cuda)
        out = run.step_device(q, k, k)
        run.stream.synchronize()
        assert torch.equal(run.out_all[:, 0], out), f"gathered buffer differs from the local output at token {t}"
        # warm-up token used epoch 1; this replay is epoch t + 2, and the device counter already points at the next one
        assert int(run.xchg.epoch) == t + 3
        assert bool((run.xchg.flags == t + 2).all()), run.xchg.flags


@pytest.mark.parametrize("B,layers", [(1, 6), (4, 3)])
def test_output_row_handover_equals_plain_dependent_launches(cuda, monkeypatch, B, layers):
    """One-GPU hand-over of a layer's attention output into the next layer's scoring launch (parallel.RowHandover,
    SKV_GEOM_WAIT_FLAGS_ONLY): the scoring launch starts on the previous fused launch's row words and waits for that
    launch's completion at its end.  Chunk tables, K / V buffers and outputs must be bit-identical to the runner with
    plain dependent launches (SKV_OUT_FLAGS=0) over many tokens, and every row word must carry the token's epoch."""
    from shadowkv_b200.engine import DecodeStepGraph
    shape = synth.MODELS["llama-3.1-8b"]
    prompt_len, budget, tokens = 8192, 1024, 24
    cs = synth.cos_sin_table(prompt_len + 1024)
    d_cs = cs.to(cuda)
    runs = []
    for on in ("1", "0"):
        monkeypatch.setenv("SKV_OUT_FLAGS", on)
        c = _cache(cuda, shape, B, prompt_len, budget, layers, cs)
        runs.append(DecodeStepGraph(c, d_cs, host_io=False).capture(warmup=1))
    assert runs[0].handover is not None and runs[1].handover is None
    g = torch.Generator().manual_seed(11)
    base = torch.randn(layers, B, 32, 1, 128, generator=g)
    for t in range(tokens):
        q = (0.95 * base + 0.3 * torch.randn(layers, B, 32, 1, 128, generator=g)).to(torch.bfloat16).to(cuda)
        k = torch.randn(layers, B, 8, 1, 128, generator=g).to(torch.bfloat16).to(cuda)
        v = torch.randn(layers, B, 8, 1, 128, generator=g).to(torch.bfloat16).to(cuda)
        outs = [r.step_device(q, k, v) for r in runs]
        torch.cuda.synchronize()
        assert torch.equal(outs[0], outs[1]), f"attention outputs differ at token {t}"
        assert torch.equal(runs[0].cache.position_ids, runs[1].cache.position_ids), f"chunk table differs at token {t}"
        assert bool((runs[0].handover.flags == t + 2).all()), runs[0].handover.flags
    assert torch.equal(runs[0].cache.k_cache_buffer, runs[1].cache.k_cache_buffer)
    assert torch.equal(runs[0].cache.v_cache_buffer, runs[1].cache.v_cache_buffer)
