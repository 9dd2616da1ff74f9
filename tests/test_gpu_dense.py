"""Dense remainder of the decode layer against oracle/dense_oracle.py: rmsnorm, the stream-K
tcgen05 linear with its four epilogues (lm_head f32, residual, SiLU*mul, QKV + rotary embedding), argmax, and the
reference-named layer methods of shadowkv_b200.llama.Llama.

Bars: element-wise stages on identical inputs <= 1 bf16 ulp (the fp32 summation order of the contraction differs from
the oracle's, so an accumulator can round to the neighbouring bf16 value); stages that feed a rounded value into
another rounding (residual add, SiLU*mul, RoPE) are bounded by the propagated 1-ulp input difference; logits-level
results: max-abs 2e-2, cosine > 0.999 (north_star)."""
import pytest
import torch

from oracle import dense_oracle as DO
from shadowkv_b200 import synth
from tests.tolerances import assert_close_bf16, assert_rope_rows_close, assert_ulp_bf16

pytestmark = pytest.mark.gpu
BF16 = torch.bfloat16


def _rn(g, *s, scale=1.0):
    return (torch.randn(*s, generator=g) * scale).to(BF16)


def _assert_prop(got, want, y_mag, *, ulps=1, frac_exact=0.95, what=""):
    """|got - want| <= ulps * 2^-7 * (|y| + |want|) + 1e-5: `y` is the rounded intermediate whose 1-ulp difference
    propagates into the result (the result itself may be smaller after cancellation)."""
    g, w = got.float().cpu(), want.float().cpu()
    tol = ulps * 2.0 ** -7 * (y_mag.float().cpu().abs() + w.abs()) + 1e-5
    bad = (g - w).abs() > tol
    assert not bool(bad.any()), f"{what}: {int(bad.sum())} of {bad.numel()} outside the propagated-ulp tolerance"
    exact = float((got.cpu().view(torch.int16) == want.cpu().view(torch.int16)).float().mean())
    assert exact >= frac_exact, f"{what}: only {exact:.4f} bit-identical"


@pytest.mark.parametrize("B,hidden", [(1, 4096), (3, 4096), (32, 4096), (2, 1024), (2, 8192), (1, 136)])
def test_rmsnorm(cuda, B, hidden):
    from shadowkv_b200 import dense
    g = torch.Generator().manual_seed(hidden + B)
    x, w = _rn(g, B, hidden, scale=3.0), (1.0 + 0.1 * torch.randn(hidden, generator=g)).to(BF16)
    got = dense.rmsnorm(x.to(cuda), w.to(cuda), 1e-5)
    assert_ulp_bf16(got, DO.rmsnorm(x, w, 1e-5), ulps=1, frac_exact=0.99, what="rmsnorm")


def test_embed_rmsnorm(cuda):
    from shadowkv_b200 import dense
    g = torch.Generator().manual_seed(3)
    emb, w = _rn(g, 1000, 4096, scale=0.02), (1.0 + 0.1 * torch.randn(4096, generator=g)).to(BF16)
    tok = torch.tensor([999, 0, 17, 17], dtype=torch.int64)
    x_out = torch.zeros(4, 4096, dtype=BF16, device=cuda)
    h_out = torch.zeros(4, 4096, dtype=BF16, device=cuda)
    dense.embed_rmsnorm(tok.to(cuda), emb.to(cuda), w.to(cuda), 1e-5, x_out, h_out)
    assert torch.equal(x_out.cpu(), emb[tok])
    assert_ulp_bf16(h_out, DO.rmsnorm(emb[tok], w, 1e-5), ulps=1, frac_exact=0.99, what="embed_rmsnorm")


@pytest.mark.parametrize("B,K,N,bias,resid", [
    (1, 4096, 4096, False, True),        # o_proj at batch 1: 32 row tiles split over 148 CTAs
    (5, 14336, 4096, False, True),       # down_proj
    (16, 4096, 4096, True, False),       # plain F.linear with bias
    (17, 1024, 200, True, True),         # N=32 accumulator, ragged last tile
    (32, 4096, 4096, False, True),
    (48, 4096, 4096, True, True),        # N=64 accumulator (the reference's batch 48 / 50 rows, test/e2e.py:56-60)
    (64, 14336, 4096, False, True),
    (2, 64, 128, False, False),          # one slab, one tile, one CTA
    (3, 128, 40000, False, False),       # many tiles per CTA, none split
])
def test_linear_residual(cuda, B, K, N, bias, resid):
    from shadowkv_b200 import dense
    g = torch.Generator().manual_seed(B * 7 + K + N)
    x, W = _rn(g, B, K), _rn(g, N, K, scale=0.02)
    b = _rn(g, N, scale=0.5) if bias else None
    r = _rn(g, B, N) if resid else None
    y = DO.linear(x, W, b)
    want = DO.add_bf16(r, y) if resid else y
    got = dense.linear(x.to(cuda), W.to(cuda), bias=b.to(cuda) if bias else None,
                       residual=r.to(cuda) if resid else None)
    _assert_prop(got, want, y, what=f"linear B={B} K={K} N={N}")
    # in-place residual stream
    if resid:
        rr = r.to(cuda).clone()
        dense.linear(x.to(cuda), W.to(cuda), bias=b.to(cuda) if bias else None, residual=rr, out=rr)
        assert torch.equal(rr, got)


@pytest.mark.parametrize("B,K,N,bias", [(1, 4096, 4096, False), (5, 14336, 4096, False), (17, 1024, 200, True),
                                        (32, 4096, 4096, True), (2, 64, 128, False), (3, 4096, 8192, False),
                                        (50, 4096, 4096, False)])
def test_partial_planes_plus_finish_equal_the_unfused_path_bit_for_bit(cuda, B, K, N, bias):
    """linear_partial + finish_rmsnorm (no cross-CTA fix-up inside the linear; the norm kernel adds the planes) against
    linear(residual=) + rmsnorm: same summation order, so the same bits -- and both against the oracle."""
    from shadowkv_b200 import dense
    g = torch.Generator().manual_seed(B + K + N)
    x, W = _rn(g, B, K), _rn(g, N, K, scale=0.02)
    b = _rn(g, N, scale=0.5) if bias else None
    r, w = _rn(g, B, N), (1.0 + 0.1 * torch.randn(N, generator=g)).to(BF16)
    xd, Wd, bd, wd = x.to(cuda), W.to(cuda), (b.to(cuda) if bias else None), w.to(cuda)
    x1 = dense.linear(xd, Wd, bias=bd, residual=r.to(cuda))
    h1 = dense.rmsnorm(x1, wd, 1e-5)
    x2 = r.to(cuda).clone()
    planes = dense.linear_partial(xd, Wd, static_weights=True)
    h2 = dense.finish_rmsnorm(planes, K, x2, wd, 1e-5, bias=bd)
    assert torch.equal(x1, x2) and torch.equal(h1, h2)
    y = DO.linear(x, W, b)
    want_x = DO.add_bf16(r, y)
    _assert_prop(x2, want_x, y, what="fused residual")
    # the norm of a 1-ulp-different x: compare on the kernel's own x (the stage itself), then loosely end to end
    assert_ulp_bf16(h2, DO.rmsnorm(x2.cpu(), w, 1e-5), ulps=1, frac_exact=0.99, what="fused rmsnorm")


def test_linear_is_deterministic_and_leaves_the_counters_at_zero(cuda):
    from shadowkv_b200 import dense
    g = torch.Generator().manual_seed(1)
    x, W = _rn(g, 4, 4096).to(cuda), _rn(g, 4096, 4096, scale=0.02).to(cuda)
    a = dense.linear(x, W)
    for _ in range(5):
        assert torch.equal(dense.linear(x, W), a)
    for _ in range(3):                                   # weights streamed ahead of the previous kernel's completion
        assert torch.equal(dense.linear(x, W, static_weights=True), a)
    torch.cuda.synchronize()
    ws = dense.workspace(cuda)
    assert int(ws[:16384].view(torch.int32).abs().sum()) == 0
    # another stream gets its own scratch
    with torch.cuda.stream(torch.cuda.Stream()):
        assert dense.workspace(cuda).data_ptr() != ws.data_ptr()
        assert torch.equal(dense.linear(x, W), a)


@pytest.mark.parametrize("B,vocab", [(1, 128256), (4, 64000), (40, 32000)])
def test_lm_head_f32_and_argmax(cuda, B, vocab):
    from shadowkv_b200 import dense
    g = torch.Generator().manual_seed(vocab)
    x, W = _rn(g, B, 4096), _rn(g, vocab, 4096, scale=0.02)
    want = DO.linear(x, W)
    got = dense.linear_f32(x.to(cuda), W.to(cuda))
    assert got.dtype == torch.float32
    assert torch.equal(got.to(BF16).float(), got), "logits must be bf16-representable (F.linear(...).float())"
    assert_ulp_bf16(got.to(BF16), want, ulps=1, frac_exact=0.97, what="lm_head")
    # every logit is within 1 bf16 ulp (checked above); the absolute bar is therefore one ulp of the largest logit where
    # that exceeds the 2e-2 of north_star (|logit| >= 4 -> ulp 2^-5)
    assert_close_bf16(got, want, max_abs=max(2e-2, 2.0 ** -7 * float(want.float().abs().max())), what="logits")
    tok = dense.argmax(got)
    assert torch.equal(tok.cpu(), DO.greedy(got.cpu()))
    # ties resolve to the lowest index
    t = torch.zeros(2, 5000, device=cuda)
    t[0, 4000] = t[0, 123] = 2.0
    t[1, :] = -1.0
    assert dense.argmax(t).tolist() == [123, 0]


@pytest.mark.parametrize("B,K,F", [(1, 4096, 14336), (7, 4096, 11008), (32, 1024, 1024), (16, 4096, 13696),
                                   (48, 4096, 14336)])
def test_linear_silu_mul(cuda, B, K, F):
    from shadowkv_b200 import dense
    g = torch.Generator().manual_seed(B + F)
    x, W = _rn(g, B, K), _rn(g, 2 * F, K, scale=0.03)
    gu = DO.linear(x, W)
    want = DO.silu_and_mul(gu)
    got = dense.linear_silu_mul(x.to(cuda), W.to(cuda))
    # gate and up each within 1 ulp -> silu(gate) within ~1.1 ulp (|d silu / silu| <= 1.1 |dg / g| for |g| <~ 4)
    # -> product within 3 ulp
    # typically; for negative gates |d ln silu / d ln g| = |1 + g (1 - sigmoid(g))| reaches ~2, hence the 6-ulp tail
    assert_ulp_bf16(got, want, ulps=6, frac_exact=0.93, abs_floor=1e-4, what="silu_mul")
    # entries reach |16| here (1 bf16 ulp = 0.0625), so the absolute bar scales with the magnitude
    assert_close_bf16(got, want, max_abs=3 * 2.0 ** -8 * float(want.float().abs().max()), what="silu_mul")


@pytest.mark.parametrize("model,B,use_dev_pos", [("llama-3.1-8b", 1, False), ("llama-3.1-8b", 4, True),
                                                 ("yi-9b-200k", 2, False), ("glm-4-9b-1m", 3, True),
                                                 ("glm-4-9b-1m", 20, False), ("llama-3.1-8b", 48, True)])
def test_qkv_rope(cuda, model, B, use_dev_pos):
    from shadowkv_b200 import dense
    sh = synth.MODELS[model]
    Hq, H, hid = sh.num_attention_heads, sh.num_key_value_heads, sh.hidden_size
    g = torch.Generator().manual_seed(B + Hq + H)
    x, w = _rn(g, B, hid, scale=2.0), (1.0 + 0.1 * torch.randn(hid, generator=g)).to(BF16)
    Wqkv = _rn(g, (Hq + 2 * H) * 128, hid, scale=0.03)
    bias = _rn(g, (Hq + 2 * H) * 128, scale=0.5) if sh.glm_rope else None
    cs = synth.cos_sin_table(9000, theta=sh.rope_theta, glm=sh.glm_rope)
    pos = 8191
    wq, wk, wv = DO.pre_attention(x, w, 1e-5, Wqkv, bias, Hq, H, cs, pos, 2 if sh.glm_rope else 1)
    h = dense.rmsnorm(x.to(cuda), w.to(cuda), 1e-5)
    pos_dev = torch.tensor([91], dtype=torch.int32, device=cuda) if use_dev_pos else None
    q, k, v = dense.linear_qkv_rope(h, Wqkv.to(cuda), Hq, H, cs.to(cuda), pos - (91 if use_dev_pos else 0),
                                    bias=bias.to(cuda) if bias is not None else None, glm=sh.glm_rope, pos_dev=pos_dev)
    assert q.shape == (B, Hq, 1, 128) and k.shape == (B, H, 1, 128) and v.shape == (B, H, 1, 128)
    assert_rope_rows_close(q.view(B, Hq, 128), wq, glm=sh.glm_rope, ulps_in=2, frac_exact=0.90, what="q")
    assert_rope_rows_close(k.view(B, H, 128), wk, glm=sh.glm_rope, ulps_in=2, frac_exact=0.90, what="k")
    # v is a plain linear of the normed input.  Stage-wise (the oracle's linear on the KERNEL's own h): 1 ulp.  End to end
    # the ~1 % of h elements where the kernel's rmsnorm is 1 ulp off the oracle's reach v as a sum of ~40 terms of
    # ulp(h) * |W| ~ 8e-3 * 0.03, i.e. ~1.5e-3 absolutely, which only shows on the few outputs that cancel to nearly zero
    # (2-6 of 49 152 at batch 48): 2 ulps with a 5e-3 floor there.
    v_rows = slice((Hq + H) * 128, (Hq + 2 * H) * 128)
    wv_stage = DO.linear(h.cpu(), Wqkv[v_rows], bias[v_rows] if bias is not None else None)
    assert_ulp_bf16(v.view(B, H * 128), wv_stage, ulps=1, frac_exact=0.97, what="v (stage)")
    assert_ulp_bf16(v.view(B, H, 128), wv, ulps=2, frac_exact=0.95, abs_floor=5e-3, what="v")
    # without a table: the plain split of pre_attention_compute (llama.py:380)
    q0, k0, v0 = dense.linear_qkv_rope(h, Wqkv.to(cuda), Hq, H, None, 0, bias=bias.to(cuda) if bias is not None else None)
    n0 = DO.pre_attention(x, w, 1e-5, Wqkv, bias, Hq, H, None, 0, 0)
    assert_ulp_bf16(q0.view(B, Hq, 128), n0[0], ulps=2, frac_exact=0.95, abs_floor=5e-3, what="q (no rope)")   # as for v
    assert torch.equal(v0, v)


@pytest.mark.parametrize("model,B", [("llama-3.1-8b", 2), ("glm-4-9b-1m", 3)])
def test_qkv_head_ranges_and_in_place_append(cuda, model, B):
    """The q heads and the k / v heads projected by two launches (row slices of the fused wqkv), k / v appended in
    place at row `row + *pos_dev` of strided [B,H,rows,128] cache buffers: same values as the oracle's fused call."""
    from shadowkv_b200 import dense
    sh = synth.MODELS[model]
    Hq, H, hid = sh.num_attention_heads, sh.num_key_value_heads, sh.hidden_size
    g = torch.Generator().manual_seed(B + H)
    h = _rn(g, B, hid)
    Wqkv = _rn(g, (Hq + 2 * H) * 128, hid, scale=0.03)
    bias = _rn(g, (Hq + 2 * H) * 128, scale=0.5) if sh.glm_rope else None
    cs = synth.cos_sin_table(9000, theta=sh.rope_theta, glm=sh.glm_rope)
    pos, gen, row0, rows = 8100, 37, 2496, 2592
    qkv = DO.linear(h, Wqkv, bias)
    wq, wk, wv = qkv.split([Hq * 128, H * 128, H * 128], dim=-1)
    wq = DO.rotary_vllm(wq.reshape(B, Hq, 128), cs[pos + gen], neox=not sh.glm_rope)
    wk = DO.rotary_vllm(wk.reshape(B, H, 128), cs[pos + gen], neox=not sh.glm_rope)
    wv = wv.reshape(B, H, 128)
    dev_h, dev_W, dev_cs = h.to(cuda), Wqkv.to(cuda), cs.to(cuda)
    dev_b = bias.to(cuda) if bias is not None else None
    pos_dev = torch.tensor([gen], dtype=torch.int32, device=cuda)
    q, k0, v0 = dense.linear_qkv_rope(dev_h, dev_W[:Hq * 128], Hq, H, dev_cs, pos, bias=dev_b, glm=sh.glm_rope,
                                      pos_dev=pos_dev, heads=(0, Hq))
    assert k0 is None and v0 is None
    assert_rope_rows_close(q.view(B, Hq, 128), wq, glm=sh.glm_rope, ulps_in=1, frac_exact=0.93, what="q heads")
    kbuf = torch.full((B, H, rows, 128), 7.0, dtype=BF16, device=cuda)
    vbuf = torch.full((B, H, rows, 128), 7.0, dtype=BF16, device=cuda)
    dense.linear_qkv_rope(dev_h, dev_W[Hq * 128:], Hq, H, dev_cs, pos, bias=dev_b, glm=sh.glm_rope, pos_dev=pos_dev,
                          heads=(Hq, 2 * H), kv_cache=(kbuf, vbuf, row0))
    assert_rope_rows_close(kbuf[:, :, row0 + gen], wk, glm=sh.glm_rope, ulps_in=1, frac_exact=0.93, what="k row")
    assert_ulp_bf16(vbuf[:, :, row0 + gen], wv, ulps=1, frac_exact=0.97, what="v row")
    kbuf[:, :, row0 + gen] = 7.0
    vbuf[:, :, row0 + gen] = 7.0
    assert bool((kbuf == 7.0).all()) and bool((vbuf == 7.0).all()), "only the appended row may be written"


def test_argument_errors(cuda):
    from shadowkv_b200 import dense
    x = torch.zeros(65, 4096, dtype=BF16, device=cuda)
    W = torch.zeros(256, 4096, dtype=BF16, device=cuda)
    with pytest.raises(ValueError):
        dense.linear(x, W)                                   # batch > 64
    with pytest.raises(ValueError):
        dense.linear(x[:2, :100].contiguous(), W[:, :100].contiguous())      # in_features % 64
    with pytest.raises(TypeError):
        dense.linear(x[:2].float(), W)
    with pytest.raises(ValueError):
        dense.linear(x[:2], W[:, :128].contiguous())         # shape mismatch


@pytest.mark.parametrize("model", ["llama-3.1-8b", "glm-4-9b-1m"])
def test_layer_methods_against_oracle(cuda, model):
    """pre_attention_compute / post_attention_compute of the host mirror at the model's real layer shapes."""
    from shadowkv_b200.llama import Llama
    B = 2
    m = Llama(model, batch_size=B, max_length=8192, device="cuda:0", num_layers=1, seed=5)
    sh, lay = m.shape, m.layers[0]
    g = torch.Generator().manual_seed(11)
    x = _rn(g, B, sh.hidden_size, scale=1.5)
    pos = 8000
    q, k, v = m.pre_attention_compute(x.to(cuda), lay, m.num_heads, m.num_key_value_heads, 128, position=pos)
    cpu = lambda t: None if t is None else t.cpu()   # noqa: E731
    wq, wk, wv = DO.pre_attention(x, cpu(lay.input_layernorm_weight), 1e-5, cpu(lay.wqkv), cpu(lay.bqkv), m.num_heads,
                                  m.num_key_value_heads, cpu(m.cos_sin_cache), pos, 2 if sh.glm_rope else 1)
    assert_rope_rows_close(q.view(B, -1, 128), wq, glm=sh.glm_rope, ulps_in=2, frac_exact=0.90, what="q")
    assert_rope_rows_close(k.view(B, -1, 128), wk, glm=sh.glm_rope, ulps_in=2, frac_exact=0.90, what="k")
    attn = _rn(g, B, 1, m.num_heads, 128, scale=0.3)
    got = m.post_attention_compute(attn.to(cuda), x.to(cuda), lay)
    want = DO.post_attention(attn.reshape(B, -1), x, cpu(lay.wo), cpu(lay.post_attention_layernorm_weight), 1e-5,
                             cpu(lay.gate_up_proj), cpu(lay.down_proj))
    # hidden states reach |8| here (1 bf16 ulp = 0.0625): the absolute bar scales with the magnitude
    assert_close_bf16(got, want, max_abs=max(2e-2, 2.0 ** -6 * float(want.float().abs().max())),
                      what="post_attention_compute")
    exact = float((got.cpu().view(torch.int16) == want.view(torch.int16)).float().mean())
    assert exact > 0.75, exact


def _build_model(layers, B, ctx, hbm_store_ways=0, split_qkv=True):
    from shadowkv_b200.llama import Llama
    sh = synth.MODELS["llama-3.1-8b"]
    m = Llama("llama-3.1-8b", batch_size=B, max_length=ctx, device="cuda:0", num_layers=layers, seed=7,
              hbm_store_ways=hbm_store_ways, split_qkv=split_qkv)
    states = [synth.layer_state(sh, B, ctx, seed=100 + l, device="cuda:0", cos_sin=m.cos_sin_cache) for l in range(layers)]
    synth.install(m.kv_cache, states)
    torch.cuda.synchronize()
    return m


def test_token_runner_matches_the_reference_named_call_chain(cuda):
    """One layer of the Llama-3.1-8B shape at a short context: the whole-token runner (device-side counters, norms
    fused into the previous step, in-place residual stream) against the eager reference-named chain
    inference -> layer_compute -> pre/post_attention_compute, token after token on the same token ids.  (One layer:
    the two paths split the attention differently, and with more layers a last-bit difference can flip a chunk of the
    next layer's top-S selection, which is a discontinuity, not an error.  For the same reason the runner projects
    q, k, v in ONE launch here, like the eager chain: the two-launch split of the default runner sums the same products
    in another order, and a last-bit difference in q can flip a borderline chunk of this layer's own selection; the
    split launches are checked against the oracle in test_qkv_head_ranges_and_in_place_append and, end to end, in the
    graph-replay test below.)"""
    B, ctx = 2, 8192
    a, b = _build_model(1, B, ctx), _build_model(1, B, ctx, split_qkv=False)
    tok = torch.tensor([11, 4242], dtype=torch.int64, device=cuda)
    for t in range(3):
        pos = torch.full((B, 1), ctx + t, dtype=torch.int64, device=cuda)
        want = a.inference(tok.view(B, 1), pos)[:, 0]
        b.stream.wait_stream(torch.cuda.current_stream())        # tok was produced on the default stream
        with torch.cuda.stream(b.stream):
            b.tok.copy_(tok)
        b.decode_token()
        b.stream.synchronize()
        assert_close_bf16(b.logits, want, max_abs=max(2e-2, 2.0 ** -6 * float(want.abs().max())), what=f"token {t} logits")
        tok = want.argmax(dim=-1)
    assert a.kv_cache.gen_offset == b.kv_cache.gen_offset == 3


def test_whole_token_graph_replay_is_bit_identical_to_eager_launches(cuda):
    """Two layers: the captured CUDA graph (programmatic dependent launches, side-stream fork/join inside) replays to
    exactly the logits and tokens of the same kernels launched eagerly, and the greedy loop runs."""
    B, ctx = 2, 8192
    first = torch.tensor([11, 4242], dtype=torch.int64, device=cuda)
    a = _build_model(2, B, ctx)
    with torch.cuda.stream(a.stream):
        a.tok.copy_(first)
    want_logits, want_tok = [], []
    for t in range(3):
        a.decode_token()                         # no graph captured: eager launches
        a.stream.synchronize()
        want_logits.append(a.logits.clone())
        want_tok.append(a.tok.clone())
    b = _build_model(2, B, ctx)
    with torch.cuda.stream(b.stream):
        b.tok.copy_(first)
    b.capture(warmup=1)                          # token 0 eagerly, then the graph is captured
    for t in range(3):
        if t > 0:
            b.decode_token()
        b.stream.synchronize()
        assert torch.equal(b.logits, want_logits[t]), f"token {t} logits"
        assert torch.equal(b.tok, want_tok[t]), f"token {t}"
    assert b.kv_cache.gen_offset == 3
    c = _build_model(2, B, ctx)
    ids = c.generate(first, 3)
    assert ids.shape == (B, 3) and torch.equal(ids[:, 2], want_tok[2])
    # the reference's batch_generate loop (base.py:626-655): greedy = the same tokens; with a temperature the graph
    # ends at the logits and sample_token draws from the top-k set
    d = _build_model(2, B, ctx)
    ids_g, tps = d.batch_generate(first, gen_len=3, temperature=0.0, benchmark=True)
    assert torch.equal(ids_g, torch.stack(want_tok, dim=1)) and tps > 0
    e = _build_model(2, B, ctx)
    torch.manual_seed(0)
    ids_s = e.batch_generate(first, gen_len=2, temperature=0.8, top_k=5, top_p=-1)
    assert ids_s.shape == (B, 2)
    kth = want_logits[0].topk(5, dim=-1).values[:, -1]
    assert bool((want_logits[0].gather(1, ids_s[:, :1]).squeeze(1) >= kth).all()), "sampled token outside the top-k set"


def test_epilogues_against_library_golden(cuda):
    """The element-wise epilogues isolated with an identity weight matrix (x W^T = x exactly), against what the
    libraries themselves produced on a B200 (tests/golden/dense_golden.npz: vllm rotary_embedding / silu_and_mul,
    flashinfer rmsnorm, torch's bf16 add): rotary embedding and residual add bit for bit."""
    import os
    import numpy as np
    from shadowkv_b200 import dense
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dense_golden.npz"))
    bf = lambda a: torch.from_numpy(np.ascontiguousarray(a)).view(BF16)   # noqa: E731
    # rmsnorm
    got = dense.rmsnorm(bf(z["x"]).to(cuda), bf(z["norm_w"]).to(cuda), 1e-5)
    assert_ulp_bf16(got, bf(z["rmsnorm"]), ulps=1, frac_exact=0.995, what="rmsnorm vs flashinfer")
    # rotary embedding: 4 q heads + 2 k heads + 2 v heads = 1024 = in_features, W = I
    q, k = bf(z["rope_q"]).view(3, 512), bf(z["rope_k"]).view(3, 256)
    x = torch.cat([q, k, k], dim=1).contiguous().to(cuda)
    eye = torch.eye(1024, dtype=BF16, device=cuda)
    pos = int(z["rope_pos"])
    for name, glm in (("neox", False), ("glm", True)):
        rows = 128 if not glm else 64
        cs = torch.zeros(pos + 1, rows, dtype=BF16)
        cs[pos] = bf(z[f"cs_{name}_row"])
        qo, ko, vo = dense.linear_qkv_rope(x, eye, 4, 2, cs.to(cuda), pos, glm=glm)
        assert torch.equal(qo.view(3, 512).cpu(), bf(z[f"rope_q_{name}"]).view(3, 512)), name
        assert torch.equal(ko.view(3, 256).cpu(), bf(z[f"rope_k_{name}"]).view(3, 256)), name
        assert torch.equal(vo.view(3, 256).cpu(), k)
    # SiLU * mul: gate/up = the fixture's input through W = I (3072 x 3072)
    gu = bf(z["silu_in"]).to(cuda)
    eye = torch.eye(gu.shape[1], dtype=BF16, device=cuda)
    got = dense.linear_silu_mul(gu, eye)
    assert_ulp_bf16(got, bf(z["silu_and_mul"]), ulps=1, frac_exact=0.999, what="silu_and_mul vs vllm")
    # residual add
    a, b = bf(z["add_a"]).to(cuda), bf(z["add_b"]).to(cuda)
    eye = torch.eye(a.shape[1], dtype=BF16, device=cuda)
    assert torch.equal(dense.linear(a, eye, residual=b).cpu(), bf(z["add"]))
