#!/usr/bin/env python
"""Context number, NOT the headline: decode tokens/s of the WHOLE Llama-3.1-8B layer stack at 122K context --
the sparse-KV decode step of this repo plus the dense remainder of every layer (rmsnorm, fused QKV, o_proj,
gate/up, SiLU*mul, down; final norm + lm_head) with random-init bf16 weights through library GEMMs (torch / cuBLAS),
all captured in one CUDA graph.  SURVEY.md section 8(f2) lists the dense part as "next"; it is outside the path
this repo re-implements and is timed here only to put the hot-path number in proportion.

The dense kernels run with the right shapes and their q/k/v outputs are computed, but the hot path consumes the
synthetic q/k/v of bench.py (q_t = sqrt(1-n^2) base + n eps_t), because random weights have no temporal locality
and the chunk-cache hit rate (~60 % in the reference's own test) is an input of the workload.

    python tools/bench_full_layer.py [--batch 1] [--steps 64] > gpurun_out/full_layer.json
"""
import argparse
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from shadowkv_b200 import synth  # noqa: E402
from shadowkv_b200.kv_cache import ShadowKVCache_CPU  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--ctx", type=int, default=124928)
    ap.add_argument("--noise", type=float, default=0.28)
    ap.add_argument("--hbm-store-ways", type=int, default=0)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    shape = synth.MODELS["llama-3.1-8b"]
    L, B, Hq, H, D, hid, ffn, vocab = shape.num_hidden_layers, args.batch, 32, 8, 128, 4096, 14336, shape.vocab_size
    bf = torch.bfloat16
    cs = synth.cos_sin_table(args.ctx + 1024, theta=shape.rope_theta, device=dev)
    cache = ShadowKVCache_CPU(shape.config(L), batch_size=B, max_length=args.ctx, device="cuda:0",
                              hbm_store_ways=args.hbm_store_ways)
    states = []
    for l in range(L):
        st = synth.layer_state(shape, B, args.ctx, seed=1000 + l, device=dev, cos_sin=cs)
        cache.v_cache_cpu[l][:, :, :st["v_chunks"].shape[2]].copy_(st["v_chunks"])
        st["v_chunks"] = None
        states.append(st)
    synth.install(cache, states, skip_v_store=True)
    del states
    g = torch.Generator(device=dev).manual_seed(0)

    def w(*s):
        return (torch.randn(*s, generator=g, device=dev) * 0.02).to(bf)

    Wqkv = [w((Hq + 2 * H) * D, hid) for _ in range(L)]
    Wo = [w(hid, Hq * D) for _ in range(L)]
    Wgu = [w(2 * ffn, hid) for _ in range(L)]
    Wdn = [w(hid, ffn) for _ in range(L)]
    n1 = [torch.ones(hid, dtype=bf, device=dev) for _ in range(L)]
    n2 = [torch.ones(hid, dtype=bf, device=dev) for _ in range(L)]
    nf = torch.ones(hid, dtype=bf, device=dev)
    lm_head = w(vocab, hid)
    emb = w(vocab, hid)
    weight_bytes = sum(t.numel() * 2 for ws in (Wqkv, Wo, Wgu, Wdn) for t in ws) + lm_head.numel() * 2

    total = args.warmup + args.steps
    a = math.sqrt(1 - args.noise ** 2)
    base = torch.randn(L, B, Hq, 1, D, generator=g, device=dev)
    qs = [(a * base + args.noise * torch.randn(L, B, Hq, 1, D, generator=g, device=dev)).to(bf) for _ in range(total)]
    kv = [torch.randn(L, B, H, 1, D, generator=g, device=dev).to(bf) for _ in range(total)]
    q_in, k_in = torch.zeros_like(qs[0]), torch.zeros_like(kv[0])
    tok = torch.zeros(B, dtype=torch.long, device=dev)
    attn_out = torch.zeros(L, B, 1, Hq, D, dtype=bf, device=dev)
    sink = torch.zeros(B, (Hq + 2 * H) * D, dtype=bf, device=dev)
    stream = torch.cuda.Stream()

    def token(dense: bool):
        x = emb[tok] if dense else None
        for l in range(L):
            if dense:
                h = F.rms_norm(x, (hid,), n1[l], 1e-5)
                qkv = h @ Wqkv[l].t()                       # the real model's q/k/v (pre-RoPE); computed, see docstring
                sink.copy_(qkv)
            o = cache.decode_layer(l, q_in[l], k_in[l], k_in[l], cs, out=attn_out[l], device_counters=True)
            if dense:
                x = x + o.view(B, Hq * D) @ Wo[l].t()
                h2 = F.rms_norm(x, (hid,), n2[l], 1e-5)
                gu = h2 @ Wgu[l].t()
                x = x + (F.silu(gu[:, :ffn]) * gu[:, ffn:]) @ Wdn[l].t()
        cache.join_side_streams()
        cache.advance_device_counters()
        if dense:
            logits = F.rms_norm(x, (hid,), nf, 1e-5) @ lm_head.t()
            tok.copy_(logits.argmax(dim=-1))

    def run(dense: bool):
        cache.gen_offset = 0
        cache.reset_device_counters()
        with torch.cuda.stream(stream):
            q_in.copy_(qs[0]); k_in.copy_(kv[0]); token(dense)
        stream.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            token(dense)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cap = cache.buf_len - cache.sparse_end - 2
        with torch.cuda.stream(stream):
            for t in range(total):
                if t == args.warmup:
                    e0.record()
                if t % cap == cap - 1:
                    cache._counters_dev.copy_(torch.tensor([0, cache.sparse_end + 1], dtype=torch.int32, device=dev))
                q_in.copy_(qs[t]); k_in.copy_(kv[t]); graph.replay()
            e1.record()
        stream.synchronize()
        return e0.elapsed_time(e1) / args.steps

    ms_hot = run(False)
    ms_full = run(True)
    hits = float(cache.cnts_all.float().mean()) / cache.select_sets
    print(json.dumps({
        "what": "Llama-3.1-8B decode, 122K context: sparse-KV hot path alone vs whole layer stack incl. dense GEMMs (cuBLAS) + lm_head",
        "batch": B, "hot_path_ms_per_token": ms_hot, "full_model_ms_per_token": ms_full,
        "hot_path_tokens_per_s": B / ms_hot * 1e3, "full_model_tokens_per_s": B / ms_full * 1e3,
        "dense_ms_per_token": ms_full - ms_hot, "dense_weight_GB": weight_bytes / 1e9,
        "dense_weight_stream_GBps": weight_bytes / ((ms_full - ms_hot) * 1e-3) / 1e9,
        "chunk_cache_hit_rate_last_token": round(hits, 3), "hbm_store_ways": args.hbm_store_ways}))


if __name__ == "__main__":
    main()
