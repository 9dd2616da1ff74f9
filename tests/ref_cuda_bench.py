#!/usr/bin/env python
"""Times the REFERENCE's own CUDA path on this B200 (the bar BASELINE.json's north_star names): the unmodified
reference extension built by oracle/build_ref.sh (oracle/_ref, CUTLASS 2.x Sm80 kernels recompiled for sm_100a)
plus flash_attn_with_kvcache from the flash-attn wheel, driven with exactly the call sequence of the reference's
ShadowKVCache_CPU decode methods (models/kv_cache.py:983-1176) and base.py:320-341's stream choreography, on the
same synthetic decode state and queries bench.py uses.

Checker-side tool (it lives under tests/ because it executes oracle/_ref): not part of the product path, not
collected by pytest, and bench.py does not call it.  Output: one JSON line
(tokens/s of the hot path for all layers, us per layer-step, per-stage CUDA-event times).

    python tests/ref_cuda_bench.py [--steps 48] [--warmup 8] [--batch 1] > gpurun_out/ref_cuda.json
"""
import argparse
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))   # tests/ -> repo root
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref", "kernels"))

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=48)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--ctx", type=int, default=124928)
    ap.add_argument("--model", default="llama-3.1-8b")
    ap.add_argument("--layers", type=int, default=0)
    ap.add_argument("--noise", type=float, default=0.28)
    args = ap.parse_args()

    import shadowkv as ref                         # the reference extension (oracle/_ref)
    from flash_attn import flash_attn_with_kvcache   # what base.py:38 imports
    from shadowkv_b200 import synth                # synthetic state only (no kernels of ours are launched)

    dev = torch.device("cuda:0")
    shape = synth.MODELS[args.model]
    L = args.layers or shape.num_hidden_layers
    B, H, Hq, D, C_ = args.batch, shape.num_key_value_heads, shape.num_attention_heads, 128, 8
    G = Hq // H
    lay = synth.layout(args.ctx, 2048)
    S, N, R = lay.select, lay.n_landmarks, 160
    budget = 2048
    cs = synth.cos_sin_table(args.ctx + 1024, theta=shape.rope_theta, glm=shape.glm_rope, device=dev)

    # ---- state, laid out as ShadowKVCache_CPU keeps it (kv_cache.py:554-649, 694-696, 767-780) -----------------
    v_cache_cpu = torch.zeros(L, B, H, args.ctx // C_, C_ * D, dtype=torch.bfloat16, pin_memory=True)
    k_buf = torch.zeros(L, B, H, lay.buf_len, D, dtype=torch.bfloat16, device=dev)
    v_buf = torch.zeros_like(k_buf)
    U, SV, LM, LMI = [], [], [], []
    position_ids = torch.zeros(L, B, H, S, dtype=torch.int64, device=dev)
    for l in range(L):
        st = synth.layer_state(shape, B, args.ctx, seed=1000 + l, device=dev, cos_sin=cs)
        v_cache_cpu[l][:, :, :st["v_chunks"].shape[2]].copy_(st["v_chunks"])
        k_buf[l].copy_(st["k_buf"]); v_buf[l].copy_(st["v_buf"])
        U.append(st["U"]); SV.append(st["SV"]); LM.append(st["k_landmark"]); LMI.append(st["k_landmark_idx"])
        position_ids[l].copy_(st["resident"])
        del st
    gemm_o = torch.zeros(B, H, G, N, dtype=torch.bfloat16, device=dev)
    softmax_o = torch.zeros_like(gemm_o)
    T = (N + 255) // 256
    norm = torch.zeros(B * H, G, T, device=dev)
    ssum = torch.zeros(B * H, G, T, device=dev)
    temp = torch.zeros(B, H, S, C_ * D, dtype=torch.bfloat16, device=dev)
    output = torch.zeros(B, H, budget, D, dtype=torch.bfloat16, device=dev)
    offsets = torch.zeros(B * H * S, dtype=torch.int32, device=dev)
    cnts = torch.zeros(B * H, dtype=torch.int32, device=dev)
    signals = torch.zeros(B * H, dtype=torch.int32, device=dev)
    copy_stream = torch.cuda.Stream()
    kernel_offset, kernel_stride = lay.sparse_start * D, lay.buf_len * D
    max_ctx_chunks_len = (args.ctx // C_) * C_

    def layer_step(l, q, kn, vn, gen, stage_ev=None):
        """kv_cache.py:1227-1271, 983-1176 and base.py:320-341, call for call."""
        def mark(i):
            if stage_ev is not None:
                stage_ev[i].record()
        mark(0)
        pos = lay.sparse_end + gen
        v_buf[l][:, :, pos:pos + 1].copy_(vn, non_blocking=True)                 # update_kv_cache
        k_buf[l][:, :, pos:pos + 1].copy_(kn, non_blocking=True)
        mark(1)
        ref.batch_gemm_softmax(q.contiguous(), LM[l].contiguous(), gemm_o, norm, ssum, softmax_o, B * H, G, N, D,
                               1 / math.sqrt(128), 0)
        chunk_attn, _ = torch.max(softmax_o.view(B, H, G, -1), dim=-2)
        top = torch.topk(chunk_attn.view(B, H, -1), k=S, dim=-1).indices
        selected = LMI[l].gather(dim=-1, index=top)
        ref.reorder_keys_and_compute_offsets(position_ids[l], selected, offsets, cnts, B, H, S)
        mark(2)
        cur = torch.cuda.current_stream()
        with torch.cuda.stream(copy_stream):                                       # base.py:329-332
            copy_stream.wait_stream(cur)
            ref.gather_copy_with_offsets(v_cache_cpu[l], v_buf[l], temp, offsets, cnts, signals, B, H,
                                         int(max_ctx_chunks_len * D), int(budget * D), kernel_offset, kernel_stride, S)
        ref.gather_copy_d2d_with_offsets(k_buf[l], offsets, cnts, B, H, int(budget * D), kernel_offset, kernel_stride, S)
        pid32 = position_ids[l].to(torch.int32).contiguous()                      # tensor_op.py:218
        csc = cs.contiguous()
        ref.batch_gather_gemm(U[l].contiguous(), SV[l].contiguous(), csc, csc, pid32, output, B, H, args.ctx, D, R,
                              budget, cs.shape[0], C_, cnts)
        x = output
        ref.apply_rotary_pos_emb_push_cache_opt(
            x, cs, pid32, k_buf[l], cnts, B, H, budget, D, x.stride(0), x.stride(1), x.stride(2), x.stride(3),
            cs.stride(0), pid32.stride(0), pid32.stride(1), pid32.stride(2), k_buf[l].stride(0), k_buf[l].stride(1),
            k_buf[l].stride(2), lay.sparse_start, lay.sparse_end, D // 2, C_)
        mark(3)
        cur.wait_stream(copy_stream)                                              # base.py:338
        mark(4)
        kv_len = lay.sparse_end + gen + 1
        out = flash_attn_with_kvcache(q=q.transpose(1, 2), k_cache=k_buf[l][:, :, :kv_len].transpose(1, 2),
                                      v_cache=v_buf[l][:, :, :kv_len].transpose(1, 2), causal=True)
        mark(5)
        return out

    total = args.warmup + args.steps
    a = math.sqrt(max(0.0, 1.0 - args.noise ** 2))
    gq = torch.Generator(device=dev).manual_seed(4242)
    base = torch.randn(L, B, Hq, 1, D, generator=gq, device=dev)

    def inputs(t):
        g = torch.Generator(device=dev).manual_seed(100000 + t)
        q = (a * base + args.noise * torch.randn(L, B, Hq, 1, D, generator=g, device=dev)).to(torch.bfloat16)
        k = torch.randn(L, B, H, 1, D, generator=g, device=dev).to(torch.bfloat16)
        v = torch.randn(L, B, H, 1, D, generator=g, device=dev).to(torch.bfloat16)
        return q, k, v

    ins = [inputs(t) for t in range(total)]
    cap = lay.buf_len - lay.sparse_end
    hits = 0
    for t in range(args.warmup):
        q, k, v = ins[t]
        for l in range(L):
            layer_step(l, q[l], k[l], v[l], t % cap)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    hit_acc = torch.zeros((), dtype=torch.int64, device=dev)
    e0.record()
    for t in range(args.warmup, total):
        q, k, v = ins[t]
        for l in range(L):
            layer_step(l, q[l], k[l], v[l], t % cap)
            hit_acc += cnts.sum()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    hit_rate = float(hit_acc.item()) / (args.steps * L * B * H * S)
    assert int(signals.abs().sum()) == 0, "reference handshake counters must be back at zero"

    # per-stage times of one token (events inside the layer; adds a little launch overhead)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(6)] for _ in range(L)]
    q, k, v = ins[-1]
    for l in range(L):
        layer_step(l, q[l], k[l], v[l], (total - 1) % cap, evs[l])
    torch.cuda.synchronize()
    names = ["append", "score+topk+reorder", "K path (d2d+gemm+rope)", "wait V fetch", "flash_attn"]
    stages = {n: sum(evs[l][i].elapsed_time(evs[l][i + 1]) for l in range(L)) / L * 1e3 for i, n in enumerate(names)}

    print(json.dumps({
        "impl": "reference-cuda (oracle/_ref extension + flash-attn wheel, eager launches as in models/base.py)",
        "metric": "decode tok/s Llama-3.1-8B 122K ctx (sparse-KV decode step, all layers)",
        "value": args.steps * B / (ms * 1e-3), "unit": "tokens/s", "ms_per_step": ms / args.steps,
        "us_per_layer_step": ms / args.steps / L * 1e3, "steps": args.steps, "warmup": args.warmup,
        "chunk_cache_hit_rate": round(hit_rate, 4), "batch": B, "ctx": args.ctx, "layers": L,
        "us_per_stage": stages, "gpu": torch.cuda.get_device_name(0)}))


if __name__ == "__main__":
    main()
