#!/usr/bin/env python
"""Kernel tuning aid: per-layer time of the captured decode step with individual stages removed, to see what
each stage costs ON the critical path (as opposed to in isolation).  Numerics are meaningless in the ablated
variants; only the timing is read.  Not part of the product path or of bench.py."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
from shadowkv_b200 import _lib, synth  # noqa: E402
from shadowkv_b200.kv_cache import ShadowKVCache_CPU  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    shape = synth.MODELS["llama-3.1-8b"]
    ctx, L, B = 124928, int(os.environ.get("LAYERS", "16")), 1
    cs = synth.cos_sin_table(ctx + 1024, device=dev)
    cache = ShadowKVCache_CPU(shape.config(L), batch_size=B, max_length=ctx, device="cuda:0")
    for l in range(L):
        st = synth.layer_state(shape, B, ctx, seed=1000 + l, device=dev, cos_sin=cs)
        cache.v_cache_cpu[l][:, :, :st["v_chunks"].shape[2]].copy_(st["v_chunks"])
        st["v_chunks"] = None
        if l == 0:
            states = []
        states.append(st)
    synth.install(cache, states, skip_v_store=True)
    del states
    g = torch.Generator(device=dev).manual_seed(1)
    base = torch.randn(L, B, 32, 1, 128, generator=g, device=dev)
    steps = 24
    qs = [(0.96 * base + 0.28 * torch.randn(L, B, 32, 1, 128, generator=g, device=dev)).to(torch.bfloat16) for _ in range(steps)]
    kv = torch.randn(L, B, 8, 1, 128, generator=g, device=dev).to(torch.bfloat16)
    q_in = torch.zeros_like(qs[0])
    out = torch.zeros(L, B, 1, 32, 128, dtype=torch.bfloat16, device=dev)
    stream = torch.cuda.Stream()

    def body(skip):
        c = cache
        for l in range(L):
            c.incoming_q_len = 1
            geo = c._geometry(cs)
            ap = _lib.Append(k_new=kv[l].data_ptr(), v_new=kv[l].data_ptr(), k_buf=c.k_cache_buffer[l].data_ptr(),
                             v_buf=c.v_cache_buffer[l].data_ptr(), pos_dev=c.gen_dev.data_ptr(), pos=c.sparse_end)
            if "select" in skip and "score" not in skip:
                import ctypes as C
                _lib.check(_lib.lib.skv_score_landmarks(C.byref(geo), q_in[l].data_ptr(), c.k_landmark[l].data_ptr(),
                                                        c._sel_ws.data_ptr(), C.byref(ap),
                                                        torch.cuda.current_stream().cuda_stream), "score")
                pid = c.position_ids[l]
            elif "score" in skip:
                pid = c.position_ids[l]
            else:
                pid = c.get_retrieval_position_ids(l, q_in[l], append=ap)
            cur, side = torch.cuda.current_stream(), c.copy_stream
            if "nofork" in skip:                 # everything on one stream: what the side-stream fork / join costs
                if "rebuild" not in skip:
                    c.get_key_cache(l, pid, None, cs)
                if "fetch" not in skip:
                    c.get_value_cache(l, pid)
            else:
                side.wait_stream(cur)
                c._side_used = True
                with torch.cuda.stream(side):
                    if "rebuild" not in skip:
                        c.get_key_cache(l, pid, None, cs)
                    c._k_done.record(side)
                if "fetch" not in skip:
                    c.get_value_cache(l, pid)
                cur.wait_event(c._k_done)
            if "attention" not in skip:
                c.attention(l, q_in[l], None, out[l], True)
        if "nofork" not in skip:
            c.join_side_streams()

    def run(skip):
        cache.reset_device_counters()
        with torch.cuda.stream(stream):
            for t in range(2):
                q_in.copy_(qs[t]); body(set())               # real steps first so offsets / cnts are realistic
        stream.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            body(skip)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            for t in range(4):
                q_in.copy_(qs[t]); graph.replay()
            e0.record()
            for t in range(4, steps):
                q_in.copy_(qs[t]); graph.replay()
            e1.record()
        stream.synchronize()
        return e0.elapsed_time(e1) * 1e3 / (steps - 4) / L

    full = run(set())
    print(f"full step              {full:7.2f} us/layer   (hits {float(cache.cnts_all.float().mean()):.0f}/256)")
    for skip in (["nofork"], ["rebuild", "nofork"], ["attention"], ["fetch"], ["rebuild"], ["fetch", "rebuild"], ["fetch", "rebuild", "attention"],
                 ["select", "fetch", "rebuild", "attention"], ["score", "select", "rebuild", "attention"],
                 ["score", "select", "fetch", "rebuild"]):
        t = run(set(skip))
        print(f"without {'+'.join(skip):38s} {t:7.2f} us/layer   (delta {full - t:6.2f})")


if __name__ == "__main__":
    main()
