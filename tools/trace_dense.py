#!/usr/bin/env python
"""Kernel tuning aid: wall-clock (%globaltimer) phase timeline of dense_streamk_kernel over a chain of back-to-back
launches (the o_proj / QKV / gate-up / down shapes over distinct weights, captured in a CUDA graph).  Needs the trace
build:  make trace && python tools/trace_dense.py [o|qkv|gu|down] [batch]   Not part of the product path or of bench.py."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("SKV_LIB", os.path.join(ROOT, "shadowkv_b200", "lib", "libshadowkv_b200_trace.so"))

import torch  # noqa: E402
from shadowkv_b200 import _lib, dense  # noqa: E402

LABELS = ["start", "prefill", "pdl_wait", "setup", "mma0", "mma_end", "acc_last", "part_fence", "flag", "fixup", "epi", "end"]


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "o"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    shapes = {"o": (4096, 4096), "qkv": (6144, 4096), "gu": (28672, 4096), "down": (4096, 14336)}
    N, K = shapes[which]
    dev = torch.device("cuda:0")
    n = 12
    g = torch.Generator(device=dev).manual_seed(0)
    Ws = [(torch.randn(N, K, generator=g, device=dev) * 0.02).to(torch.bfloat16) for _ in range(n)]
    x = torch.randn(B, K, generator=g, device=dev).to(torch.bfloat16)
    out = torch.zeros(B, N // 2 if which == "gu" else N, dtype=torch.bfloat16, device=dev)

    def chain():
        for W in Ws:
            if which == "gu":
                dense.linear_silu_mul(x, W, out=out, static_weights=True)
            else:
                dense.linear(x, W, out=out, static_weights=True)

    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        chain()
    st.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=st):
        chain()
    with torch.cuda.stream(st):
        gr.replay()
        gr.replay()
    st.synchronize()
    buf = (C.c_longlong * 8192)()
    fn = _lib.lib.skv_debug_read_trace_dense
    fn.argtypes = [C.c_void_p, C.c_int]
    assert fn(buf, 8192) == 0
    rows = [[buf[r * 32 + s] for s in range(12)] for r in range(256)]
    launches = sorted({r // 16 for r in range(256) if rows[r][0]}, key=lambda L: min(rows[L * 16 + i][0] for i in range(16) if rows[L * 16 + i][0]))
    t0 = min(rows[launches[0] * 16 + i][0] for i in range(16) if rows[launches[0] * 16 + i][0])
    print(f"shape {which}: W [{N},{K}], {n} launches per replay; times in us since the first traced CTA start")
    for L in launches[-n:]:
        for i in (0, 1, 8, 15):
            r = rows[L * 16 + i]
            if not r[0]:
                continue
            print(f"launch {L:2d} cta_sel {i:2d}: " + " ".join(f"{LABELS[s]}={(r[s] - t0) / 1e3:.2f}" if r[s] else f"{LABELS[s]}=-" for s in range(12)))
        print()


if __name__ == "__main__":
    main()
