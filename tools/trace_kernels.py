#!/usr/bin/env python
"""Kernel tuning aid: phase timelines (clock64 stamps) of select_kernel and attention_decode_kernel at the bench
shape.  Needs the trace build:  make trace && SKV_LIB=shadowkv_b200/lib/libshadowkv_b200_trace.so python tools/trace_kernels.py
Not part of the product path or of bench.py."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("SKV_LIB", os.path.join(ROOT, "shadowkv_b200", "lib", "libshadowkv_b200_trace.so"))

import torch  # noqa: E402
from shadowkv_b200 import _lib, synth  # noqa: E402
from shadowkv_b200.kv_cache import ShadowKVCache_CPU  # noqa: E402


def read(unit, n_cta, n_slot):
    buf = (C.c_longlong * 8192)()
    fn = getattr(_lib.lib, f"skv_debug_read_trace_{unit}")
    fn.argtypes = [C.c_void_p, C.c_int]
    assert fn(buf, 8192) == 0
    return [[buf[c * 32 + s] for s in range(n_slot)] for c in range(n_cta)]


def show(name, rows, labels):
    print(f"--- {name}: cycles since the CTA's first stamp (delta to previous phase)")
    for c, r in enumerate(rows):
        if r[0] == 0:
            continue
        out, prev = [], r[0]
        for i, v in enumerate(r):
            if v == 0:
                out.append(f"{labels[i]}=-")
                continue
            out.append(f"{labels[i]}={v - r[0]}(+{v - prev})")
            prev = v
        print(f"cta {c}: " + " ".join(out))


def main():
    dev = torch.device("cuda:0")
    shape = synth.MODELS["llama-3.1-8b"]
    ctx, L, B = 124928, 2, 1
    cs = synth.cos_sin_table(ctx + 1024, device=dev)
    cache = ShadowKVCache_CPU(shape.config(L), batch_size=B, max_length=ctx, device="cuda:0")
    states = [synth.layer_state(shape, B, ctx, seed=1000 + l, device=dev, cos_sin=cs) for l in range(L)]
    synth.install(cache, states)
    q, k, v = synth.queries(shape, B, seed=1, device=dev)
    for it in range(3):
        for l in range(L):
            cache.decode_layer(l, q, k, v, cs)
    torch.cuda.synchronize()
    show("select_kernel (bh 0, ranks 0..7)", read("score", 8, 15),
         ["start", "stats", "keys+h1", "csync1", "thr1", "h2", "csync2", "thr2", "counts", "csync3", "before", "emit",
          "csync4", "ids", "table"])
    show("score_tcgen05_kernel (bh 0, first CTAs)", read("score_mma", 6, 16),
         ["start", "setup", "tma0", "tma1", "tma2", "tma3", "full0", "full1", "full2", "full3", "acc0", "acc1", "acc2",
          "acc3", "epi_end", "end"])
    show("attention_decode_kernel (row 0, splits)", read("attention", 18, 8),
         ["start", "q", "loop", "partials", "fence+atomic", "last:fence", "weights", "end"])


if __name__ == "__main__":
    main()
