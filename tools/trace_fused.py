#!/usr/bin/env python
"""Kernel tuning aid: wall-clock (%globaltimer) timeline of one layer step at the bench shape -- scoring, selection and
the fused fetch + rebuild + attention launch (tile CTAs and split CTAs of row 0).  Needs the trace build:
    make trace && SKV_LIB=shadowkv_b200/lib/libshadowkv_b200_trace.so python tools/trace_fused.py [--hbm]
Not part of the product path or of bench.py."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("SKV_LIB", os.path.join(ROOT, "shadowkv_b200", "lib", "libshadowkv_b200_trace.so"))

import torch  # noqa: E402
from shadowkv_b200 import _lib, synth  # noqa: E402
from shadowkv_b200.engine import DecodeStepGraph  # noqa: E402
from shadowkv_b200.kv_cache import ShadowKVCache_CPU  # noqa: E402


def read(unit):
    buf = (C.c_longlong * 8192)()
    fn = getattr(_lib.lib, f"skv_debug_read_trace_{unit}")
    fn.argtypes = [C.c_void_p, C.c_int]
    assert fn(buf, 8192) == 0
    return [[buf[c * 32 + s] for s in range(32)] for c in range(256)]


def main():
    dev = torch.device("cuda:0")
    shape = synth.MODELS["llama-3.1-8b"]
    ctx, L, B = 124928, 4, 1
    cs = synth.cos_sin_table(ctx + 1024, device=dev)
    cache = ShadowKVCache_CPU(shape.config(L), batch_size=B, max_length=ctx, device="cuda:0",
                              v_residency="hbm" if "--hbm" in sys.argv else "host")
    states = [synth.layer_state(shape, B, ctx, seed=1000 + l, device=dev, cos_sin=cs) for l in range(L)]
    synth.install(cache, states)
    run = DecodeStepGraph(cache, cs, host_io=False).capture(warmup=1)
    base = torch.randn(L, B, 32, 1, 128, device=dev)
    noise = 0.28                                   # --noise N: query innovation between tokens (1.0: every chunk misses)
    if "--noise" in sys.argv:
        noise = float(sys.argv[sys.argv.index("--noise") + 1])
    for t in range(6):
        q = ((1.0 - noise * noise) ** 0.5 * base + noise * torch.randn(L, B, 32, 1, 128, device=dev)).to(torch.bfloat16)
        k = torch.randn(L, B, 8, 1, 128, device=dev).to(torch.bfloat16)
        run.step_device(q, k, k)
    torch.cuda.synchronize()
    sc, se, fa = read("score_mma"), read("score"), read("fused_attend")
    t0 = sc[0][20]
    us = lambda t: "-" if t == 0 else f"{(t - t0) / 1e3:7.2f}"   # noqa: E731
    print(f"last layer of the last token; times in us since the scoring kernel's CTA (0,0) started; hits row 0: {int(cache.cnts_all[L - 1][0])}")
    for r, name in ((0, "cta(0,0)"), (1, "last cta")):
        print(f"score  {name}: start {us(sc[r][20])} past its wait {us(sc[r][21])} q staged {us(sc[r][22])} tiles in smem "
              f"{us(sc[r][24])} {us(sc[r][25])} {us(sc[r][26])} {us(sc[r][27])} end {us(sc[r][23])}")
    merged = [fa[t][14] for t in range(256) if fa[t][14] > t0]
    print(f"fused launch: row 0 merged {us(max(merged)) if merged else '-'} (a step takes one period; the previous layer's merge = this minus the period)")
    print(f"select row 0   : past its wait {us(se[0][21])} table written {us(se[0][22])}")
    k2 = se[17]
    if k2[1]:
        names_k = {1: "stats", 2: "keys+hist", 3: "exchange1", 4: "thr1", 5: "h2", 6: "exchange2", 7: "thr2", 10: "counts", 11: "emit",
                   12: "csync", 13: "ids", 14: "table"}
        print("select cluster kernel, row 0 CTA 0 (wall clock): " + " ".join(f"{n} {us(k2[i])}" for i, n in names_k.items() if k2[i] > t0))
    f = se[16]
    if f[0]:
        names_f = ["start", "past wait", "stats", "filter+scan", "compact", "keys+hist", "threshold", "emit", "table"]
        print("select_fast row 0: " + " ".join(f"{n} {us(f[i])}" for i, n in enumerate(names_f)) + f"  candidates {f[12]}")
    names = ["start", "idx", "ldV", "mma", "stats", "V1st", "issued", "landed", "Vall", "sync", "part", "waited", "dots", "end", "MERGED"]
    print("tile CTAs of row 0:  " + " ".join(f"{n:>7}" for n in names))
    for t in range(64):
        if fa[t][0] and fa[t][1]:
            print(f"  tile {t:2d}          " + " ".join(f"{us(fa[t][i]):>7}" for i in range(15)))
    names = ["start", "mask", "loop", "part", "waited", "end"]
    print("split CTAs of row 0: " + " ".join(f"{n:>7}" for n in names) + "  MERGED")
    for s in range(64):
        r = fa[64 + s]
        if r[0]:
            print(f"  split {s:2d}         " + " ".join(f"{us(r[i]):>7}" for i in range(6)) + f" {us(r[14]):>7}")


if __name__ == "__main__":
    main()
