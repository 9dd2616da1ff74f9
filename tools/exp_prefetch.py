"""Experiment: does an L2 prefetch make a later dense linear faster?  (tuning aid, not a product path)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from shadowkv_b200 import dense
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(1, 4096, generator=g, device=dev).to(torch.bfloat16)
out = torch.zeros(1, 4096, dtype=torch.bfloat16, device=dev)
flush = torch.zeros(512 << 20, dtype=torch.uint8, device=dev)
for mb, N in ((33.5, 4096), (67, 8192), (117, 14336)):
    W = (torch.randn(N, 4096, generator=g, device=dev) * 0.02).to(torch.bfloat16)
    o = torch.zeros(1, N, dtype=torch.bfloat16, device=dev)
    def t(prefetch, static):
        ts = []
        for _ in range(5):
            flush.fill_(1)
            torch.cuda.synchronize()
            if prefetch:
                dense.l2_prefetch(W)
                torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); dense.linear(x, W, out=o, static_weights=static); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        return min(ts)
    # time of the prefetch kernel itself + drain
    flush.fill_(1); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); dense.l2_prefetch(W); e1.record(); torch.cuda.synchronize()
    print(f"W {W.numel()*2/1e6:.1f} MB: cold {t(False, False):.1f} us, after l2_prefetch {t(True, False):.1f} us, prefetch kernel alone {e0.elapsed_time(e1)*1e3:.1f} us")
