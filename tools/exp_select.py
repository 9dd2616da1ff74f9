#!/usr/bin/env python
"""Experiment: isolated time of skv_select_chunks (scoring + selection) at the cfg2 shape, and how many clusters fit."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from shadowkv_b200 import _lib, synth
lib = _lib.lib
lib.skv_debug_max_active_clusters.argtypes = [C.c_int, C.c_int]
for csz in (8, 16):
    for smem in (100 * 1024, 113 * 1024, 160 * 1024, 215 * 1024):
        print("clusters of", csz, "with", smem // 1024, "KiB:", lib.skv_debug_max_active_clusters(csz, smem))
dev = torch.device("cuda:0")
B, H, G, N, S, L = int(os.environ.get("B", 1)), 8, 4, 15560, 256, 16
geom = _lib.Geom(batch=B, kv_heads=H, q_groups=G, head_dim=128, chunk=8, select=S, n_landmarks=N, rank=160, seq_len=N * 8 + 448,
                 buf_len=2592, sparse_start=448, table_mode=1)
ws = torch.empty(lib.skv_select_workspace_bytes(C.byref(geom)), dtype=torch.uint8, device=dev)
lm = torch.randn(L, B, H, N, 128, device=dev).to(torch.bfloat16)
idx = torch.arange(N, device=dev).expand(L, B, H, N).contiguous()
q = torch.randn(L, B, H * G, 1, 128, device=dev).to(torch.bfloat16)
res = torch.stack([torch.randperm(N, device=dev)[:S] for _ in range(L * B * H)]).view(L, B, H, S).contiguous()
off = torch.zeros(L, B * H * S, dtype=torch.int32, device=dev); cnt = torch.zeros(L, B * H, dtype=torch.int32, device=dev)
def run(l):
    _lib.check(lib.skv_select_chunks(C.byref(geom), q[l].data_ptr(), lm[l].data_ptr(), idx[l].data_ptr(), res[l].data_ptr(),
                                     off[l].data_ptr(), cnt[l].data_ptr(), 0, 0, ws.data_ptr(), None, torch.cuda.current_stream().cuda_stream), "sel")
for l in range(L): run(l)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for i in range(4 * L): run(i % L)
g.replay(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
print(f"B={B} SKV_FUSED_SELECT={os.environ.get('SKV_FUSED_SELECT')} SKV_SELECT_CLUSTER={os.environ.get('SKV_SELECT_CLUSTER')}: {e0.elapsed_time(e1) * 1e3 / (4 * L):.2f} us per launch")

if "trace" in os.environ.get("SKV_LIB", ""):
    buf = (C.c_longlong * 8192)()
    fn = lib.skv_debug_read_trace_score_select
    fn.argtypes = [C.c_void_p, C.c_int]
    assert fn(buf, 8192) == 0
    names = ["start", "setup", "stream", "stats", "keys+h1", "sync2", "thr1", "h2", "sync3", "thr2", "emit", "sync4", "table"]
    print("row 0 CTAs, us since the first CTA's start: " + " ".join(f"{n:>8}" for n in names))
    t0 = min(buf[r * 32] for r in range(16) if buf[r * 32])
    for r in range(16):
        if buf[r * 32]:
            print(f"  rank {r:2d}                                  " + " ".join(f"{(buf[r * 32 + i] - t0) / 1e3:8.2f}" if buf[r * 32 + i] else "       -" for i in range(13)))
