"""Experiment (tuning aid, not a product path): are SM-issued reads of the pinned V store faster when the store is
allocated write-combined (cudaHostAllocWriteCombined)?  The store is written once (prefill) and then only read by the
GPU.  Times the chunk gather kernel and a copy-engine H2D over a default-pinned and a write-combined buffer."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from shadowkv_b200._lib import check, lib  # noqa: E402

rt = C.CDLL("libcudart.so.12")
rt.cudaHostAlloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_uint]
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
dev = torch.device("cuda:0")
torch.zeros(1, device=dev)
rows, chunks, n_ids = 8, 32768, 4096                       # 8 x 64 MB store, 64 MB gathered per launch
nbytes = rows * chunks * 2048
dst = torch.empty(rows, n_ids * 1024, dtype=torch.bfloat16, device=dev)
g = torch.Generator(device=dev).manual_seed(0)
ids = torch.stack([torch.randperm(chunks, generator=g, device=dev)[:n_ids].sort().values for _ in range(rows)]).contiguous()
big = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def bench(ptr, label):
    st = torch.cuda.current_stream()
    ts = []
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        check(lib.skv_gather_chunks(ptr, chunks * 1024, dst.data_ptr(), n_ids * 1024, 0, ids.data_ptr(), n_ids, rows,
                                    st.cuda_stream), "gather_chunks")
        e1.record(st)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    gather = rows * n_ids * 2048 / (min(ts[1:]) * 1e-3) / 1e9
    ms = []
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        rc = rt.cudaMemcpyAsync(big.data_ptr(), ptr, 256 << 20, 1, st.cuda_stream)
        e1.record(st)
        torch.cuda.synchronize()
        assert rc == 0, rc
        ms.append(e0.elapsed_time(e1))
    print(f"{label}: SM-issued chunk gather {gather:.1f} GB/s, copy-engine H2D {(256 << 20) / (min(ms[1:]) * 1e-3) / 1e9:.1f} GB/s", flush=True)


for label, flags in (("default pinned", 1), ("write-combined", 1 | 4), ("default pinned (again)", 1)):
    p = C.c_void_p()
    rc = rt.cudaHostAlloc(C.byref(p), nbytes, flags)
    assert rc == 0, f"cudaHostAlloc rc={rc}"
    C.memset(p, 1, nbytes)
    bench(p.value, label)
    rt.cudaFreeHost(p)
