"""Experiment (tuning aid, not a product path): does the SM-issued chunk gather over PCIe run faster when the pinned V
store sits on transparent huge pages (anonymous mmap + madvise(MADV_HUGEPAGE) + cudaHostRegister) than in a cudaHostAlloc
buffer?  Scattered 2 KiB reads touch a different 4 KiB page each, so IOMMU / ATS translation misses would show here."""
import ctypes as C
import mmap
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from shadowkv_b200._lib import check, lib  # noqa: E402

rt = C.CDLL("libcudart.so.12")
libc = C.CDLL("libc.so.6", use_errno=True)
rt.cudaHostAlloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_uint]
rt.cudaHostRegister.argtypes = [C.c_void_p, C.c_size_t, C.c_uint]
rt.cudaHostUnregister.argtypes = [C.c_void_p]
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
libc.mmap.restype = C.c_void_p
libc.mmap.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_long]
libc.madvise.argtypes = [C.c_void_p, C.c_size_t, C.c_int]
libc.munmap.argtypes = [C.c_void_p, C.c_size_t]
dev = torch.device("cuda:0")
torch.zeros(1, device=dev)
for f in ("enabled", "defrag", "shmem_enabled"):
    try:
        print("thp", f, open(f"/sys/kernel/mm/transparent_hugepage/{f}").read().strip(), flush=True)
    except OSError as e:
        print("thp", f, "unreadable", e)
print(open("/proc/meminfo").read().split("AnonHugePages")[0].splitlines()[0], [l for l in open("/proc/meminfo") if "Huge" in l])
rows, chunks, n_ids = 8, 32768, 4096                       # 8 x 64 MB store, 64 MB gathered per launch
nbytes = rows * chunks * 2048
dst = torch.empty(rows, n_ids * 1024, dtype=torch.bfloat16, device=dev)
g = torch.Generator(device=dev).manual_seed(0)
ids = torch.stack([torch.randperm(chunks, generator=g, device=dev)[:n_ids].sort().values for _ in range(rows)]).contiguous()
big = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def bench(ptr, label):
    st = torch.cuda.current_stream()
    ts = []
    for _ in range(8):
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


def anon_huge():
    MAP_PRIVATE, MAP_ANONYMOUS, MADV_HUGEPAGE = 2, 0x20, 14
    p = libc.mmap(None, nbytes + (2 << 20), mmap.PROT_READ | mmap.PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0)
    assert p not in (None, C.c_void_p(-1).value), C.get_errno()
    a = (p + (2 << 20) - 1) & ~((2 << 20) - 1)                  # 2 MiB aligned start
    rc = libc.madvise(a, nbytes, MADV_HUGEPAGE)
    print("madvise rc", rc, "errno", C.get_errno() if rc else 0, flush=True)
    C.memset(a, 1, nbytes)
    print([l.strip() for l in open("/proc/meminfo") if "AnonHugePages" in l], flush=True)
    rc = rt.cudaHostRegister(a, nbytes, 0)
    assert rc == 0, f"cudaHostRegister rc={rc}"
    return p, a


p = C.c_void_p()
assert rt.cudaHostAlloc(C.byref(p), nbytes, 1) == 0
C.memset(p, 1, nbytes)
bench(p.value, "cudaHostAlloc")
base, a = anon_huge()
bench(a, "mmap + MADV_HUGEPAGE + cudaHostRegister")
bench(p.value, "cudaHostAlloc (again)")
rt.cudaHostUnregister(a)
libc.munmap(base, nbytes + (2 << 20))
rt.cudaFreeHost(p)
