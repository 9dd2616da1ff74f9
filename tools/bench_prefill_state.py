#!/usr/bin/env python
"""Measurement of the prefill-side state build at the headline shape: one layer of
Llama-3.1-8B at 124 928 tokens, batch 1.  Per stage: CUDA-event time and bytes moved against the measured HBM
copy peak; next to it the same stage done with the torch ops the reference calls (models/kv_cache.py:856-921 on the
GPU) and the reference's torch.svd next to the Gram-matrix factorisation.

    python tools/bench_prefill_state.py > gpurun_out/prefill_state.json
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
from shadowkv_b200 import _lib, synth  # noqa: E402
from shadowkv_b200.kv_cache import ShadowKVCache_CPU  # noqa: E402


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def main():
    dev = torch.device("cuda:0")
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    shape = synth.MODELS["llama-3.1-8b"]
    B, H, L, D, R = 1, 8, 124928, 128, 160
    g = torch.Generator(device=dev).manual_seed(0)
    n_all = L // 8
    k = (torch.randn(B, H, n_all, 1, D, generator=g, device=dev) + 0.7 * torch.randn(B, H, n_all, 8, D, generator=g, device=dev)
         ).reshape(B, H, L, D).to(torch.bfloat16)
    v = torch.randn(B, H, L, D, generator=g, device=dev).to(torch.bfloat16)
    k_pre = torch.randn(B, L, H * D, generator=g, device=dev).to(torch.bfloat16)
    last_q = torch.randn(B, 32, D, generator=g, device=dev).to(torch.bfloat16)
    lay = synth.layout(L)
    n, n_out = lay.chunks, lay.outlier_chunk
    means = torch.empty(B, H, n, D, dtype=torch.bfloat16, device=dev)
    mincos = torch.empty(B, H, n, dtype=torch.bfloat16, device=dev)
    outl = torch.empty(B, H, n_out, dtype=torch.int64, device=dev)
    rest = torch.empty(B, H, n - n_out, dtype=torch.int64, device=dev)
    lm = torch.empty(B, H, n - n_out, D, dtype=torch.bfloat16, device=dev)
    lib, check = _lib.lib, _lib.check
    res = {}
    t = timed(lambda: check(lib.skv_chunk_stats(k.data_ptr(), B, H, n, k.stride(0), k.stride(1), k.stride(2), means.data_ptr(),
                                                mincos.data_ptr(), 0), "stats"))
    by = B * H * n * 8 * D * 2 + means.numel() * 2
    res["chunk_stats"] = {"us": t, "GBps": by / t * 1e-3, "frac_hbm": by / t * 1e-3 / peaks["hbm_gbs"], "bytes": by}
    t = timed(lambda: check(lib.skv_select_outliers(mincos.data_ptr(), B * H, n, n_out, outl.data_ptr(), rest.data_ptr(), 0), "outl"))
    res["select_outliers"] = {"us": t}
    t = timed(lambda: check(lib.skv_gather_rows(means.data_ptr(), rest.data_ptr(), lm.data_ptr(), B * H, n, n - n_out, 0), "rows"))
    res["gather_landmarks"] = {"us": t, "GBps": 2 * lm.numel() * 2 / t * 1e-3}

    def torch_stats():        # the reference's ops (kv_cache.py:856-868) on the GPU
        kc = k[:, :, :n * 8].view(B, H, n, 8, D)
        cand = kc.mean(dim=-2)
        cs = torch.nn.functional.cosine_similarity(cand.unsqueeze(3).expand(-1, -1, -1, 8, -1), kc, dim=-1)
        return cs.min(dim=-1).values.topk(n_out, largest=False).indices

    res["reference_torch_ops_stats_plus_topk"] = {"us": timed(torch_stats, 3)}

    cache = ShadowKVCache_CPU(shape.config(1), batch_size=B, max_length=L, device="cuda:0")
    t = timed(lambda: (setattr(cache, "prefilled_batch", 0), cache.prefill_kv_cache(v, 0, k, last_q)), 3)
    res["prefill_kv_cache_whole_call"] = {"us": t, "note": "includes the 256 MB device -> pinned-host copy of V (PCIe) "
                                          "and the landmark copy to the host the reference keeps until H2D()"}
    t_gram = timed(lambda: (setattr(cache, "prefilled_batch", 0), cache.get_svd(k_pre, 0)), 3)
    cache.svd_method = "svd"
    t_svd = timed(lambda: (setattr(cache, "prefilled_batch", 0), cache.get_svd(k_pre, 0)), 1)
    res["get_svd"] = {"gram_eigh_us": t_gram, "reference_full_svd_us": t_svd, "speedup": t_svd / t_gram}
    print(json.dumps({"what": "prefill-side state build, one layer, Llama-3.1-8B shape, 124928 tokens, batch 1", "stages": res,
                      "hbm_peak_GBps": peaks["hbm_gbs"]}))


if __name__ == "__main__":
    main()
