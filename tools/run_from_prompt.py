#!/usr/bin/env python
"""`test/e2e.py`-shaped run FROM TOKEN IDS (reference test/e2e.py:161-166, models/base.py:
500-655): a random prompt of --ctx tokens goes through ``Llama.batch_prefill`` (dense stack through library kernels,
decode state of every layer through this repo's get_svd / prefill_kv_cache), then --gen-len tokens are decoded through
the captured whole-token graph.  Random-init weights of the named architecture (no hub access), so the text is
meaningless; what is measured is the prompt time, the decode tokens/s and the chunk-cache hit rate the model's OWN
queries produce.

    python tools/run_from_prompt.py [--model llama-3.1-8b] [--ctx 124928] [--batch 1] [--gen-len 100]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
from shadowkv_b200 import synth  # noqa: E402
from shadowkv_b200.llama import Llama  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="llama-3.1-8b")
    ap.add_argument("--ctx", type=int, default=124928)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--gen-len", type=int, default=90,
                    help="clamped to the generation region of the K/V buffers (96 rows at 122K, reference kv_cache.py:574)")
    ap.add_argument("--layers", type=int, default=0)
    ap.add_argument("--v-residency", default="host", choices=["host", "hbm", "auto"])
    a = ap.parse_args()
    dev = "cuda:0"
    shape = synth.MODELS[a.model]
    t0 = time.perf_counter()
    from shadowkv_b200.kv_cache import ShadowKVCache_CPU
    layers = a.layers or shape.num_hidden_layers
    cache = ShadowKVCache_CPU(shape.config(layers), batch_size=a.batch, max_length=a.ctx, device=dev,
                              v_residency=a.v_residency)
    m = Llama(a.model, batch_size=a.batch, max_length=a.ctx, device=dev, num_layers=layers, kv_cache=cache)
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t0
    ids = torch.randint(0, shape.vocab_size, (a.batch, a.ctx), generator=torch.Generator().manual_seed(0)).to(dev)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    logits = m.batch_prefill(ids)
    torch.cuda.synchronize()
    prefill_s = time.perf_counter() - t1
    first = logits[:, -1, :].argmax(dim=-1, keepdim=True)
    t2 = time.perf_counter()
    m.kv_cache.H2D()
    torch.cuda.synchronize()
    h2d_s = time.perf_counter() - t2
    m.graph, m._prepared = None, False
    gen_len = min(a.gen_len, m.kv_cache.buf_len - m.kv_cache.sparse_end - 1)
    out, tps = m.batch_generate(first, gen_len=gen_len, benchmark=True)
    torch.cuda.synchronize()
    c = m.kv_cache
    hits = float(c.cnts_all.float().mean()) / c.select_sets
    print(json.dumps({
        "what": "prompt token ids -> batch_prefill -> H2D -> captured decode loop (reference test/e2e.py shape), random-init weights",
        "model": shape.name, "ctx": a.ctx, "batch": a.batch, "layers": layers, "gen_len": gen_len,
        "v_residency": c.v_residency if hasattr(c, "v_residency") else a.v_residency,
        "setup_s": round(setup_s, 2), "prefill_s": round(prefill_s, 2),
        "prefill_tokens_per_s": round(a.batch * a.ctx / prefill_s, 1), "h2d_s": round(h2d_s, 3),
        "decode_tokens_per_s": round(tps, 1), "decode_ms_per_token": round(1e3 * a.batch / tps, 3),
        "chunk_cache_hit_rate_last_token": round(hits, 3),
        "first_tokens": out[:, :8].tolist(), "kv_len": c.get_kv_len(),
        "peak_hbm_GB": round(torch.cuda.max_memory_allocated() / 1e9, 1),
    }))


if __name__ == "__main__":
    main()
