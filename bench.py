#!/usr/bin/env python
"""Benchmark of the B200-native ShadowKV sparse-KV decode step (BASELINE.json metric:
decode tok/s, Llama-3.1-8B, 122K context; per-layer step us vs roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one decode token = the per-layer sparse-KV decode step of ALL layers of the model
(append -> landmark scoring -> top-k selection + chunk-table update -> {V hit compaction + miss fetch from
pinned host || K hit compaction + low-rank rebuild with RoPE} -> split-K attention), on synthetic decode state
of the named shape (random-init; there is no hub access).  The dense QKV / o_proj / MLP work between the layers
is outside this path and is NOT in the timed region -- `config.scope` says so.

  value  : tokens/s with the per-token inputs (q, k_new, v_new of every layer) already resident in HBM,
           all layers replayed as one CUDA graph, timed with CUDA events on the launching stream.
  e2e    : the same tokens/s through the public host API (DecodeStepGraph.step_host): HOST tensors in (pinned host
           memory, copied to the device inside the timed region every step), attention outputs of every layer copied
           back D2H, one host synchronisation per token.
  roofline : the landmark-scoring kernel (the dominant HBM kernel), timed alone with CUDA events, cycling through
           the landmark matrices of all layers (1 GB > L2) -- algorithmic bytes H*N*D*2 (+q) per (b) launch.
  cpu_baseline : the oracle's restatement of the reference PyTorch path (kv_cache.py:421-470) on the host cores,
           one layer per sample, scaled to the model's layer count.

N > 1 (torchrun): the batch is sharded, every rank owns `--batch` sequences with their own decode state and
pinned V store; no collective on the data path ("scaling": "weak").
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "decode tok/s Llama-3.1-8B 122K ctx (sparse-KV decode step, all layers)"
UNIT = "tokens/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=192)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="llama-3.1-8b")
    ap.add_argument("--ctx", type=int, default=124928)          # test/e2e.py:63-68 "122K"
    ap.add_argument("--batch", type=int, default=1)             # per GPU
    ap.add_argument("--layers", type=int, default=0)            # 0 = the model's layer count
    ap.add_argument("--budget", type=int, default=2048)
    ap.add_argument("--rank", type=int, default=160)
    ap.add_argument("--noise", type=float, default=0.28,
                    help="per-token query innovation; 0.28 gives ~60%% chunk-cache hits (reference test default)")
    ap.add_argument("--table-mode", default="inplace", choices=["inplace", "reference"],
                    help="chunk-table slot policy: in place (no hit compaction) or the reference's slot order")
    ap.add_argument("--shard", default="batch", choices=["batch", "heads"],
                    help="N > 1: shard the batch (weak scaling, no collective) or split the kv heads of the same "
                         "sequences over the ranks (strong scaling, one all-gather of the attention output per layer)")
    ap.add_argument("--hbm-store-ways", type=int, default=4096,
                    help="ways per (layer,b,h) of the optional HBM chunk store measured as a SECOND configuration "
                         "after the headline run (0 = skip); the headline numbers never use it")
    ap.add_argument("--no-kernel-table", action="store_true")
    ap.add_argument("--no-whole-model", action="store_true",
                    help="skip the whole-layer-stack section (dense remainder of every layer + lm_head, SURVEY.md 8 f2)")
    ap.add_argument("--whole-model-steps", type=int, default=48)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=float(d["hbm_gbs"]), bf16=float(d["bf16_tflops"]), src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, bf16=1590.0, src="fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.path = index, None, f"/tmp/skv_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "20"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for n, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        try:
            os.remove(self.path)
        except OSError:
            pass
        return out


# ------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's PyTorch path restated for the CPU (oracle), timed on host cores
# ------------------------------------------------------------------------------------------------------
def cpu_reference_layer_time(args, shape, n_layers, *, warmup: int, steps: int, max_seconds: float):
    """Times oracle.decode_step_uncached (reference models/kv_cache.py:421-470 + attention) for ONE layer of the
    named shape on the host cores.  Returns (seconds per layer-step, threads, steps done)."""
    import torch
    from oracle import shadowkv_oracle as O
    from shadowkv_b200 import synth
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    B = args.batch
    lay = synth.layout(args.ctx, args.budget)
    H, Hq, D = shape.num_key_value_heads, shape.num_attention_heads, 128
    g = torch.Generator().manual_seed(1000)
    bf = torch.bfloat16
    landmarks = torch.randn(B, H, lay.n_landmarks, D, generator=g).to(bf)
    lm_idx = torch.arange(lay.n_landmarks).expand(B, H, -1).contiguous()
    U = (torch.randn(B, args.ctx, args.rank, generator=g) / math.sqrt(args.rank)).to(bf)
    SV_rd = torch.randn(B, H, args.rank, D, generator=g).to(bf)
    v_full = torch.randn(B, H, args.ctx, D, generator=g).to(bf)
    k_buf = torch.randn(B, H, lay.buf_len, D, generator=g).to(bf)
    v_buf = torch.randn(B, H, lay.buf_len, D, generator=g).to(bf)
    cs = synth.cos_sin_table(args.ctx + 1024, theta=shape.rope_theta, glm=shape.glm_rope)
    geo = O.Geometry(sparse_start=lay.sparse_start, sparse_end=lay.sparse_end)
    rope = lambda x, tok: O.rope_torch(x, cs, tok)

    def one(i):
        q = torch.randn(B, Hq, 1, D, generator=g).to(bf)
        kn = torch.randn(B, H, 1, D, generator=g).to(bf)
        vn = torch.randn(B, H, 1, D, generator=g).to(bf)
        geo.gen_offset = i % 64
        O.decode_step_uncached(landmarks, lm_idx, U, SV_rd, v_full, k_buf, v_buf, geo, q, kn, vn, rope)

    for i in range(warmup):
        one(i)
    t0 = time.perf_counter()
    done = 0
    for i in range(steps):
        one(i)
        done += 1
        if time.perf_counter() - t0 > max_seconds:
            break
    dt = (time.perf_counter() - t0) / done
    return dt, threads, done


# ------------------------------------------------------------------------------------------------------
_JSON_OUT = None


def emit(line: dict) -> None:
    """The ONE JSON line goes to the real stdout; everything else the process (or NCCL, or torchrun's children)
    prints to fd 1 was redirected to stderr at start-up."""
    text = json.dumps(line) + "\n"
    if _JSON_OUT is None:
        sys.stdout.write(text)
        sys.stdout.flush()
    else:
        os.write(_JSON_OUT, text.encode())


def main():
    global _JSON_OUT
    # keep stdout clean for the driver: libraries (NCCL prints its version there) write to fd 1 directly
    sys.stdout.flush()
    _JSON_OUT = os.dup(1)
    os.dup2(2, 1)
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_gpus = max(args.gpus, world) if world > 1 else args.gpus

    if args.impl == "reference":
        if rank != 0:
            return 0
        from shadowkv_b200 import synth      # shapes only: the reference arm runs on the host cores
        shape = synth.MODELS[args.model]
        L = args.layers or shape.num_hidden_layers
        dt, threads, done = cpu_reference_layer_time(args, shape, L, warmup=args.warmup, steps=args.steps,
                                                     max_seconds=150.0)
        tok_s = args.batch / (dt * L)
        sample = (f"1 of {L} layers per step ({done} layer-steps timed after {args.warmup} warm-ups), full "
                  f"{args.ctx}-token shape, tokens/s = batch / (layer time x {L})")
        line = {"impl": "reference", "metric": METRIC, "value": tok_s, "unit": UNIT, "n_gpus": n_gpus,
                "steps": done, "warmup": args.warmup, "ms_per_step": dt * L * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": workload_config(args, shape, L, hit_rate=0.0, n_gpus=n_gpus),
                "cpu_baseline": {"value": tok_s, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
                "e2e": {"value": tok_s, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return 0

    return run_ours(args, rank, world, local_rank, n_gpus)


def workload_config(args, shape, L, *, hit_rate, n_gpus, heads_mode=False):
    from shadowkv_b200 import synth
    lay = synth.layout(args.ctx, args.budget)
    return {"workload": f"{shape.name} random-init decode state, {args.ctx} ctx ('122K', test/e2e.py shape), "
                        f"batch {args.batch}/GPU, sparse budget {args.budget}, rank {args.rank}, chunk 8, {L} layers",
            "scope": "sparse-KV decode step of every layer; the dense QKV/o_proj/MLP of the layers is timed around it in `whole_model` (SURVEY.md 8 f2), not in value / e2e",
            "batch_per_gpu": args.batch, "global_batch": args.batch * (1 if heads_mode else n_gpus), "ctx": args.ctx,
            "layers": L,
            "landmarks": lay.n_landmarks, "select_chunks": lay.select, "kv_rows_attended": lay.sparse_end + 1,
            "chunk_cache_hit_rate": round(hit_rate, 4), "query_noise": args.noise, "table_mode": args.table_mode,
            "fetch_rebuild_launch": "two launches, two streams" if os.environ.get("SKV_FUSED_FETCH") == "0" else "one launch",
            "l2_policy": "inputs larger than L2: per token the step streams every layer's landmarks, "
                         f"{L} x {lay.n_landmarks * shape.num_key_value_heads * 256 / 1e6:.1f} MB per sequence",
            "parallelism": (f"kv heads split x{n_gpus}, one all-gather of the attention output per layer" if heads_mode
                            else f"batch-sharded x{n_gpus}, no collective")}


def run_ours(args, rank, world, local_rank, n_gpus):
    import torch
    import torch.distributed as dist
    from shadowkv_b200 import _lib, synth
    from shadowkv_b200.engine import DecodeStepGraph
    from shadowkv_b200.kv_cache import ShadowKVCache_CPU
    import ctypes as C

    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = None
    if world > 1:
        from shadowkv_b200 import parallel as _par
        numa = _par.bind_to_gpu_numa_node(local_rank)       # before the pinned V store is allocated
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    full_shape = synth.MODELS[args.model]
    shape = full_shape
    heads_mode = args.shard == "heads" and world > 1
    if heads_mode:        # this rank's slice of the kv heads (and their query heads); same sequences on every rank
        import dataclasses
        from shadowkv_b200 import parallel
        cfg = parallel.head_shard_config(full_shape.config(1), world, rank)
        shape = dataclasses.replace(full_shape, num_key_value_heads=cfg.num_key_value_heads,
                                    num_attention_heads=cfg.num_attention_heads, hidden_size=cfg.hidden_size)
    L = args.layers or shape.num_hidden_layers
    B = args.batch
    H, Hq, D = shape.num_key_value_heads, shape.num_attention_heads, 128
    lay = synth.layout(args.ctx, args.budget)
    S = lay.select

    # ---- synthetic decode state, generated on the device ---------------------------------------------------
    t_setup = time.perf_counter()
    cs = synth.cos_sin_table(args.ctx + 1024, theta=shape.rope_theta, glm=shape.glm_rope, device=dev)
    cache = ShadowKVCache_CPU(shape.config(L), batch_size=B, max_length=args.ctx, device=str(dev),
                              sparse_budget=args.budget, rank=args.rank, table_mode=args.table_mode)
    states = []
    for l in range(L):
        st = synth.layer_state(shape, B, args.ctx, seed=1000 + l + 7919 * rank, rank=args.rank,
                               sparse_budget=args.budget, device=dev, cos_sin=cs)
        # move the V store to the pinned host tensor right away (8.2 GB per sequence for 32 layers)
        cache.v_cache_cpu[l][:, :, :st["v_chunks"].shape[2]].copy_(st["v_chunks"])
        st["v_chunks"] = None
        states.append(st)
    synth.install(cache, states, skip_v_store=True)
    del states
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t_setup

    total = args.warmup + args.steps
    gq = torch.Generator(device=dev).manual_seed(4242 + rank)
    a = math.sqrt(max(0.0, 1.0 - args.noise ** 2))
    base = torch.randn(L, B, Hq, 1, D, generator=gq, device=dev)
    capacity = lay.buf_len - lay.sparse_end            # generated tokens the buffer holds (96 at 122K)

    def inputs(t):
        g = torch.Generator(device=dev).manual_seed(100000 * (rank + 1) + t)
        q = (a * base + args.noise * torch.randn(L, B, Hq, 1, D, generator=g, device=dev)).to(torch.bfloat16)
        k = torch.randn(L, B, H, 1, D, generator=g, device=dev).to(torch.bfloat16)
        v = torch.randn(L, B, H, 1, D, generator=g, device=dev).to(torch.bfloat16)
        return q, k, v

    dev_inputs = [inputs(t) for t in range(total)]
    host_inputs = [tuple(x.cpu().pin_memory() for x in dev_inputs[t]) for t in range(total)]    # pinned host memory

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- value: device-resident inputs ------------------------------------------------------------------------
    run_d = DecodeStepGraph(cache, cs, host_io=False, gather_heads=heads_mode).capture(warmup=1)
    run_d.rewind(0)
    for t in range(args.warmup):
        run_d.step_device(*dev_inputs[t])
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    hits = torch.zeros(L, B * H, dtype=torch.int64, device=dev)
    with torch.cuda.stream(run_d.stream):
        ev0.record()
    for t in range(args.warmup, total):
        if cache.sparse_end + cache.gen_offset + 1 > cache.buf_len:
            run_d.rewind_async(0)
        run_d.step_device(*dev_inputs[t])
        with torch.cuda.stream(run_d.stream):
            hits += run_d.hits_d
    with torch.cuda.stream(run_d.stream):
        ev1.record()
    barrier()
    ms_dev = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    hit_rate = float(hits.sum().item()) / (args.steps * L * B * H * S)

    # ---- e2e: host tensors through the public API -----------------------------------------------------------------
    run_h = DecodeStepGraph(cache, cs, host_io=True, gather_heads=heads_mode).capture(warmup=0)
    run_h.rewind(0)
    for t in range(args.warmup):
        run_h.step_host(*host_inputs[t])
    barrier()
    t0 = time.perf_counter()
    for t in range(args.warmup, total):
        if cache.sparse_end + cache.gen_offset + 1 > cache.buf_len:
            run_h.rewind_async(0)
        out_h = run_h.step_host(*host_inputs[t])
    torch.cuda.synchronize()
    ms_e2e = (time.perf_counter() - t0) * 1e3
    checksum = float(out_h.float().abs().sum())

    # ---- second configuration: HBM chunk store between the sparse slots and the host store -----------------------
    store = None
    if args.hbm_store_ways > 0:
        cache.enable_hbm_store(args.hbm_store_ways)
        run_s = DecodeStepGraph(cache, cs, host_io=False, gather_heads=heads_mode).capture(warmup=0)
        run_s.rewind(0)
        fill_steps = 32                                     # tokens that warm the store before timing
        for t in range(fill_steps):
            if cache.sparse_end + cache.gen_offset + 1 > cache.buf_len:
                run_s.rewind_async(0)
            run_s.step_device(*dev_inputs[t % total])
        barrier()
        cache.pcie_count.zero_()
        hits.zero_()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(run_s.stream):
            s0.record()
        for t in range(args.warmup, total):
            if cache.sparse_end + cache.gen_offset + 1 > cache.buf_len:
                run_s.rewind_async(0)
            run_s.step_device(*dev_inputs[t])
            with torch.cuda.stream(run_s.stream):
                hits += run_s.hits_d
        with torch.cuda.stream(run_s.stream):
            s1.record()
        barrier()
        ms_store = s0.elapsed_time(s1)
        slots = args.steps * L * B * H * S
        store = {"ways": args.hbm_store_ways, "ms": ms_store,
                 "sparse_slot_hit_rate": float(hits.sum().item()) / slots,
                 "pcie_chunk_rate": float(cache.pcie_count.sum().item()) / slots,
                 "hbm_bytes": int(cache.vstore_data.numel() * 2 + cache.vstore_tags.numel() * 4),
                 "store_fill_tokens": fill_steps}
        cache.enable_hbm_store(0)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ms_dev = max_over_ranks(ms_dev)
    ms_e2e = max_over_ranks(ms_e2e)
    if store is not None:
        ms_store = max_over_ranks(store.pop("ms"))
        store.update(value=args.steps * B * (1 if heads_mode else n_gpus) / (ms_store * 1e-3), unit=UNIT,
                     ms_per_step=ms_store / args.steps, us_per_layer_step=ms_store / args.steps / L * 1e3,
                     note="same workload and outputs, byte for byte; chunks that fell out of the sparse slots are "
                          "served from a direct-mapped HBM store instead of crossing PCIe again (not part of the "
                          "headline value / e2e)")
    tokens = args.steps * B * (1 if heads_mode else n_gpus)      # the head split works on the SAME sequences
    value = tokens / (ms_dev * 1e-3)
    e2e = tokens / (ms_e2e * 1e-3)

    # ---- per-kernel timing (rank 0): CUDA events around back-to-back launches, cycling through the layers --------
    kernels, roofline, pcie = {}, None, None
    if rank == 0:
        kernels, roofline, pcie = kernel_table(args, cache, cs, shape, L, B, lay, peaks, dev,
                                               full=not args.no_kernel_table)

    # ---- whole layer stack (section 8 f2): dense remainder of every layer + lm_head around the same decode step ------
    whole = None
    if world == 1 and not args.no_whole_model and B <= 32 and not heads_mode:
        try:
            from shadowkv_b200.llama import Llama
            from tools.bench_decode_model import whole_model
            model = Llama(full_shape, batch_size=B, max_length=args.ctx, device=f"cuda:{local_rank}", num_layers=L,
                          kv_cache=cache)
            whole = whole_model(model, steps=args.whole_model_steps, warmup=4, noise=args.noise,
                                modes=("hot", "native", "library"), ops=True, store_ways=args.hbm_store_ways)
            whole["note"] = ("context for the headline, not part of value / e2e: decode tokens/s of the whole model with "
                             "random-init weights; the sparse-KV step consumes the same synthetic q/k/v as the headline")
            del model
        except Exception as e:  # noqa: BLE001  -- the headline must not die with the extra section
            whole = {"error": repr(e)[:300]}

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        dt, threads, done = cpu_reference_layer_time(args, shape, L, warmup=2, steps=1000, max_seconds=args.cpu_seconds)
        cpu = {"value": B / (dt * L), "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"oracle restatement of reference kv_cache.py:421-470 + attention, 1 of {L} layers "
                         f"x {done} layer-steps at the full {args.ctx}-token shape ({dt * 1e3:.2f} ms per layer-step), "
                         f"scaled by {L} layers"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
                "scaling": "strong" if heads_mode else "weak", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic",
                "config": workload_config(args, full_shape, L, hit_rate=hit_rate, n_gpus=n_gpus, heads_mode=heads_mode),
                "us_per_layer_step": ms_dev / args.steps / L * 1e3,
                "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": run_h.h2d_bytes,
                        "d2h_bytes_per_step": run_h.d2h_bytes, "ms_per_step": ms_e2e / args.steps,
                        "api": "shadowkv_b200.engine.DecodeStepGraph.step_host"},
                "gpu_launches": run_d.launches_per_step * args.steps,
                "hbm_chunk_store": store, "whole_model": whole, "roofline": roofline, "kernels": kernels, "pcie": pcie, "cpu_baseline": cpu, "clocks": clocks,
                "numa_binding_rank0": numa, "peaks": peaks, "lib": _lib.version(), "setup_s": round(setup_s, 1), "checksum": checksum}
        emit(line)
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        if heads_mode:
            # destroy_process_group() does not return while CUDA graphs that captured NCCL collectives are alive
            # (observed on torch 2.11 / NCCL 2.28): drop the graphs and leave without tearing the group down
            for r in (run_d, run_h):
                r.graph.reset()
            sys.stderr.flush()
            os._exit(0)
        dist.destroy_process_group()
    return 0


def kernel_table(args, cache, cs, shape, L, B, lay, peaks, dev, *, full=True):
    """Isolated per-kernel durations: each kernel launched back to back `reps` times between two CUDA events
    on the launching stream, cycling over the layers so the inputs never sit in L2 (32 layers of landmarks =
    1 GB).  Algorithmic bytes per launch are SURVEY.md section 8(d)'s per-(layer, sequence) figures x batch."""
    import ctypes as C
    import torch
    from shadowkv_b200._lib import lib, check
    H, G, D, S, N, R = cache.num_key_value_heads, cache.num_key_value_groups, 128, lay.select, lay.n_landmarks, args.rank
    g = cache._geometry(cs)
    q = torch.randn(L, B, H * G, 1, D, device=dev).to(torch.bfloat16)
    out = torch.empty(B, 1, H * G, D, dtype=torch.bfloat16, device=dev)

    def cur():
        return torch.cuda.current_stream().cuda_stream

    def timed(fn, reps):
        """`reps` back-to-back launches captured in one CUDA graph (no host launch gaps), replayed between two
        CUDA events on the launching stream."""
        for i in range(min(reps, L)):
            fn(i % L)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for i in range(reps):
                fn(i % L)
        graph.replay()
        torch.cuda.synchronize()
        st = torch.cuda.current_stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        graph.replay()
        e1.record(st)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / reps          # us per launch

    def score(l):
        check(lib.skv_score_landmarks(C.byref(g), q[l].data_ptr(), cache.k_landmark[l].data_ptr(),
                                      cache._sel_ws.data_ptr(), None, cur()), "score")

    reps = 4 * L
    t_score = timed(score, reps)
    score_bytes = B * (H * N * D * 2 + H * G * D * 2)
    roofline = {"kernel": "score_tcgen05_kernel (landmark scoring, TMA ring + tcgen05)", "bound": "hbm",
                "achieved": score_bytes / t_score * 1e-3, "peak": peaks["hbm"], "unit": "GB/s",
                "frac": score_bytes / t_score * 1e-3 / peaks["hbm"],
                # dram__bytes_read.sum + dram__bytes_write.sum of one launch at batch 1 / 8 kv heads / 15 560 landmarks,
                # from the ncu --set full capture summarised in profiles/r01_ncu_full_summary.txt
                "traffic": 31910400 if (B, H, N) == (1, 8, 15560) else None,
                "us_per_launch": t_score, "algorithmic_bytes_per_launch": score_bytes,
                "peak_source": peaks["src"], "launches_timed": reps}
    kernels = {"score": {"us": t_score, "GBps": score_bytes / t_score * 1e-3}}
    pcie = None
    if not full:
        return kernels, roofline, pcie

    sel = torch.zeros(B, H, S, dtype=torch.int64, device=dev)

    def select(l):
        check(lib.skv_select_from_scores(C.byref(g), cache.k_landmark_idx[l].data_ptr(),
                                         cache.position_ids[l].data_ptr(), cache.offsets.data_ptr(),
                                         cache.cnts.data_ptr(), 0, 0, cache._sel_ws.data_ptr(), cur()), "select")

    score(0)
    t_select = timed(select, 2 * L)
    kernels["select"] = {"us": t_select, "GBps": B * H * G * N * 2 / t_select * 1e-3}

    def attn(l):
        cache.attention(l, q[l], kv_len=lay.sparse_end + 1, out=out)

    t_attn = timed(attn, reps)
    attn_bytes = B * 2 * H * (lay.sparse_end + 1) * D * 2
    kernels["attention"] = {"us": t_attn, "GBps": attn_bytes / t_attn * 1e-3,
                            "frac_hbm": attn_bytes / t_attn * 1e-3 / peaks["hbm"]}

    # controlled hit rates for the data-movement kernels: offsets/cnts built directly
    for label, hit in (("h0", 0.0), ("h60", 154 / 256), ("h100", 1.0)):
        cnt = int(round(hit * S))
        inplace = cache.table_mode == "inplace"
        offs = torch.zeros(B * H, S, dtype=torch.int32, device=dev)
        res = torch.zeros(B * H, S, dtype=torch.int64, device=dev)
        for r in range(B * H):
            ids = torch.randperm(lay.chunks, device=dev)[:S]
            if inplace:      # resident[slot] = chunk id; offsets[i] = slot of the i-th miss (ascending slot and id)
                slots = torch.sort(torch.randperm(S, device=dev)[:S - cnt]).values
                ids[slots] = torch.sort(ids[slots]).values
                res[r] = ids
                offs[r, :S - cnt] = slots.to(torch.int32)
            else:            # hits: ascending source slots >= i; misses: ascending chunk ids (offset = id)
                if cnt:
                    offs[r, :cnt] = torch.sort(torch.randperm(S, device=dev)[:cnt]).values.to(torch.int32)
                if cnt < S:
                    offs[r, cnt:] = torch.sort(ids[cnt:]).values.to(torch.int32)
                res[r] = offs[r].to(torch.int64)
        cache.offsets.copy_(offs.view(-1))
        cache.cnts.fill_(cnt)
        res = res.view(B, H, S)
        miss = S - cnt

        def fetch(l):
            vh = cache.v_cache_cpu[l]
            check(lib.skv_fetch_misses(C.byref(g), vh.data_ptr(), vh.stride(1), cache.v_cache_buffer[l].data_ptr(),
                                       cache.offsets.data_ptr(), cache.cnts.data_ptr(), res.data_ptr(), None, cur()), "fetch")

        def compact(l):
            check(lib.skv_compact_hits(C.byref(g), cache.k_cache_buffer[l].data_ptr(), cache.v_cache_buffer[l].data_ptr(),
                                       cache.offsets.data_ptr(), cache.cnts.data_ptr(), cur()), "compact")

        def rebuild(l):
            check(lib.skv_rebuild_keys(C.byref(g), cache.U[l].data_ptr(), cache.SV[l].data_ptr(), cs.data_ptr(),
                                       res.data_ptr(), cache.cnts.data_ptr(), cache.offsets.data_ptr(), cache.k_cache_buffer[l].data_ptr(), 0, cur()),
                  "rebuild")

        def fetch_rebuild(l):            # the launch the step uses: both transfers of the missed chunks in one grid
            vh = cache.v_cache_cpu[l]
            check(lib.skv_fetch_and_rebuild(C.byref(g), vh.data_ptr(), vh.stride(1), cache.v_cache_buffer[l].data_ptr(),
                                            cache.U[l].data_ptr(), cache.SV[l].data_ptr(), cs.data_ptr(),
                                            cache.k_cache_buffer[l].data_ptr(), res.data_ptr(), cache.cnts.data_ptr(),
                                            cache.offsets.data_ptr(), None, cur()), "fetch_and_rebuild")

        row = {}
        if miss:
            by = B * H * miss * 2048
            t = timed(fetch_rebuild, 2 * L)
            row["fetch_rebuild_us"] = t
            row["fetch_rebuild_GBps_pcie"] = by / t * 1e-3
            t = timed(fetch, 2 * L)
            row["fetch_us"] = t
            row["fetch_GBps_pcie"] = by / t * 1e-3
            t = timed(rebuild, 2 * L)
            rd = B * H * miss * 8 * R * 2 + B * H * D * R * 2
            wr = B * H * miss * 8 * D * 2
            fl = 2.0 * B * H * miss * 8 * R * D
            row["rebuild_us"] = t
            row["rebuild_GBps"] = (rd + wr) / t * 1e-3
            row["rebuild_TFLOPs"] = fl / t * 1e-6
            row["rebuild_frac_tensor"] = fl / t * 1e-6 / peaks["bf16"]
        if cnt and not inplace:
            t = timed(compact, 2 * L)
            row["compact_kv_us"] = t
            row["compact_GBps"] = 2 * 2 * B * H * cnt * 2048 / t * 1e-3
        kernels[label] = row

    # PCIe denominator: pinned host -> device copy of 256 MiB with the copy engine
    n = 256 << 20
    hbuf = torch.empty(n, dtype=torch.uint8).pin_memory()
    dbuf = torch.empty(n, dtype=torch.uint8, device=dev)
    dbuf.copy_(hbuf, non_blocking=True)
    torch.cuda.synchronize()
    st = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(4):
        dbuf.copy_(hbuf, non_blocking=True)
    e1.record(st)
    torch.cuda.synchronize()
    pcie = {"h2d_memcpy_GBps": 4 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9}
    for label in ("h0", "h60"):
        if "fetch_GBps_pcie" in kernels.get(label, {}):
            kernels[label]["fetch_frac_pcie"] = kernels[label]["fetch_GBps_pcie"] / pcie["h2d_memcpy_GBps"]
            kernels[label]["fetch_rebuild_frac_pcie"] = kernels[label]["fetch_rebuild_GBps_pcie"] / pcie["h2d_memcpy_GBps"]
    return kernels, roofline, pcie


if __name__ == "__main__":
    sys.exit(main())
