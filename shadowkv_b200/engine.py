"""Whole-token runner of the sparse-KV decode step: all layers of one decode token captured once in a CUDA
graph and replayed per token (the loop of reference models/base.py:147-148 + 315-341, minus the dense
QKV / MLP work that sits between the layers and is outside this path).

Two public entry points:

* ``step_device(q, k, v)``  -- inputs already on the GPU ([layers,B,Hq,1,D] / [layers,B,H,1,D] bf16);
  returns the device tensor of attention outputs [layers,B,1,Hq,D].
* ``step_host(q, k, v)``    -- inputs are HOST tensors; pinned tensors are copied to the device as they are,
  pageable ones are staged through the runner's pinned buffers first; the graph is replayed behind the three
  copies and ends with the copy of the attention outputs back to pinned host memory.  This is the call
  ``bench.py`` times as ``e2e``.

The graph holds, per layer: score (+ fused append) -> select -> V miss fetch + tcgen05 K rebuild (one launch) ->
split-K attention, i.e. 4 kernels chained by programmatic dependent launches (5 with table_mode="reference": one hit
compaction more; +1 with the HBM chunk store: its fill), plus one counter-advance kernel per token.  The generation offset lives on the
device (``ShadowKVCache_CPU.gen_dev`` / ``kv_len_dev``) so the same graph is valid for every token.
"""
from __future__ import annotations

import os

import torch

from . import parallel
from .kv_cache import ShadowKVCache_CPU


class DecodeStepGraph:
    def __init__(self, cache: ShadowKVCache_CPU, cos_sin_cache: torch.Tensor, *, host_io: bool = True,
                 gather_heads: bool = False):
        """gather_heads: kv-head split over the ranks of the default process group (parallel.head_shard_config):
        the attention output of every layer is all-gathered inside the graph (the one exchange of that mode)."""
        self.cache = cache
        self.gather = gather_heads
        self.cos_sin = cos_sin_cache
        self.host_io = host_io
        L, B = cache.num_layers, cache.batch_size
        Hq, H, D = cache.num_attention_heads, cache.num_key_value_heads, cache.head_dim
        dev, dt = cache.device, cache.dtype
        self.q_d = torch.zeros(L, B, Hq, 1, D, dtype=dt, device=dev)
        self.k_d = torch.zeros(L, B, H, 1, D, dtype=dt, device=dev)
        self.v_d = torch.zeros(L, B, H, 1, D, dtype=dt, device=dev)
        self.out_d = torch.zeros(L, B, 1, Hq, D, dtype=dt, device=dev)
        self.out_all = None
        self.xchg = None
        self.exchange = "none"
        if gather_heads:
            import torch.distributed as dist
            hooks_free = cache.fused_fetch_rebuild and cache.can_fuse_attend()
            if hooks_free and os.environ.get("SKV_HEAD_EXCHANGE", "p2p") != "nccl":
                # peer stores from inside the fused launch + one flag wait per layer (parallel.HeadExchange)
                self.xchg = parallel.HeadExchange(L, B, Hq, dev, D, rows=B * H)
                self.out_all = self.xchg.out_all
                self.exchange = ("P2P stores over NVLink from the merging CTA of skv_fetch_rebuild_attend into every rank's "
                                 "symmetric-memory buffer; the next layer's scoring launch waits for the peers' flags (no collective launch)")
            else:
                world = dist.get_world_size() if dist.is_initialized() else 1
                self.out_all = torch.zeros(L, world, B, 1, Hq, D, dtype=dt, device=dev)
                self.exchange = "NCCL all_gather_into_tensor, one node per layer inside the captured graph"
        # one GPU: per-row hand-over of a layer's attention output into the next layer's scoring launch
        # (parallel.RowHandover).  A measured experiment, OFF by default (SKV_OUT_FLAGS=1 enables it): bit-identical, but
        # not faster than the plain dependent launch -- 61.4 / 61.7 against 61.4 / 60.9 us per layer-step with V on the
        # host, 32.6 against 31.9 with V in HBM: the fence + flag store of the merging CTA and the poll round trip cost
        # what the grid's completion signal costs, so the gap between two layers is work (late rows), not launch latency
        self.handover = None
        if (not gather_heads and L > 1 and cache.fused_fetch_rebuild and cache.can_fuse_attend()
                and os.environ.get("SKV_OUT_FLAGS", "0") == "1"):
            self.handover = parallel.RowHandover(L, B * H, dev)
        if host_io:
            self.q_h = torch.zeros(L, B, Hq, 1, D, dtype=dt).pin_memory()
            self.k_h = torch.zeros(L, B, H, 1, D, dtype=dt).pin_memory()
            self.v_h = torch.zeros(L, B, H, 1, D, dtype=dt).pin_memory()
            self.out_h = (torch.zeros_like(self.out_all, device="cpu") if gather_heads
                          else torch.zeros(L, B, 1, Hq, D, dtype=dt)).pin_memory()
        self.h2d_bytes = (self.q_d.numel() + self.k_d.numel() + self.v_d.numel()) * 2
        self.d2h_bytes = (self.out_all if gather_heads else self.out_d).numel() * 2
        # kernels of libshadowkv_b200.so per token: score(+append), select, fetch(V) + rebuild(K) (one launch, or two
        # with SKV_FUSED_FETCH=0), attention (+ the hit compactions when the table keeps the reference's slot order:
        # one launch for both buffers in front of the fused launch, one per buffer otherwise)
        fused = cache.fused_fetch_rebuild and cache.rebuild_impl == 0
        per_layer = (4 if fused else 5) + (0 if cache.table_mode == "inplace" else (1 if fused else 2))
        if fused and cache.can_fuse_attend():
            per_layer = 3                      # score (+append), select, fetch + rebuild + attention
        if cache.select_fast_active():
            per_layer += 1                     # select_fast_kernel in front of the cluster selection
        self.launches_per_step = L * (per_layer + (1 if cache.hbm_store_ways else 0))
        self._per_layer = per_layer
        self.inline_wait = os.environ.get("SKV_HEAD_WAIT", "inline") != "launch"
        if self.xchg is not None:
            self.launches_per_step += 1 if self.inline_wait else L      # flag-wait launches of the peer-store exchange
        self.graph = None
        self.stream = torch.cuda.Stream(device=dev)
        self.replays = 0
        self._presets = {}

    @property
    def hits_d(self):
        """Chunk-cache hits per (layer, b*h) of the last token (the per-layer `cnts` the selection kernels wrote)."""
        return self.cache.cnts_all

    # ----------------------------------------------------------------------------------------------
    def _h2d(self, q=None, k=None, v=None):
        """Host -> device copies of one token's inputs on the current stream (sources must be pinned)."""
        self.q_d.copy_(self.q_h if q is None else q, non_blocking=True)
        self.k_d.copy_(self.k_h if k is None else k, non_blocking=True)
        self.v_d.copy_(self.v_h if v is None else v, non_blocking=True)

    def _body(self):
        c = self.cache
        for l in range(c.num_layers):
            # peer-store exchange: what consumes a layer's gathered output (o_proj; here the next layer's query) must not
            # start before every peer's rows have arrived.  Between layers the next scoring launch waits for the flags
            # itself (SKV_HEAD_WAIT=launch keeps a flag-wait launch per layer); the last layer gets the wait launch
            inline_wait = self.xchg is not None and self.inline_wait
            if self.handover is not None:
                # a layer's scoring launch directly follows the previous layer's fused launch on the stream: it starts on
                # that launch's row words (layer 0 follows the token's bookkeeping nodes and waits as usual)
                c.decode_layer(l, self.q_d[l], self.k_d[l], self.v_d[l], self.cos_sin, out=self.out_d[l],
                               device_counters=True, exchange=self.handover.struct(l),
                               wait_for=self.handover.wait_spec(l - 1) if l > 0 else None, wait_only=l > 0)
                continue
            c.decode_layer(l, self.q_d[l], self.k_d[l], self.v_d[l], self.cos_sin, out=self.out_d[l],
                           device_counters=True, exchange=self.xchg.struct(l) if self.xchg is not None else None,
                           wait_for=self.xchg.wait_spec(l - 1) if inline_wait and l > 0 else None)
            if self.xchg is not None:
                if not inline_wait or l == c.num_layers - 1:
                    self.xchg.wait(l)
            elif self.gather:
                parallel.gather_heads(self.out_d[l], self.out_all[l])
        c.join_side_streams()
        c.advance_device_counters()
        if self.xchg is not None:
            self.xchg.advance()
        if self.handover is not None:
            self.handover.advance()
        if self.host_io:
            self.out_h.copy_(self.out_all if self.gather else self.out_d, non_blocking=True)

    def capture(self, warmup: int = 1):
        """Warm up eagerly (allocates scratch, builds tensor maps, sets kernel attributes), then capture.
        The warm-up tokens are real decode steps: they advance the cache like any other token."""
        c = self.cache
        c.reset_device_counters()
        with torch.cuda.stream(self.stream):
            for _ in range(warmup):
                if self.host_io:
                    self._h2d()                  # the warm-up token's inputs are what the caller left in q_h / k_h / v_h
                self._body()
                c.note_replayed_steps(1)
        self.stream.synchronize()
        self.rewind_async(c.gen_offset)          # builds the preset for "here" before capture; no-op otherwise
        self._presets.setdefault(0, torch.tensor([0, c.sparse_end + 1], dtype=torch.int32, device=c.device))
        self.stream.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=self.stream):
            self._body()
        # capture does not execute: the cache state is unchanged by it
        return self

    # ----------------------------------------------------------------------------------------------
    def replay(self):
        """Enqueue one token on the runner's stream (no host synchronisation)."""
        c = self.cache
        if c.sparse_end + c.gen_offset + 1 > c.buf_len:
            raise RuntimeError(f"generation buffer full ({c.buf_len - c.sparse_end} tokens)")
        with torch.cuda.stream(self.stream):
            self.graph.replay()
        c.note_replayed_steps(1)
        self.replays += 1

    def step_device(self, q, k, v):
        with torch.cuda.stream(self.stream):
            self.q_d.copy_(q, non_blocking=True)
            self.k_d.copy_(k, non_blocking=True)
            self.v_d.copy_(v, non_blocking=True)
        self.replay()
        return self.out_d

    def step_host(self, q, k, v):
        """Host tensors in, pinned host tensor out; synchronises on the result (a user needs it to sample)."""
        if not self.host_io:
            raise RuntimeError("runner was built with host_io=False")
        srcs = []
        for t, stage in ((q, self.q_h), (k, self.k_h), (v, self.v_h)):
            direct = t.dtype == stage.dtype and t.shape == stage.shape and t.is_contiguous() and t.is_pinned()
            srcs.append(t if direct else stage.copy_(t))      # pageable / strided inputs go through the pinned stage
        with torch.cuda.stream(self.stream):
            self._h2d(*srcs)
        self.replay()
        self.stream.synchronize()                 # also covers the reads of the caller's pinned tensors
        return self.out_h

    def rewind_async(self, gen_offset: int = 0):
        """Stream-ordered rewind (no host synchronisation): the counters are reset by a device-to-device copy
        enqueued behind the tokens already in flight."""
        c = self.cache
        key = int(gen_offset)
        if key not in self._presets:
            self._presets[key] = torch.tensor([key, c.sparse_end + key + 1], dtype=torch.int32, device=c.device)
        c.kv_offset -= c.gen_offset - gen_offset
        c.gen_offset = gen_offset
        with torch.cuda.stream(self.stream):
            c._counters_dev.copy_(self._presets[key], non_blocking=True)

    def rewind(self, gen_offset: int = 0):
        """Reset the generation region (bench loops longer than the buffer's capacity): the sparse region,
        chunk table and landmarks are left as they are; only the append position goes back."""
        c = self.cache
        self.stream.synchronize()
        c.kv_offset -= c.gen_offset - gen_offset
        c.gen_offset = gen_offset
        c.reset_device_counters()
        torch.cuda.synchronize(c.device)
