"""Host-side mirror of the reference's ShadowKV caches (reference models/kv_cache.py).

``ShadowKVCache_CPU`` keeps the reference's method names, argument meaning and state layout
 so ``models/base.py:296-341`` can drive it unchanged, but every decode-step
method lands in the fused sm_100a kernels behind the C ABI:

=============================  =====================================================================
reference method (kv_cache.py)  here
=============================  =====================================================================
update_kv_cache  (:1227-1271)   ``skv_append_kv`` (``decode_layer`` fuses it into the scoring launch)
get_retrieval_position_ids      ``skv_select_chunks``: TMA + tcgen05 scoring, then softmax + GQA max +
  (:983-1057)                   top-k + id gather + chunk-table update    (2 launches, was 8+)
get_value_cache  (:1059-1106)   ``skv_fetch_misses`` from pinned host memory (+ ``skv_compact_hits`` only
                                when ``table_mode="reference"``; optional HBM chunk store)
get_key_cache    (:1108-1176)   ``skv_rebuild_keys`` (tcgen05, RoPE in the epilogue; no bf16 scratch round
                                trip; + ``skv_compact_hits`` only in reference table mode)
(base.py:341 flash-attn call)   ``attention`` / ``shadowkv_b200.attention.flash_attn_with_kvcache``
=============================  =====================================================================

``decode_layer`` runs the whole per-layer step: score+append -> select -> V fetch + K rebuild of the missed chunks
(ONE launch, ``skv_fetch_and_rebuild``; ``SKV_FUSED_FETCH=0``: two launches on two streams as in base.py:326-338)
-> attention, every kernel launched programmatically behind its single predecessor.  The prefill-side methods (``get_svd``, ``prefill_kv_cache``) build the state with torch
ops; they are outside the hot path and kept only so the class is usable end to end.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import torch

from . import _lib
from ._lib import Geom, check, lib


def _cs() -> int:
    return torch.cuda.current_stream().cuda_stream


def hbm_store_ways_for(free_bytes: int, layers: int, rows: int, chunks_per_row: int, fraction: float = 0.5,
                       chunk_bytes: int = 2048) -> int:
    """Ways per (layer, b, h) of the HBM chunk store: the largest power of two whose store (2 KiB of V + a 4-byte tag
    per way) fits in `fraction` of `free_bytes`, capped at the first power of two that holds every chunk of a row
    (a direct-mapped store of that size never evicts).  0 when not even 128 ways fit."""
    budget = int(free_bytes * fraction) // max(1, layers * rows * (chunk_bytes + 4))
    cap = 1 << max(0, (max(1, chunks_per_row) - 1).bit_length())
    ways = 128
    if budget < ways:
        return 0
    while ways * 2 <= budget and ways * 2 <= cap:
        ways *= 2
    return min(ways, cap)


class ShadowKVCache_CPU:
    """Batched ShadowKV cache: V chunked in pinned host memory, sparse K/V ring on the GPU, low-rank
    K factors and landmarks on the GPU after ``H2D``.  Reference: models/kv_cache.py:509-1319."""

    def __init__(self, config, batch_size: int = 1, max_length: int = 32 * 1024, device: str = "cuda:0",
                 dtype=torch.bfloat16, sparse_budget: int = 2048, chunk_size: int = 8, rank: int = 160,
                 table_mode: str = "inplace", hbm_store_ways: int = 0, v_residency: str = "host") -> None:
        """table_mode: "inplace" (default; chunk-cache hits keep their slots, no compaction traffic) or
        "reference" (slot order of reference map.cuh:754-796: hits compacted to the front).
        hbm_store_ways: > 0 adds a direct-mapped HBM chunk store of that many 2 KiB ways per (layer, b, h)
        between the sparse slots and the pinned host store (power of two; 0 = the reference's behaviour: every
        chunk that is not in a sparse slot crosses PCIe).  Results are byte-identical either way.
        v_residency: where the full chunk-major V store lives (reference kv_cache.py:554-563 keeps it in pinned host
        memory, the only choice on an 80 GB part).  "host" = the reference's; "hbm" = in HBM (a 122K-token Llama-3.1-8B
        sequence is 8.2 GB of V against 180 GB), the miss fetch is then an HBM gather; "auto" = HBM when the whole
        store fits in `V_HBM_FRACTION` of the free HBM next to the factors and landmarks that follow, else host (and
        `apply_residency_policy()` then sizes an HBM chunk store from what is left).  Byte-identical results."""
        if dtype != torch.bfloat16:
            raise ValueError("the sm_100a kernels compute in bf16 only")
        if table_mode not in ("inplace", "reference"):
            raise ValueError("table_mode must be 'inplace' or 'reference'")
        self.table_mode = table_mode
        self.svd_method = "gram"     # "svd": the reference's full torch.svd (kv_cache.py:711)
        self.config = config
        self.batch_size = batch_size
        self.max_length = max_length
        self.device = torch.device(device)
        self.dtype = dtype
        self.num_attention_heads = config.num_attention_heads
        self.num_key_value_heads = config.num_key_value_heads
        self.num_key_value_groups = config.num_attention_heads // config.num_key_value_heads
        self.head_dim = config.hidden_size // config.num_attention_heads
        self.num_layers = config.num_hidden_layers

        self.sparse_budget = int(sparse_budget)
        self.chunk_size = chunk_size
        self.rank = rank
        self.local_chunk = 4                                             # kv_cache.py:547
        self.outlier_chunk = self._outlier_chunks()                      # kv_cache.py:548 (:206 for ShadowKVCache)
        self.select_sets = self.sparse_budget // self.chunk_size
        assert self.select_sets * self.chunk_size == self.sparse_budget, \
            f"({self.select_sets}) * {self.chunk_size} != {self.sparse_budget}"
        if self.head_dim != 128 or self.chunk_size != 8:
            raise ValueError("kernels are built for head_dim 128 and chunk_size 8 (2 KiB chunks)")
        if self.select_sets not in (128, 256, 512, 1024):
            raise ValueError("sparse_budget must be 1024, 2048, 4096 or 8192")

        L, B, H, D = self.num_layers, batch_size, self.num_key_value_heads, self.head_dim
        if v_residency not in ("host", "hbm", "auto"):
            raise ValueError("v_residency must be 'host', 'hbm' or 'auto'")
        self.v_residency = self._decide_v_residency(v_residency)
        # chunk-major V store (pinned host: kv_cache.py:554-563; the attribute keeps the reference's name in HBM too)
        self.v_cache_cpu = self._alloc_v_store(L, B, H, max_length // chunk_size, D * chunk_size)
        self.buf_len = self._buffer_len()
        self.k_cache_buffer = torch.zeros(L, B, H, self.buf_len, D, dtype=dtype, device=self.device)
        self.v_cache_buffer = torch.zeros(L, B, H, self.buf_len, D, dtype=dtype, device=self.device)

        self.kv_offset = 0
        self.prefill = 0
        self.gen_offset = 0
        self.prefilled_batch = 0
        self.k_landmark = None
        self.k_landmark_idx = None
        self.U = None
        self.SV = None

        self.block_num = B * H
        # offsets / cnts of the reference (kv_cache.py:626-627) are one scratch pair shared by all layers; here
        # every layer owns its pair, so kernels of layer l that run behind the critical path (the HBM store fill,
        # programmatically launched successors) never race with the selection of layer l+1.  `offsets` / `cnts`
        # are views of the layer that ran last.
        self.offsets_all = torch.zeros(L, self.block_num * self.select_sets, dtype=torch.int32, device=self.device)
        self.cnts_all = torch.zeros(L, self.block_num, dtype=torch.int32, device=self.device)
        self.offsets, self.cnts = self.offsets_all[0], self.cnts_all[0]
        self.position_ids = torch.full((L, B, H, self.select_sets), -1, dtype=torch.int64, device=self.device)

        self.hbm_store_ways = int(hbm_store_ways)
        if self.hbm_store_ways & (self.hbm_store_ways - 1):
            raise ValueError("hbm_store_ways must be 0 or a power of two")
        self.vstore_tags = self.vstore_data = None
        self.pcie_count = torch.zeros(L, B * H, dtype=torch.int32, device=self.device)
        if self.hbm_store_ways:
            W = self.hbm_store_ways
            self.vstore_tags = torch.full((L, B * H, W), -1, dtype=torch.int32, device=self.device)
            self.vstore_data = torch.empty(L, B * H, W, D * chunk_size, dtype=dtype, device=self.device)

        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._fetch_done = torch.cuda.Event()
        self._k_done = torch.cuda.Event()
        self._geom = None
        self._row_flags = None
        self._sel_hint = None
        self._sel_ws = None
        self._attn_ws = None
        self.rebuild_impl = 0      # 0: tcgen05 (product); 1: SIMT cross-check (tests only)
        # decode_layer: V fetch and K rebuild of the missed chunks as ONE launch (default) or as two launches on two
        # streams like the reference's choreography (SKV_FUSED_FETCH=0; kept for A/B measurements)
        self.fused_fetch_rebuild = os.environ.get("SKV_FUSED_FETCH", "1") != "0"
        self.programmatic_attention = os.environ.get("SKV_ATTN_PDL", "1") != "0"     # A/B knob, single-launch path only
        # decode_layer: V fetch + K rebuild + attention as ONE launch (skv_fetch_rebuild_attend; in-place table, G <= 8).
        # SKV_FUSED_ATTEND=0 keeps the attention as its own launch behind skv_fetch_and_rebuild (A/B measurements)
        self.fused_attend = os.environ.get("SKV_FUSED_ATTEND", "1") != "0"
        self._attend_ws = None
        self._side_used = False
        self.rope_glm = False
        # device-side copies of the generation counters, so a captured CUDA graph of the step can be replayed
        # while the cache grows (decode_layer(..., device_counters=True) + advance_device_counters())
        self._counters_dev = torch.zeros(2, dtype=torch.int32, device=self.device)
        self.gen_dev = self._counters_dev[0:1]
        self.kv_len_dev = self._counters_dev[1:2]

    # --- sizing hooks (the batch-1 GPU-resident variant overrides them) ---------------------------
    def _outlier_chunks(self) -> int:
        return int((self.sparse_budget // 1024) * 24)

    def _buffer_len(self) -> int:
        return self.sparse_budget + 128 + (self.outlier_chunk + self.local_chunk) * self.chunk_size   # kv_cache.py:574

    V_HBM_FRACTION = 0.6      # "auto": share of the free HBM the V store + factors + landmarks may take

    def state_bytes(self):
        """(V store bytes, bytes of the factors + landmarks + K/V buffers that live in HBM in every configuration)."""
        L, B, H, D = self.num_layers, self.batch_size, self.num_key_value_heads, self.head_dim
        v = L * B * H * (self.max_length // self.chunk_size) * D * self.chunk_size * 2
        n_lm = self.max_length // self.chunk_size
        other = L * B * (self.max_length * self.rank * 2 + H * D * self.rank * 2 + H * n_lm * (D * 2 + 8)
                         + 2 * H * self._buffer_len() * D * 2)
        return v, other

    def _decide_v_residency(self, want: str) -> str:
        if want != "auto":
            return want
        free, _ = torch.cuda.mem_get_info(self.device)
        v, other = self.state_bytes()
        return "hbm" if v + other <= self.V_HBM_FRACTION * free else "host"

    def _alloc_v_store(self, *shape):
        if self.v_residency == "hbm":
            return torch.zeros(*shape, dtype=self.dtype, device=self.device)
        return torch.zeros(*shape, dtype=self.dtype, device="cpu", pin_memory=True)

    def migrate_v_store(self, residency: str):
        """Move the V store between pinned host memory and HBM ("host" / "hbm" / "auto"), layer by layer.  Captured CUDA
        graphs of the step hold the old pointers and must be re-captured.  Returns the residency now in effect."""
        target = self._decide_v_residency(residency) if residency == "auto" else residency
        if target not in ("host", "hbm"):
            raise ValueError("residency must be 'host', 'hbm' or 'auto'")
        if target == self.v_residency:
            return target
        torch.cuda.synchronize(self.device)
        old = self.v_cache_cpu
        self.v_residency = target
        new = self._alloc_v_store(*old.shape)
        for l in range(old.shape[0]):
            new[l].copy_(old[l])
        torch.cuda.synchronize(self.device)
        self.v_cache_cpu = new
        return target

    def apply_residency_policy(self, fraction: float = 0.5):
        """The tiered policy in one call (after H2D / load_state): V store in HBM when it fits ("auto"), otherwise the
        pinned host store with a direct-mapped HBM chunk store sized from `fraction` of what is free.  Returns a dict
        describing the tiers chosen."""
        tier = self.migrate_v_store("auto")
        ways = 0
        if tier == "host":
            self.enable_hbm_store("auto", fraction)
            ways = self.hbm_store_ways
        v, other = self.state_bytes()
        return {"v_residency": tier, "hbm_chunk_store_ways": ways, "v_store_bytes": v, "other_state_bytes": other}

    # ------------------------------------------------------------------------------------------
    def print_stats(self):
        print(f"ShadowKV_CPU | sparse budget {self.sparse_budget} | chunk size {self.chunk_size} |rank {self.rank} "
              f"| cached {self.kv_offset} | local_chunk {self.local_chunk} | outlier_chunk {self.outlier_chunk}")

    def get_kv_len(self):
        return self.kv_offset

    def clear(self):
        self.k_cache_buffer.zero_()
        self.v_cache_buffer.zero_()
        self.k_landmark = self.k_landmark_idx = self.U = self.SV = None
        self.kv_offset = self.prefill = self.gen_offset = self.prefill_local = 0
        self.prefilled_batch = 0
        self.position_ids.fill_(-1)
        if self.vstore_tags is not None:
            self.vstore_tags.fill_(-1)
        self.pcie_count.zero_()
        self._geom = None

    # ------------------------------------------------------------------------------------------
    # geometry / scratch
    # ------------------------------------------------------------------------------------------
    def _set_layout(self, incoming: int) -> None:
        """Region boundaries of the K/V buffers for a prompt of `incoming` tokens (kv_cache.py:819-894)."""
        self.prefill = incoming
        self.max_ctx_chunks_len = (incoming // self.chunk_size) * self.chunk_size
        self.chunks = incoming // self.chunk_size - self.local_chunk
        self.chunks -= self.chunks % 8                                   # :836
        self.prefill_local = incoming - self.chunks * self.chunk_size
        self.sparse_start = self.prefill_local + self.outlier_chunk * self.chunk_size
        self.sparse_end = self.sparse_start + self.sparse_budget
        self.kernel_offset = self.sparse_start * self.head_dim
        self.kernel_stride = self.buf_len * self.head_dim
        if self.sparse_end > self.buf_len:
            raise ValueError("prompt leaves no room for the sparse region in the K/V buffers")

    def _geometry(self, cos_sin_cache=None) -> Geom:
        if self._geom is None:
            if self.k_landmark is None or self.U is None:
                raise RuntimeError("cache state not built: run prefill (or load_state) and H2D() first")
            g = Geom()
            g.batch, g.kv_heads, g.q_groups = self.batch_size, self.num_key_value_heads, self.num_key_value_groups
            g.head_dim, g.chunk, g.select = self.head_dim, self.chunk_size, self.select_sets
            g.n_landmarks = self.k_landmark.shape[-2]
            g.rank = self.U.shape[-1]
            g.seq_len = self.U.shape[-2]
            g.buf_len, g.sparse_start = self.buf_len, self.sparse_start
            g.table_mode = 1 if self.table_mode == "inplace" else 0
            self._geom = g
            self._sel_ws = torch.empty(lib.skv_select_workspace_bytes(C.byref(g)), dtype=torch.uint8,
                                       device=self.device)
            self._attn_ws = torch.zeros(
                lib.skv_attention_workspace_bytes(g.batch, g.kv_heads, g.q_groups, g.head_dim, 0),
                dtype=torch.uint8, device=self.device)
            self._attend_ws = torch.zeros(lib.skv_attend_workspace_bytes(g.batch, g.kv_heads, g.q_groups),
                                          dtype=torch.uint8, device=self.device)
            # per-row hand-over words between the selection and the fused fetch launch (skv_geom.row_flags), one array
            # per layer; SKV_ROW_FLAGS=0 keeps plain dependent launches (A/B runs)
            self._row_flags = (torch.zeros(self.num_layers, g.batch * g.kv_heads, dtype=torch.int32, device=self.device)
                               if os.environ.get("SKV_ROW_FLAGS", "1") != "0" else None)
            # threshold hints of the selection (skv_geom.select_hint): per (layer, b, kv head) the bf16 key the previous
            # token's top-S threshold ended on; a speed hint only (the selection is the same with any content)
            self._sel_hint = torch.zeros(self.num_layers, g.batch * g.kv_heads, dtype=torch.int32, device=self.device)
        if cos_sin_cache is not None:
            width = cos_sin_cache.shape[-1]
            if width not in (self.head_dim, self.head_dim // 2):
                raise ValueError(f"Invalid cos_sin shape {tuple(cos_sin_cache.shape)}")     # tensor_op.py:197
            self._geom.rope_glm = 1 if width == self.head_dim // 2 else 0
            self._geom.cos_sin_rows = cos_sin_cache.shape[0]
        return self._geom

    # ------------------------------------------------------------------------------------------
    # prefill side (torch ops; outside the hot path)
    # ------------------------------------------------------------------------------------------
    def get_svd(self, new_k_cache, layer_idx):
        """Rank-`rank` factorisation of the pre-RoPE keys (kv_cache.py:666-737): U [L, rank], SV [H, D, rank].
        The reference runs a full fp32 torch.svd of the [L, H*D] key matrix.  Here (svd_method="gram", default) the
        right singular vectors come from the eigen-decomposition of the H*D x H*D Gram matrix K^T K -- one GEMM
        over the keys and a 1024 x 1024 symmetric eigensolve instead of an SVD of a 124928 x 1024 matrix -- and
        U = K V / sigma.  Same subspace and the same product U.SV up to fp32 rounding; the factors themselves are
        only defined up to sign.  svd_method="svd" keeps the reference's call."""
        if new_k_cache.shape[1] <= 32:
            k = new_k_cache.transpose(1, 2).reshape(new_k_cache.shape[0], -1, self.num_key_value_heads * self.head_dim)
        else:
            k = new_k_cache
        bsz, L = k.shape[0], k.shape[1]
        if layer_idx == 0 and self.prefilled_batch == 0:
            self.U = torch.zeros(self.num_layers, self.batch_size, L, self.rank, dtype=self.dtype, device="cpu")
            self.SV = torch.zeros(self.num_layers, self.batch_size, self.num_key_value_heads, self.head_dim,
                                  self.rank, dtype=self.dtype, device="cpu")
        kf = k.float()
        if self.svd_method == "gram":
            gram = kf.transpose(1, 2) @ kf                                           # [bsz, HD, HD] fp32
            evals, evecs = torch.linalg.eigh(gram)                                   # ascending
            top = evecs[:, :, -self.rank:].flip(-1)                                  # [bsz, HD, rank], largest first
            sigma = evals[:, -self.rank:].flip(-1).clamp_min(0).sqrt()               # [bsz, rank]
            u = (kf @ top) / sigma.clamp_min(1e-20).unsqueeze(1)                     # [bsz, L, rank]
            sv = sigma.unsqueeze(-1) * top.transpose(1, 2)                           # [bsz, rank, HD]
        else:
            uu, ss, vh = torch.linalg.svd(kf, full_matrices=False)
            u = uu[:, :, :self.rank]
            sv = ss[:, :self.rank, None] * vh[:, :self.rank]
        sl = slice(self.prefilled_batch, self.prefilled_batch + bsz)
        self.U[layer_idx][sl].copy_(u.to(self.dtype))
        sv = sv.to(self.dtype).view(bsz, self.rank, self.num_key_value_heads, self.head_dim).permute(0, 2, 3, 1)
        self.SV[layer_idx][sl].copy_(sv)                                             # [bsz, H, D, rank]

    def register_k_landmark(self, k_landmark, k_landmark_idx, layer_idx):
        """kv_cache.py:739-786"""
        n = k_landmark.shape[-2]
        bsz = k_landmark.shape[0]
        if layer_idx == 0 and self.prefilled_batch == 0:
            self.k_landmark = torch.zeros(self.num_layers, self.batch_size, self.num_key_value_heads, n,
                                          self.head_dim, dtype=self.dtype, device="cpu")
            self.k_landmark_idx = torch.zeros(self.num_layers, self.batch_size, self.num_key_value_heads, n,
                                              dtype=torch.long, device="cpu")
        sl = slice(self.prefilled_batch, self.prefilled_batch + bsz)
        self.k_landmark[layer_idx][sl].copy_(k_landmark)
        self.k_landmark_idx[layer_idx][sl].copy_(k_landmark_idx)

    def prefill_kv_cache(self, new_v_cache, layer_idx, key_states_roped, last_query_states=None, outliers_out=None):
        """Builds one layer of decode state from the prompt's K/V (kv_cache.py:788-980).  The chunk statistics
        (bf16 chunk means, min cosine similarity per chunk), the outlier / landmark split, the landmark compaction
        and all chunk gathers run in the CUDA kernels of csrc/prefill.cu and csrc/gather.cu; the first fill of the
        sparse region reuses the decode path's scoring + selection kernels.  outliers_out (optional int64
        [bsz,H,outlier_chunk]) receives the outlier chunk ids (tests)."""
        bsz, H, incoming, D = new_v_cache.shape
        C_ = self.chunk_size
        self._set_layout(incoming)
        sl = slice(self.prefilled_batch, self.prefilled_batch + bsz)
        st = _cs()
        n_chunks_all = incoming // C_
        self.v_cache_cpu[layer_idx][sl, :, :n_chunks_all].copy_(
            new_v_cache[:, :, :self.max_ctx_chunks_len].reshape(bsz, H, n_chunks_all, C_ * D), non_blocking=True)
        kb, vb = self.k_cache_buffer[layer_idx][sl], self.v_cache_buffer[layer_idx][sl]
        # local window = the tail of the prompt
        kb[:, :, :self.prefill_local].copy_(key_states_roped[:, :, -self.prefill_local:])
        vb[:, :, :self.prefill_local].copy_(new_v_cache[:, :, -self.prefill_local:])

        n, n_out = self.chunks, self.outlier_chunk
        k = key_states_roped if key_states_roped.is_contiguous() else key_states_roped.contiguous()
        v = new_v_cache if new_v_cache.is_contiguous() else new_v_cache.contiguous()
        means = torch.empty(bsz, H, n, D, dtype=self.dtype, device=k.device)
        mincos = torch.empty(bsz, H, n, dtype=self.dtype, device=k.device)
        check(lib.skv_chunk_stats(k.data_ptr(), bsz, H, n, k.stride(0), k.stride(1), k.stride(2), means.data_ptr(),
                                  mincos.data_ptr(), st), "chunk_stats")
        outliers = torch.empty(bsz, H, n_out, dtype=torch.int64, device=k.device)
        rest_idx = torch.empty(bsz, H, n - n_out, dtype=torch.int64, device=k.device)
        check(lib.skv_select_outliers(mincos.data_ptr(), bsz * H, n, n_out, outliers.data_ptr(), rest_idx.data_ptr(), st),
              "select_outliers")
        if outliers_out is not None:
            outliers_out.copy_(outliers)
        buf_stride = self.buf_len * D
        for src, dst in ((k, kb), (v, vb)):             # outlier chunks, exact K and V (kv_cache.py:870-907)
            check(lib.skv_gather_chunks(src.data_ptr(), src.stride(1), dst.data_ptr(), buf_stride, self.prefill_local * D,
                                        outliers.data_ptr(), n_out, bsz * H, st), "gather_chunks(outliers)")
        lm = torch.empty(bsz, H, n - n_out, D, dtype=self.dtype, device=k.device)
        check(lib.skv_gather_rows(means.data_ptr(), rest_idx.data_ptr(), lm.data_ptr(), bsz * H, n, n - n_out, st),
              "gather_rows")
        self.register_k_landmark(lm, rest_idx, layer_idx)

        # first fill of the sparse region with the prompt's last query (kv_cache.py:926-970): the decode path's
        # scoring + selection kernels on an empty chunk table (every selected chunk is a miss)
        g = Geom(batch=bsz, kv_heads=H, q_groups=self.num_key_value_groups, head_dim=D, chunk=C_, select=self.select_sets,
                 n_landmarks=n - n_out, rank=self.rank, seq_len=incoming, buf_len=self.buf_len,
                 sparse_start=self.sparse_start, table_mode=1,
                 flags=_lib.GEOM_DYNAMIC_LANDMARKS)      # `lm` is the previous kernel's output: no early landmark loads
        ws = torch.empty(lib.skv_select_workspace_bytes(C.byref(g)), dtype=torch.uint8, device=k.device)
        q = last_query_states.reshape(bsz, H * self.num_key_value_groups, D).contiguous()
        pid = self.position_ids[layer_idx][sl]
        pid.fill_(-1)
        offs = torch.empty(bsz * H * self.select_sets, dtype=torch.int32, device=k.device)
        cnt = torch.empty(bsz * H, dtype=torch.int32, device=k.device)
        check(lib.skv_select_chunks(C.byref(g), q.data_ptr(), lm.data_ptr(), rest_idx.data_ptr(), pid.data_ptr(),
                                    offs.data_ptr(), cnt.data_ptr(), 0, 0, ws.data_ptr(), None, st), "select_chunks(prefill)")
        for src, dst in ((k, kb), (v, vb)):             # exact K and V of the selected chunks
            check(lib.skv_gather_chunks(src.data_ptr(), src.stride(1), dst.data_ptr(), buf_stride, self.sparse_start * D,
                                        pid.data_ptr(), self.select_sets, bsz * H, st), "gather_chunks(selected)")

        if layer_idx == self.num_layers - 1:
            assert self.sparse_budget < incoming
            self.prefilled_batch += bsz
            if self.prefilled_batch == self.batch_size:
                self.kv_offset += incoming
                assert not bool(torch.any(self.position_ids == -1)), "the cache for offloading is not built correctly"

    def H2D(self):
        """Move factors and landmarks to the GPU (kv_cache.py:1178-1225)."""
        for name in ("SV", "U", "k_landmark", "k_landmark_idx"):
            t = getattr(self, name)
            if t is not None and not t.is_cuda:
                setattr(self, name, t.to(self.device))
        self._geom = None
        torch.cuda.synchronize(self.device)

    def load_state(self, *, prompt_len: int, U, SV, k_landmark, k_landmark_idx, position_ids=None):
        """Install externally built decode state (synthetic benches, tests): tensors laid out as
        SURVEY.md section 8 a1.  V store and K/V buffers are filled by the caller through the public
        attributes.  Replaces get_svd/prefill_kv_cache + H2D."""
        self._set_layout(prompt_len)
        self.U, self.SV, self.k_landmark, self.k_landmark_idx = U, SV, k_landmark, k_landmark_idx
        if position_ids is not None:
            self.position_ids.copy_(position_ids)
        self.kv_offset = prompt_len
        self.prefilled_batch = self.batch_size
        self.H2D()

    # ------------------------------------------------------------------------------------------
    # decode step
    # ------------------------------------------------------------------------------------------
    def update_kv_cache(self, new_k_cache, new_v_cache, layer_idx, device_counters=False):
        """Append the new token's post-RoPE K and V rows (kv_cache.py:1227-1271)."""
        incoming = new_k_cache.shape[-2]
        if incoming != 1:
            raise ValueError("decode appends one token per step")
        g = self._geometry()
        pos = self.sparse_end + (0 if device_counters else self.gen_offset)
        if pos >= self.buf_len:
            raise RuntimeError(f"generation buffer full ({self.buf_len - self.sparse_end} tokens)")
        kc, vc = new_k_cache.contiguous(), new_v_cache.contiguous()     # named: temporaries must outlive the launch
        check(lib.skv_append_kv(C.byref(g), kc.data_ptr(), vc.data_ptr(),
                                self.k_cache_buffer[layer_idx].data_ptr(), self.v_cache_buffer[layer_idx].data_ptr(),
                                pos, self.gen_dev.data_ptr() if device_counters else 0, _cs()), "append_kv")
        if layer_idx == self.num_layers - 1 and not device_counters:
            self.kv_offset += incoming
            self.gen_offset += incoming

    # --- device-side counters (CUDA-graph replay) ---------------------------------------------------------
    def reset_device_counters(self):
        """Copy the host-side generation offset into the device counters (call before capture / replay)."""
        self._counters_dev.copy_(torch.tensor([self.gen_offset, self.sparse_end + self.gen_offset + 1],
                                              dtype=torch.int32), non_blocking=False)

    def advance_device_counters(self):
        """gen += 1 on the device; part of the captured step."""
        self._counters_dev.add_(1)

    def note_replayed_steps(self, n: int):
        """Host-side bookkeeping after `n` graph replays (kv_cache.py:1269-1271 done in bulk)."""
        self.gen_offset += n
        self.kv_offset += n
        if self.sparse_end + self.gen_offset > self.buf_len:
            raise RuntimeError("generation buffer overrun during graph replay")

    def get_retrieval_position_ids(self, layer_idx, query_states, selected_out=None, score_out=None, append=None):
        """Score landmarks, pick the top chunks and update the chunk table (kv_cache.py:983-1057).
        Returns position_ids[layer] (chunk id resident in every sparse slot), updated in place.
        append: optional _lib.Append -- fuses update_kv_cache's row copy into the scoring launch."""
        self.incoming_q_len = query_states.shape[-2]
        if self.incoming_q_len != 1:
            raise ValueError("decode scores one query token per step")
        g = self._geometry()
        q = query_states.contiguous()
        self.offsets, self.cnts = self.offsets_all[layer_idx], self.cnts_all[layer_idx]
        g.select_hint = self._sel_hint[layer_idx].data_ptr()
        check(lib.skv_select_chunks(C.byref(g), q.data_ptr(), self.k_landmark[layer_idx].data_ptr(),
                                    self.k_landmark_idx[layer_idx].data_ptr(), self.position_ids[layer_idx].data_ptr(),
                                    self.offsets.data_ptr(), self.cnts.data_ptr(),
                                    selected_out.data_ptr() if selected_out is not None else 0,
                                    score_out.data_ptr() if score_out is not None else 0,
                                    self._sel_ws.data_ptr(), C.byref(append) if append is not None else None, _cs()),
              "select_chunks")
        return self.position_ids[layer_idx]

    def _view_len(self, layer_idx):
        gen = self.gen_offset if layer_idx == self.num_layers - 1 else self.gen_offset + self.incoming_q_len
        return self.sparse_end + gen

    def get_value_cache(self, layer_idx, position_ids):
        """Compact the hit chunks, fetch the missed ones from pinned host memory (kv_cache.py:1059-1106)."""
        g = self._geometry()
        st = _cs()
        self.offsets, self.cnts = self.offsets_all[layer_idx], self.cnts_all[layer_idx]
        vb = self.v_cache_buffer[layer_idx]
        if not g.table_mode:
            check(lib.skv_compact_hits(C.byref(g), 0, vb.data_ptr(), self.offsets.data_ptr(), self.cnts.data_ptr(), st),
                  "compact_hits(V)")
        vh = self.v_cache_cpu[layer_idx]
        vs = self._vstore(layer_idx)
        check(lib.skv_fetch_misses(C.byref(g), vh.data_ptr(), vh.stride(1), vb.data_ptr(), self.offsets.data_ptr(),
                                   self.cnts.data_ptr(), self.position_ids[layer_idx].data_ptr(),
                                   C.byref(vs) if vs is not None else None, st), "fetch_misses")
        return vb[:, :, :self._view_len(layer_idx)]

    def enable_hbm_store(self, ways, fraction: float = 0.5):
        """(Re)allocate the HBM chunk store with `ways` ways per (layer, b, h); 0 removes it.  Any captured
        CUDA graph of the step must be re-captured afterwards.  ways="auto": the largest power of two that fits in
        `fraction` of the HBM that is free right now (`hbm_store_ways_for`)."""
        if ways == "auto":
            self.vstore_tags = self.vstore_data = None          # what an earlier store held counts as free
            free, _ = torch.cuda.mem_get_info(self.device)
            ways = hbm_store_ways_for(free, self.num_layers, self.batch_size * self.num_key_value_heads,
                                      chunks_per_row=self.max_length // self.chunk_size, fraction=fraction,
                                      chunk_bytes=self.head_dim * self.chunk_size * 2)
        if ways & (ways - 1):
            raise ValueError("ways must be 0 or a power of two")
        self.hbm_store_ways = int(ways)
        self.vstore_tags = self.vstore_data = None
        if ways:
            L, BH = self.num_layers, self.batch_size * self.num_key_value_heads
            self.vstore_tags = torch.full((L, BH, ways), -1, dtype=torch.int32, device=self.device)
            self.vstore_data = torch.empty(L, BH, ways, self.head_dim * self.chunk_size, dtype=self.dtype,
                                           device=self.device)
        self.pcie_count.zero_()

    def _vstore(self, layer_idx):
        if not self.hbm_store_ways:
            return None
        return _lib.VStore(tags=self.vstore_tags[layer_idx].data_ptr(), data=self.vstore_data[layer_idx].data_ptr(),
                           pcie_count=self.pcie_count[layer_idx].data_ptr(), ways=self.hbm_store_ways)

    def fill_value_store(self, layer_idx):
        """Install the chunks the last get_value_cache(layer_idx) brought over PCIe into the HBM chunk store.
        Call after get_value_cache on the same stream (decode_layer does); no-op without a store."""
        vs = self._vstore(layer_idx)
        if vs is None:
            return
        g = self._geometry()
        self.offsets, self.cnts = self.offsets_all[layer_idx], self.cnts_all[layer_idx]
        check(lib.skv_vstore_fill(C.byref(g), self.v_cache_buffer[layer_idx].data_ptr(), self.offsets.data_ptr(),
                                  self.cnts.data_ptr(), self.position_ids[layer_idx].data_ptr(), C.byref(vs), _cs()),
              "vstore_fill")

    def join_side_streams(self):
        """Make the current stream wait for everything queued on the copy stream (the store fill of the last
        layer runs behind the attention that did not have to wait for it)."""
        if self._side_used:                   # nothing to join when every layer took the single-launch path
            torch.cuda.current_stream().wait_stream(self.copy_stream)
            self._side_used = False

    def get_key_cache(self, layer_idx, position_ids, rope_func, cos_sin_cache):
        """Compact the hit chunks, rebuild + RoPE the missed ones from U.SV (kv_cache.py:1108-1176).
        rope_func is accepted for signature compatibility and unused, as in the reference."""
        g = self._geometry(cos_sin_cache)
        st = _cs()
        self.offsets, self.cnts = self.offsets_all[layer_idx], self.cnts_all[layer_idx]
        kb = self.k_cache_buffer[layer_idx]
        if not g.table_mode:
            check(lib.skv_compact_hits(C.byref(g), kb.data_ptr(), 0, self.offsets.data_ptr(), self.cnts.data_ptr(), st),
                  "compact_hits(K)")
        check(lib.skv_rebuild_keys(C.byref(g), self.U[layer_idx].data_ptr(), self.SV[layer_idx].data_ptr(),
                                   cos_sin_cache.data_ptr(), self.position_ids[layer_idx].data_ptr(),
                                   self.cnts.data_ptr(), self.offsets.data_ptr(), kb.data_ptr(), self.rebuild_impl, st),
              "rebuild_keys")
        return kb[:, :, :self._view_len(layer_idx)]

    def fetch_and_rebuild(self, layer_idx, cos_sin_cache):
        """get_value_cache + get_key_cache of one layer in ONE launch (the V fetch and the K rebuild of the missed
        chunks run in different CTAs of the same grid), so the attention behind it has a single predecessor."""
        g = self._geometry(cos_sin_cache)
        st = _cs()
        self.offsets, self.cnts = self.offsets_all[layer_idx], self.cnts_all[layer_idx]
        kb, vb = self.k_cache_buffer[layer_idx], self.v_cache_buffer[layer_idx]
        if not g.table_mode:
            check(lib.skv_compact_hits(C.byref(g), kb.data_ptr(), vb.data_ptr(), self.offsets.data_ptr(),
                                       self.cnts.data_ptr(), st), "compact_hits(K,V)")
        vh = self.v_cache_cpu[layer_idx]
        vs = self._vstore(layer_idx)
        check(lib.skv_fetch_and_rebuild(C.byref(g), vh.data_ptr(), vh.stride(1), vb.data_ptr(),
                                        self.U[layer_idx].data_ptr(), self.SV[layer_idx].data_ptr(),
                                        cos_sin_cache.data_ptr(), kb.data_ptr(), self.position_ids[layer_idx].data_ptr(),
                                        self.cnts.data_ptr(), self.offsets.data_ptr(),
                                        C.byref(vs) if vs is not None else None, st), "fetch_and_rebuild")

    def select_fast_active(self) -> bool:
        """Whether the selection launches select_fast_kernel in front of the cluster kernel (csrc/score.cu; an experiment
        that measured slower, enabled only by SKV_SELECT_FAST=1: threshold hints on, at most 16384 landmarks)."""
        n = self.k_landmark.shape[-2] if self.k_landmark is not None else 0
        return (0 < n <= 16384 and os.environ.get("SKV_SELECT_FAST", "0") == "1"
                and os.environ.get("SKV_SELECT_HINT", "1") != "0")

    def can_fuse_attend(self) -> bool:
        """skv_fetch_rebuild_attend handles the in-place chunk table, G <= 8 and the tcgen05 rebuild."""
        return (self.fused_attend and self.table_mode == "inplace" and self.num_key_value_groups <= 8
                and self.rebuild_impl == 0)

    def fetch_rebuild_attend(self, layer_idx, cos_sin_cache, query_states, out=None, kv_len=None, device_counters=False,
                             resident=None, exchange=None):
        """get_value_cache + get_key_cache + flash_attn_with_kvcache of one layer (base.py:326-341) in ONE launch: the CTAs
        that rebuild / fetch the missed chunks also take their softmax partial on chip, other CTAs attend over what is
        already resident while the transfer is in flight.  Returns [B,1,Hq,D].  `resident`: chunk table to use instead of
        position_ids[layer_idx] (kernel benches with constructed hit rates).  `exchange`: _lib.Exchange of the kv-head split
        (the merged rows are also stored into every peer's gathered buffer; parallel.HeadExchange)."""
        g = self._geometry(cos_sin_cache)
        if device_counters:
            kv_len = self.buf_len                    # sizes the grid; the kernel reads the real length
        elif kv_len is None:
            kv_len = self._view_len(layer_idx)
        q = query_states.contiguous()
        B, Hq = q.shape[0], q.shape[1]
        if out is None:
            out = torch.empty(B, 1, Hq, self.head_dim, dtype=self.dtype, device=self.device)
        self.offsets, self.cnts = self.offsets_all[layer_idx], self.cnts_all[layer_idx]
        kb, vb = self.k_cache_buffer[layer_idx], self.v_cache_buffer[layer_idx]
        vh = self.v_cache_cpu[layer_idx]
        vs = self._vstore(layer_idx)
        res = self.position_ids[layer_idx] if resident is None else resident
        check(lib.skv_fetch_rebuild_attend(C.byref(g), vh.data_ptr(), vh.stride(1), vb.data_ptr(),
                                           self.U[layer_idx].data_ptr(), self.SV[layer_idx].data_ptr(),
                                           cos_sin_cache.data_ptr(), kb.data_ptr(), res.data_ptr(), self.cnts.data_ptr(),
                                           self.offsets.data_ptr(), C.byref(vs) if vs is not None else None,
                                           q.data_ptr(), out.data_ptr(), kv_len,
                                           self.kv_len_dev.data_ptr() if device_counters else 0,
                                           1.0 / math.sqrt(self.head_dim), self._attend_ws.data_ptr(),
                                           C.byref(exchange) if exchange is not None else None, _cs()),
              "fetch_rebuild_attend")
        return out

    def attention(self, layer_idx, query_states, kv_len=None, out=None, device_counters=False, programmatic=False):
        """softmax(q K^T / sqrt(D)) V over rows [0, kv_len) of this layer's buffers (base.py:341).
        programmatic=True: programmatic dependent launch behind the single kernel that precedes it in the stream."""
        g = self._geometry()
        if device_counters:
            kv_len = self.buf_len                    # sizes the grid; the kernel reads the real length
        elif kv_len is None:
            kv_len = self._view_len(layer_idx)
        q = query_states.contiguous()
        B, Hq = q.shape[0], q.shape[1]
        if out is None:
            out = torch.empty(B, 1, Hq, self.head_dim, dtype=self.dtype, device=self.device)
        kb, vb = self.k_cache_buffer[layer_idx], self.v_cache_buffer[layer_idx]
        check(lib.skv_attention_decode_ex(q.data_ptr(), kb.data_ptr(), vb.data_ptr(), out.data_ptr(), B,
                                          g.kv_heads, g.q_groups, g.head_dim, kv_len, kb.stride(0), kb.stride(1),
                                          kb.stride(2), 1.0 / math.sqrt(self.head_dim), 0, self._attn_ws.data_ptr(),
                                          self.kv_len_dev.data_ptr() if device_counters else 0,
                                          1 if programmatic else 0, _cs()),
              "attention_decode")
        return out

    def decode_layer(self, layer_idx, query_states, key_states, value_states, cos_sin_cache, out=None,
                     device_counters=False, after_select=None, before_attention=None, exchange=None, wait_for=None,
                     wait_only=False):
        """The whole per-layer sparse-KV decode step, with the reference's stream choreography
        (base.py:320-341): append -> select -> {V fetch on copy_stream || K compact + rebuild} -> attention.
        query_states [B,Hq,1,D], key_states/value_states [B,H,1,D], all post-RoPE.  Returns [B,1,Hq,D].
        With device_counters=True nothing depends on host-side counters, so the call sequence can be
        captured in a CUDA graph and replayed (see reset/advance_device_counters)."""
        self.incoming_q_len = 1
        g = self._geometry(cos_sin_cache)
        # (a2) fused into the scoring launch: same rows, same position as update_kv_cache.  key_states None: the caller
        # appends the token's K / V rows itself before the attention (the whole-model runner's K/V projection writes
        # them in place from a side stream and joins through `before_attention`)
        pos = self.sparse_end + (0 if device_counters else self.gen_offset)
        if pos >= self.buf_len:
            raise RuntimeError(f"generation buffer full ({self.buf_len - self.sparse_end} tokens)")
        ap = None
        if key_states is not None:
            kc, vc = key_states.contiguous(), value_states.contiguous()
            ap = _lib.Append(k_new=kc.data_ptr(), v_new=vc.data_ptr(), k_buf=self.k_cache_buffer[layer_idx].data_ptr(),
                             v_buf=self.v_cache_buffer[layer_idx].data_ptr(),
                             pos_dev=self.gen_dev.data_ptr() if device_counters else None, pos=pos)
        cur, side = torch.cuda.current_stream(), self.copy_stream
        if self.hbm_store_ways and self._side_used and not torch.cuda.is_current_stream_capturing():
            # driven eagerly, the store fill this layer queued on the side stream for the PREVIOUS token reads
            # position_ids / offsets / cnts of this layer and writes its store: it must be complete before this token's
            # selection rewrites those tables and before its fetch reads the store.  (In a captured graph the join at the
            # end of the token body orders them; with several layers the fork / join of the next layer does.)
            cur.wait_stream(side)
        hooks = after_select is not None or before_attention is not None
        single = self.fused_fetch_rebuild and self.rebuild_impl == 0 and not hooks
        fused_attend = single and self.can_fuse_attend()
        if fused_attend and self._row_flags is not None:
            # selection -> fetch + rebuild + attention hand over per (b, kv head) row instead of per grid
            g.row_flags = self._row_flags[layer_idx].data_ptr()
        if wait_for is not None:
            # kv-head split: (flags pointer, count, epoch pointer) of the PREVIOUS layer's exchange; the scoring launch
            # waits for them itself (skv_geom.wait_flags) instead of a flag-wait launch between the layers
            g.wait_flags, g.wait_count, g.wait_epoch_dev = wait_for
            if wait_only:
                # the words are raised by the launch right in front of this layer's scoring launch and cover everything
                # it consumes from it (parallel.RowHandover): wait for them INSTEAD of for that launch's completion
                g.flags |= _lib.GEOM_WAIT_FLAGS_ONLY
        try:
            pid = self.get_retrieval_position_ids(layer_idx, query_states, append=ap)
        except Exception:
            g.row_flags = None
            raise
        finally:
            g.wait_flags, g.wait_count, g.wait_epoch_dev = None, 0, None
            g.flags &= ~_lib.GEOM_WAIT_FLAGS_ONLY
        if after_select is not None:
            # hook of the whole-model runner: from here until the attention the step waits on PCIe and leaves HBM
            # idle; the runner uses the window to pull the next dense weights into L2 on its own stream
            after_select()
        if layer_idx == self.num_layers - 1 and not device_counters:
            self.kv_offset += 1
            self.gen_offset += 1
        kv_len = self.sparse_end + self.gen_offset + (0 if layer_idx == self.num_layers - 1 else 1)
        if fused_attend:
            # everything behind the selection as ONE launch: score -> select -> fetch + rebuild + attention, three kernels
            # per layer, every link a programmatic dependent launch
            try:
                out = self.fetch_rebuild_attend(layer_idx, cos_sin_cache, query_states, out, kv_len, device_counters,
                                                exchange=exchange)
            finally:
                g.row_flags = None
            if self.hbm_store_ways:           # the store fill runs on the side stream, behind this layer's launch
                self._side_used = True
                self._fetch_done.record(cur)
                with torch.cuda.stream(side):
                    side.wait_event(self._fetch_done)
                    self.fill_value_store(layer_idx)
            return out
        if exchange is not None:
            raise RuntimeError("the peer-store exchange rides on the fused fetch + rebuild + attention launch "
                               "(in-place chunk table, G <= 8, no hooks)")
        if single:
            # one launch for both transfers of the missed chunks: select -> fetch+rebuild -> attention is a chain of
            # single predecessors, each launched programmatically (no fork / join, no launch gap before the attention)
            self.fetch_and_rebuild(layer_idx, cos_sin_cache)
        else:
            # two launches on two streams (the reference's choreography).  Also what a caller with hooks gets: it runs
            # its own kernels in the PCIe window and joins them in front of the attention, so the attention has several
            # predecessors anyway, and the rebuild CTAs (96 KiB of shared memory each) leave their SMs after ~9 us
            # instead of staying for the whole transfer as the CTAs of the single launch do.  The V fetch is the long
            # pole (PCIe), so it stays on the current stream right behind the selection kernel; the short K rebuild runs
            # beside it on the side stream (base.py:326-338 has the two swapped, which only matters for who pays the
            # fork / join)
            side.wait_stream(cur)
            self._side_used = True
            with torch.cuda.stream(side):
                self.get_key_cache(layer_idx, pid, None, cos_sin_cache)
                self._k_done.record(side)
            self.get_value_cache(layer_idx, pid)
        if self.hbm_store_ways:               # the store fill runs behind the attention, on the side stream
            self._side_used = True
            self._fetch_done.record(cur)
            with torch.cuda.stream(side):
                side.wait_event(self._fetch_done)
                self.fill_value_store(layer_idx)
        if not single:
            cur.wait_event(self._k_done)
        if before_attention is not None:
            before_attention()
        return self.attention(layer_idx, query_states, kv_len, out, device_counters,
                              programmatic=single and self.programmatic_attention)


class ShadowKVCache(ShadowKVCache_CPU):
    """Batch-1 variant with V resident on the GPU and no chunk cache (reference models/kv_cache.py:155-506):
    every step rebuilds all S chunks.  Same kernels; the resident table is cleared before each selection so
    every chunk is a miss and the V "fetch" reads the device-resident store."""

    def __init__(self, config, batch_size=1, max_length=32 * 1024, device="cuda:0", dtype=torch.bfloat16,
                 sparse_budget=2048, chunk_size=8, rank=160, table_mode="inplace") -> None:
        assert batch_size == 1, "ShadowKV class only supports batch_size=1, please use ShadowKV_CPU class for batch_size > 1"
        super().__init__(config, batch_size, max_length, device, dtype, sparse_budget, chunk_size, rank, table_mode, 0)

    def _outlier_chunks(self) -> int:
        return 48                                                          # kv_cache.py:206

    def _buffer_len(self) -> int:
        return self.sparse_budget + 4096                                   # kv_cache.py:233-251

    def _decide_v_residency(self, want: str) -> str:
        return "hbm"                                                       # GPU-resident full V (kv_cache.py:222-230)

    def print_stats(self):
        print(f"ShadowKV | sparse budget {self.sparse_budget} | chunk size {self.chunk_size} |rank {self.rank} "
              f"| cached {self.kv_offset} | local_chunk {self.local_chunk} | outlier_chunk {self.outlier_chunk}")

    def get_retrieval_position_ids(self, layer_idx, query_states, selected_out=None, score_out=None, append=None):
        self.position_ids[layer_idx].fill_(-1)
        chunk_ids = super().get_retrieval_position_ids(layer_idx, query_states, selected_out, score_out, append)
        # token-level ids, as the reference returns for this class (kv_cache.py:440-445)
        ar = torch.arange(self.chunk_size, device=chunk_ids.device)
        return (chunk_ids.unsqueeze(-1) * self.chunk_size + ar).view(self.batch_size, self.num_key_value_heads, -1)
