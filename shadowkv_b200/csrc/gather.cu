// Chunk-cache data movement: reorder entry point, hit compaction (D2D, in place) and miss fetch
// (pinned host -> device over PCIe).  Integer/byte work, bit-exact against the reference
// (kernels/gather_copy.cu, copy.cuh, map.cuh); HBM/PCIe bound, no tensor cores.
#include <cooperative_groups.h>

#include "chunk_table.cuh"
#include "common.cuh"

namespace cg = cooperative_groups;

namespace skv {

// ------------------------------------------------------------------------------------------------
// reorder_keys_and_compute_offsets
// ------------------------------------------------------------------------------------------------
template <int S>
__global__ void __launch_bounds__(S) reorder_kernel(int64_t *__restrict__ cached, const int64_t *__restrict__ cur,
                                                    int32_t *__restrict__ offsets, int32_t *__restrict__ cnts) {
  __shared__ int s_sorted[S];
  __shared__ int s_slot[S];
  __shared__ int s_warp[33];
  const int row = blockIdx.x, tid = threadIdx.x;
  const int old_id = static_cast<int>(cached[size_t(row) * S + tid]);   // int32 truncation as map.cuh:771
  const int new_id = static_cast<int>(cur[size_t(row) * S + tid]);      // map.cuh:776
  ChunkTableSmem sm{s_sorted, s_slot, s_warp};
  chunk_table_update<S>(tid, old_id, new_id, sm, cached + size_t(row) * S, offsets + size_t(row) * S, cnts + row);
}

// ------------------------------------------------------------------------------------------------
// Hit compaction.  One cluster of 8 CTAs per (b,h) and buffer; each CTA stages up to 32 chunks
// (64 KiB) in shared memory with TMA bulk copies, the cluster barrier separates "everyone has read"
// from "anyone may write", so the in-place forward move needs no temp buffer and no inter-kernel
// handshake.  Rounds of 256 slots keep it valid for S = 512 / 1024 (offsets[i] >= i and ascending,
// so slots written in round r are never read in a later round).
// ------------------------------------------------------------------------------------------------
constexpr int kCompactCluster = 8;
constexpr int kCompactPerCta = 32;
constexpr int kCompactRound = kCompactCluster * kCompactPerCta;   // 256 slots

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *gdst, const void *smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
}

__global__ void __cluster_dims__(kCompactCluster, 1, 1) __launch_bounds__(128)
    compact_hits_kernel(skv_bf16 *__restrict__ buf0, skv_bf16 *__restrict__ buf1, const int32_t *__restrict__ offsets,
                        const int32_t *__restrict__ cnts, int S, int64_t stride_bh, int64_t base_off) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char *stage = smem_raw;                                              // 32 x 2 KiB
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + kCompactPerCta * kChunkBytes);
  int *s_off = reinterpret_cast<int *>(bar + 1);                                // [32]

  cg::cluster_group cluster = cg::this_cluster();
  const int rank = blockIdx.x;                 // cluster spans blockIdx.x
  const int bh = blockIdx.y;
  skv_bf16 *buf = blockIdx.z == 0 ? buf0 : buf1;
  const int cnt = cnts[bh];
  unsigned char *base = reinterpret_cast<unsigned char *>(buf + size_t(bh) * stride_bh + base_off);
  const int32_t *off = offsets + size_t(bh) * S;

  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int rounds = (cnt + kCompactRound - 1) / kCompactRound;
  uint32_t parity = 0;
  for (int r = 0; r < rounds; ++r) {
    const int i0 = r * kCompactRound + rank * kCompactPerCta;
    const int n = max(0, min(kCompactPerCta, cnt - i0));
    if (threadIdx.x < kCompactPerCta) s_off[threadIdx.x] = threadIdx.x < n ? off[i0 + threadIdx.x] : -1;
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t bytes = 0;
      for (int j = 0; j < n; ++j) bytes += (s_off[j] != i0 + j) ? kChunkBytes : 0;
      if (bytes) {
        mbar_expect_tx(bar, bytes);
        for (int j = 0; j < n; ++j)
          if (s_off[j] != i0 + j)
            bulk_g2s(stage + j * kChunkBytes, base + size_t(s_off[j]) * kChunkBytes, kChunkBytes, bar);
        mbar_wait(bar, parity);
        parity ^= 1;
      }
    }
    cluster.sync();   // every CTA of this (b,h) has finished reading its sources for this round
    if (threadIdx.x == 0) {
      bool any = false;
      for (int j = 0; j < n; ++j)
        if (s_off[j] != i0 + j) {
          bulk_s2g(base + size_t(i0 + j) * kChunkBytes, stage + j * kChunkBytes, kChunkBytes);
          any = true;
        }
      if (any) {
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // staging reusable next round
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // writes complete before exit
}

static int launch_compact(skv_bf16 *buf0, skv_bf16 *buf1, const int32_t *offsets, const int32_t *cnts, int BH, int S,
                          int64_t stride_bh, int64_t base_off, cudaStream_t st) {
  if (!buf0 && !buf1) return SKV_OK;
  if (!buf0) { buf0 = buf1; buf1 = nullptr; }
  const int smem = kCompactPerCta * kChunkBytes + 8 + kCompactPerCta * 4 + 16;
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void *>(compact_hits_kernel), smem, "compact_hits_kernel")) return rc;
  dim3 grid(kCompactCluster, BH, buf1 ? 2 : 1);
  compact_hits_kernel<<<grid, 128, smem, st>>>(buf0, buf1, offsets, cnts, S, stride_bh, base_off);
  return launch_status("compact_hits_kernel");
}

// ------------------------------------------------------------------------------------------------
// Chunk gather from the (pinned host or device) chunk-major store.  One warp moves one 2 KiB chunk
// pair with 8 x 16 B loads in flight per lane; only slots >= cnt are touched.
// ------------------------------------------------------------------------------------------------
template <typename IdT>
__global__ void __launch_bounds__(256)
    fetch_chunks_kernel(const skv_bf16 *__restrict__ src, int64_t src_stride_bh, skv_bf16 *__restrict__ dst,
                        int64_t dst_stride_bh, int64_t dst_base_off, const IdT *__restrict__ ids,
                        const int32_t *__restrict__ cnts, const int64_t *__restrict__ resident, int S,
                        skv_vstore vs) {
  pdl_launch_dependents();
  pdl_wait();                                         // offsets / cnts / resident come from the selection kernel
  const int bh = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cnt = cnts ? cnts[bh] : 0;
  const int i = (blockIdx.x * 8 + warp) * 2;          // two slots (or two entries of the miss list) per warp
  const int4 *sbase = reinterpret_cast<const int4 *>(src + size_t(bh) * src_stride_bh);
  int4 *dbase = reinterpret_cast<int4 *>(dst + size_t(bh) * dst_stride_bh + dst_base_off);
  const IdT *id = ids + size_t(bh) * S;
  int4 v[2][4];
  bool on[2];
  int slot[2];
  if (resident) {                                     // in-place table: ids[] lists the slots of the S-cnt misses
    if (i >= S - cnt) return;
  } else if (i + 1 < cnt || i >= S) {
    return;
  }
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    long long chunk = 0;
    if (resident) {
      on[u] = (i + u) < S - cnt;
      if (on[u]) { slot[u] = static_cast<int>(id[i + u]); chunk = resident[size_t(bh) * S + slot[u]]; }
    } else {
      on[u] = (i + u) >= cnt && (i + u) < S;
      slot[u] = i + u;
      if (on[u]) chunk = static_cast<long long>(id[i + u]);
    }
    if (on[u]) {
      const int4 *p = sbase + size_t(chunk) * (kChunkBytes / 16) + lane;
      if (vs.tags) {                                  // HBM chunk store: read-only here (skv_vstore_fill updates it)
        const int way = static_cast<int>(chunk) & (vs.ways - 1);
        const size_t w = size_t(bh) * vs.ways + way;
        if (vs.tags[w] == static_cast<int>(chunk))    // warp uniform
          p = reinterpret_cast<const int4 *>(vs.data) + w * (kChunkBytes / 16) + lane;
        else if (vs.pcie_count && lane == 0)
          atomicAdd(vs.pcie_count + bh, 1);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) v[u][k] = ld_stream(p + k * 32);
    }
  }
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    if (on[u]) {
      int4 *p = dbase + size_t(slot[u]) * (kChunkBytes / 16) + lane;
#pragma unroll
      for (int k = 0; k < 4; ++k) p[k * 32] = v[u][k];
    }
  }
}

template <typename IdT>
static int launch_fetch(const skv_bf16 *src, int64_t src_stride, skv_bf16 *dst, int64_t dst_stride, int64_t dst_off,
                        const IdT *ids, const int32_t *cnts, const int64_t *resident, int BH, int S, cudaStream_t st,
                        const skv_vstore *vstore = nullptr) {
  skv_vstore vs{};
  if (vstore) vs = *vstore;
  dim3 grid((S + 15) / 16, BH);
  cudaError_t e = launch_kernel(fetch_chunks_kernel<IdT>, grid, dim3(256), 0, st, /*pdl=*/true, src, src_stride, dst,
                                dst_stride, dst_off, ids, cnts, resident, S, vs);
  if (e != cudaSuccess) { set_error("fetch_chunks_kernel: %s", cudaGetErrorString(e)); return int(e); }
  return launch_status("fetch_chunks_kernel");
}

// HBM chunk-store fill: one CTA per (b,h).  Thread t looks at miss t: it installs its chunk when the store does
// not hold it yet and no LATER miss of this row maps to the same way (so every way has exactly one writer and the
// fetch kernel, which only reads the store, never races with a fill).
__global__ void __launch_bounds__(256)
    vstore_fill_kernel(const skv_bf16 *__restrict__ v_buf, int64_t stride_bh, int64_t base_off,
                       const int32_t *__restrict__ offsets, const int32_t *__restrict__ cnts,
                       const int64_t *__restrict__ resident, int S, int table_mode, skv_vstore vs) {
  __shared__ int s_way[1024], s_slot[1024], s_chunk[1024];
  __shared__ int s_win[1024], s_nwin;
  const int bh = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cnt = cnts[bh], n_miss = S - cnt;
  if (threadIdx.x == 0) s_nwin = 0;
  for (int i = threadIdx.x; i < n_miss; i += blockDim.x) {
    int slot, chunk;
    if (table_mode) { slot = offsets[size_t(bh) * S + i]; chunk = static_cast<int>(resident[size_t(bh) * S + slot]); }
    else { slot = cnt + i; chunk = offsets[size_t(bh) * S + slot]; }
    s_slot[i] = slot; s_chunk[i] = chunk; s_way[i] = chunk & (vs.ways - 1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_miss; i += blockDim.x) {
    const int way = s_way[i];
    bool win = vs.tags[size_t(bh) * vs.ways + way] != s_chunk[i];      // not in the store yet
    for (int j = i + 1; win && j < n_miss; ++j) win = s_way[j] != way;   // the last miss of a way wins
    if (win) s_win[atomicAdd(&s_nwin, 1)] = i;
  }
  __syncthreads();
  const int nwin = s_nwin;
  const int4 *src = reinterpret_cast<const int4 *>(v_buf + size_t(bh) * stride_bh + base_off);
  int4 *dst = reinterpret_cast<int4 *>(vs.data) + size_t(bh) * vs.ways * (kChunkBytes / 16);
  for (int k = warp; k < nwin; k += blockDim.x >> 5) {
    const int i = s_win[k];
    const int4 *p = src + size_t(s_slot[i]) * (kChunkBytes / 16) + lane;
    int4 *q = dst + size_t(s_way[i]) * (kChunkBytes / 16) + lane;
    int4 v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) v[e] = p[e * 32];
#pragma unroll
    for (int e = 0; e < 4; ++e) q[e * 32] = v[e];
    if (lane == 0) vs.tags[size_t(bh) * vs.ways + s_way[i]] = s_chunk[i];
  }
}

}  // namespace skv

using namespace skv;

extern "C" {

int skv_reorder_keys_and_compute_offsets(int64_t *cached_pos_ids, const int64_t *cur_pos_ids, int32_t *offsets,
                                         int32_t *cnts, int batch_size, int heads, int map_size,
                                         skv_stream_t stream) {
  if (!cached_pos_ids || !cur_pos_ids || !offsets || !cnts) { set_error("reorder: null pointer"); return SKV_ERR_ARGUMENT; }
  const int rows = batch_size * heads;
  if (rows <= 0) return SKV_OK;
  cudaStream_t st = as_stream(stream);
  switch (map_size) {
    case 128: reorder_kernel<128><<<rows, 128, 0, st>>>(cached_pos_ids, cur_pos_ids, offsets, cnts); break;
    case 256: reorder_kernel<256><<<rows, 256, 0, st>>>(cached_pos_ids, cur_pos_ids, offsets, cnts); break;
    case 512: reorder_kernel<512><<<rows, 512, 0, st>>>(cached_pos_ids, cur_pos_ids, offsets, cnts); break;
    case 1024: reorder_kernel<1024><<<rows, 1024, 0, st>>>(cached_pos_ids, cur_pos_ids, offsets, cnts); break;
    default:
      set_error("reorder: map_size %d not in {128,256,512,1024}", map_size);
      return SKV_ERR_UNSUPPORTED;
  }
  return launch_status("reorder_kernel");
}

int skv_gather_copy(const skv_bf16 *values, skv_bf16 *v_cache_buffer, const int64_t *position_ids, int batch_size,
                    int heads, int cpu_v_length, int gpu_v_length, int map_size, skv_stream_t stream) {
  if (!values || !v_cache_buffer || !position_ids) { set_error("gather_copy: null pointer"); return SKV_ERR_ARGUMENT; }
  if (!valid_map_size(map_size)) { set_error("gather_copy: map_size %d unsupported", map_size); return SKV_ERR_UNSUPPORTED; }
  return launch_fetch<int64_t>(values, cpu_v_length, v_cache_buffer, gpu_v_length, 0, position_ids, nullptr, nullptr,
                               batch_size * heads, map_size, as_stream(stream));
}

int skv_gather_chunks(const skv_bf16 *src, int64_t src_stride, skv_bf16 *dst, int64_t dst_stride, int64_t dst_offset,
                      const int64_t *ids, int n_ids, int rows, skv_stream_t stream) {
  if (!src || !dst || !ids) { set_error("gather_chunks: null pointer"); return SKV_ERR_ARGUMENT; }
  if (n_ids < 1 || rows < 1) return SKV_OK;
  return launch_fetch<int64_t>(src, src_stride, dst, dst_stride, dst_offset, ids, nullptr, nullptr, rows, n_ids,
                               as_stream(stream));
}

int skv_gather_copy_d2d_with_offsets(skv_bf16 *keys, const int32_t *offsets, const int32_t *cnts, int batch_size,
                                     int heads, int gpu_k_length, int gpu_k_offset, int gpu_k_stride, int map_size,
                                     skv_stream_t stream) {
  (void)gpu_k_length;
  if (!keys || !offsets || !cnts) { set_error("gather_copy_d2d: null pointer"); return SKV_ERR_ARGUMENT; }
  if (!valid_map_size(map_size)) { set_error("gather_copy_d2d: map_size %d unsupported", map_size); return SKV_ERR_UNSUPPORTED; }
  return launch_compact(keys, nullptr, offsets, cnts, batch_size * heads, map_size, gpu_k_stride, gpu_k_offset,
                        as_stream(stream));
}

int skv_gather_copy_with_offsets(const skv_bf16 *values, skv_bf16 *v_cache_buffer, skv_bf16 *temp,
                                 const int32_t *offsets, const int32_t *cnts, int32_t *signals, int batch_size,
                                 int heads, int cpu_v_length, int gpu_v_length, int gpu_v_offset, int gpu_v_stride,
                                 int map_size, skv_stream_t stream) {
  (void)temp; (void)signals; (void)gpu_v_length;
  if (!values || !v_cache_buffer || !offsets || !cnts) { set_error("gather_copy_with_offsets: null pointer"); return SKV_ERR_ARGUMENT; }
  if (!valid_map_size(map_size)) { set_error("gather_copy_with_offsets: map_size %d unsupported", map_size); return SKV_ERR_UNSUPPORTED; }
  cudaStream_t st = as_stream(stream);
  int rc = launch_compact(v_cache_buffer, nullptr, offsets, cnts, batch_size * heads, map_size, gpu_v_stride,
                          gpu_v_offset, st);
  if (rc) return rc;
  return launch_fetch<int32_t>(values, cpu_v_length, v_cache_buffer, gpu_v_stride, gpu_v_offset, offsets, cnts, nullptr,
                               batch_size * heads, map_size, st);
}

int skv_compact_hits(const skv_geom *g, skv_bf16 *k_buf, skv_bf16 *v_buf, const int32_t *offsets,
                     const int32_t *cnts, skv_stream_t stream) {
  if (!g || !offsets || !cnts) { set_error("compact_hits: null pointer"); return SKV_ERR_ARGUMENT; }
  if (g->head_dim != kHeadDim || g->chunk != kChunkTokens || !valid_map_size(g->select)) {
    set_error("compact_hits: unsupported geometry"); return SKV_ERR_UNSUPPORTED;
  }
  return launch_compact(k_buf, v_buf, offsets, cnts, g->batch * g->kv_heads, g->select,
                        int64_t(g->buf_len) * g->head_dim, int64_t(g->sparse_start) * g->head_dim, as_stream(stream));
}

int skv_fetch_misses(const skv_geom *g, const skv_bf16 *v_host, int64_t host_stride, skv_bf16 *v_buf,
                     const int32_t *offsets, const int32_t *cnts, const int64_t *resident, const skv_vstore *vstore,
                     skv_stream_t stream) {
  if (!g || !v_host || !v_buf || !offsets || !cnts || (g->table_mode && !resident)) {
    set_error("fetch_misses: null pointer"); return SKV_ERR_ARGUMENT;
  }
  if (vstore && (!vstore->tags || !vstore->data || vstore->ways < 1 || (vstore->ways & (vstore->ways - 1)))) {
    set_error("fetch_misses: chunk store needs tags, data and a power-of-two way count"); return SKV_ERR_ARGUMENT;
  }
  if (g->head_dim != kHeadDim || g->chunk != kChunkTokens || !valid_map_size(g->select)) {
    set_error("fetch_misses: unsupported geometry"); return SKV_ERR_UNSUPPORTED;
  }
  return launch_fetch<int32_t>(v_host, host_stride, v_buf, int64_t(g->buf_len) * g->head_dim,
                               int64_t(g->sparse_start) * g->head_dim, offsets, cnts,
                               g->table_mode ? resident : nullptr, g->batch * g->kv_heads, g->select,
                               as_stream(stream), vstore);
}

int skv_vstore_fill(const skv_geom *g, const skv_bf16 *v_buf, const int32_t *offsets, const int32_t *cnts,
                    const int64_t *resident, const skv_vstore *vstore, skv_stream_t stream) {
  if (!g || !v_buf || !offsets || !cnts || !vstore || !vstore->tags || !vstore->data || (g->table_mode && !resident)) {
    set_error("vstore_fill: null pointer"); return SKV_ERR_ARGUMENT;
  }
  if (g->head_dim != kHeadDim || g->chunk != kChunkTokens || !valid_map_size(g->select) || vstore->ways < 1 ||
      (vstore->ways & (vstore->ways - 1))) {
    set_error("vstore_fill: unsupported geometry"); return SKV_ERR_UNSUPPORTED;
  }
  vstore_fill_kernel<<<g->batch * g->kv_heads, 256, 0, as_stream(stream)>>>(
      v_buf, int64_t(g->buf_len) * g->head_dim, int64_t(g->sparse_start) * g->head_dim, offsets, cnts, resident,
      g->select, g->table_mode, *vstore);
  return launch_status("vstore_fill_kernel");
}

}  // extern "C"
