// Low-rank pre-RoPE key rebuild for the missed chunks:  K[r,:] = U[b, id*8 + r%8, :R] . SV[b,h,:, :R]^T
// (reference kernels/batch_gather_gemm.cu:193-287 + rope_new.cu:321-370/429-493, three launches and a bf16
// scratch round trip there).  Here: one launch.  The chunk-gathered U rows (8 contiguous rows of 2R bytes per
// chunk) are written into 128-byte-swizzled K-major slabs by 16-byte cp.async from all threads, the SV panel
// comes by TMA, the 128x128xR product runs on tcgen05 with the fp32 accumulator in TMEM, and the epilogue rounds to
// bf16, applies the rotary embedding in bf16 at token position id*8+r and writes straight into the K buffer.
#include <map>
#include <mutex>
#include <tuple>

#include "sm100_ptx.cuh"

namespace skv {

// bf16 x bf16 -> f32, A and B K-major, M=128, N=128
constexpr uint32_t kIdescBf16M128N128 = umma_idesc_bf16_m128(128);

// ------------------------------------------------------------------------------------------------
constexpr int kTileRows = 128;                 // 16 chunks
constexpr int kTileChunks = kTileRows / kChunkTokens;
constexpr int kSlabCols = 64;                  // 128 B of bf16 = one swizzle span
constexpr int kMaxSlabs = 4;                   // rank <= 256
constexpr int kSlabBytesA = kTileRows * 128;   // 16 KiB
constexpr int kSlabBytesB = kHeadDim * 128;    // 16 KiB

struct RebuildParams {
  const int64_t *resident;   // [BH, S] chunk ids (after reorder)
  const int32_t *ids32;      // alternative int32 ids (drop-in batch_gather_gemm); one of the two is set
  const int32_t *cnts;       // [BH]
  const __nv_bfloat16 *U;    // [B, L, R]: rows of the missed chunks are gathered with cp.async by all threads
  const int32_t *miss_slots; // in-place table: [BH, S] slot of the i-th miss (i < S - cnt); null = slots cnt..S-1
  const __nv_bfloat16 *cos_sin;
  __nv_bfloat16 *out;        // K buffer (fused) or scratch [BH, S*8, 128] (drop-in)
  int64_t out_stride_bh;     // elements
  int64_t out_row0;          // first row of slot 0 inside a (b,h) block
  int H, S, L, R, n_slabs, k_steps;
  int rope;                  // 0 none (drop-in pre-RoPE), 1 neox, 2 glm
  int tile_skip;             // 1: skip by whole tiles like the reference and write every row of a kept tile
};

// One 128-row tile (16 missed chunks) of one (b,h), worked on by threads 0..127 of the CTA.  Shared by the stand-alone
// rebuild kernel and by the rebuild warps of fetch_rebuild_kernel.
// barrier of the 128 threads that work on a tile (the whole CTA of rebuild_tcgen05_kernel, the first four warps of
// fetch_rebuild_kernel)
__device__ __forceinline__ void tile_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

__device__ __forceinline__ void rebuild_tile(const CUtensorMap *map_sv_ptr, const RebuildParams &p, const int tile,
                                             const int bh, unsigned char *smem) {
  unsigned char *sA = smem;                                   // n_slabs x 16 KiB
  // sized by the slabs the rank needs (rank 160 -> 3 slabs -> 96 KiB, two CTAs per SM), not by the maximum
  unsigned char *sB = smem + p.n_slabs * kSlabBytesA;         // n_slabs x 16 KiB
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + p.n_slabs * (kSlabBytesA + kSlabBytesB));
  uint64_t *mma_bar = full_bar + kMaxSlabs;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(mma_bar + 1);
  int *s_ids = reinterpret_cast<int *>(tmem_slot + 2);        // [16] chunk id per tile row group (-1: skip)
  int *s_slots = s_ids + kTileChunks;                         // [16] destination slot

  const int b = bh / p.H;
  const int warp = threadIdx.x >> 5;
  const int cnt = p.cnts ? p.cnts[bh] : 0;
  const int first_slot = tile * kTileChunks;
  if (p.miss_slots) {
    if (first_slot >= p.S - cnt) return;                      // tiles run over the miss list in this mode
  } else if (first_slot + kTileChunks <= cnt) {
    return;                                                   // tile entirely cache hits (uniform per CTA)
  }

  if (threadIdx.x < kTileChunks) {
    int slot = first_slot + threadIdx.x;
    int id = -1;
    if (p.miss_slots) {
      if (slot < p.S - cnt) {
        slot = p.miss_slots[size_t(bh) * p.S + slot];
        id = static_cast<int>(p.resident[size_t(bh) * p.S + slot]);
      }
    } else if (slot < p.S && (p.tile_skip || slot >= cnt)) {
      id = p.resident ? static_cast<int>(p.resident[size_t(bh) * p.S + slot]) : p.ids32[size_t(bh) * p.S + slot];
    }
    s_ids[threadIdx.x] = id;
    s_slots[threadIdx.x] = slot;
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.n_slabs; ++s) mbarrier_init(&full_bar[s], 1);
    mbarrier_init(mma_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(map_sv_ptr) : "memory");
  }
  if (warp == 0) tcgen05_alloc(tmem_slot, 128);
  tcgen05_fence_before();
  tile_sync();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // U rows of the tile's chunks: 128 threads x 16-byte cp.async into the 128-byte-swizzled K-major layout the MMA
  // descriptor expects (piece c of row r lands at piece c ^ (r & 7) of its 128-byte line).  One TMA box per chunk
  // and slab (48 serial issues by one thread) was the longest part of this kernel; here every load is in flight
  // at once.  The SV panel still comes by TMA (three 16 KiB boxes).
  {
    const int ppr = p.R >> 3;                                  // 16-byte pieces per U row
    const __nv_bfloat16 *ub = p.U + size_t(b) * p.L * p.R;
    for (int q = threadIdx.x; q < kTileRows * ppr; q += 128) {
      const int r = q / ppr, c = q - r * ppr;
      const int id = s_ids[r >> 3];
      if (id >= 0) {
        const __nv_bfloat16 *src = ub + (size_t(id) * kChunkTokens + (r & 7)) * p.R + c * 8;
        const uint32_t dst = smem_addr(sA + (c >> 3) * kSlabBytesA + (r >> 3) * 1024 + (r & 7) * 128 + (((c & 7) ^ (r & 7)) << 4));
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
      }
    }
    if (p.R & 8) {                                             // last 16-column MMA step is half valid: zero the rest
      const int r = threadIdx.x, c = ppr;
      *reinterpret_cast<int4 *>(sA + (c >> 3) * kSlabBytesA + (r >> 3) * 1024 + (r & 7) * 128 + (((c & 7) ^ (r & 7)) << 4)) =
          make_int4(0, 0, 0, 0);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.n_slabs; ++s) {
      mbarrier_expect_tx(&full_bar[s], kSlabBytesB);
      tma_load_2d(sB + s * kSlabBytesB, map_sv_ptr, s * kSlabCols, bh * kHeadDim, &full_bar[s]);
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  fence_proxy_async_smem();                                    // cp.async wrote through the generic proxy
  tile_sync();
  if (threadIdx.x == 0) {
    // MMA issuer
    int step = 0;
    for (int s = 0; s < p.n_slabs; ++s) {
      mbarrier_wait(&full_bar[s], 0);
      tcgen05_fence_after();
      const uint64_t da = umma_desc_k_sw128(smem_addr(sA + s * kSlabBytesA));
      const uint64_t db = umma_desc_k_sw128(smem_addr(sB + s * kSlabBytesB));
      for (int kk = 0; kk < 4 && step < p.k_steps; ++kk, ++step)
        tcgen05_mma_bf16(tmem_base, da + uint64_t(kk * 2), db + uint64_t(kk * 2), kIdescBf16M128N128, step > 0);
    }
    tcgen05_commit(mma_bar);
  }
  __syncwarp();
  mbarrier_wait(mma_bar, 0);
  tcgen05_fence_after();

  // ---- epilogue: thread = tile row ---------------------------------------------------------------
  const int row = threadIdx.x;                     // warp w owns TMEM lanes 32w..32w+31
  const int j = row >> 3, r = row & 7;
  const int id = s_ids[j];
  const int slot = s_slots[j];
  const bool write = id >= 0;
  const uint32_t taddr = tmem_base + (uint32_t(warp * 32) << 16);
  __nv_bfloat16 *orow = p.out + size_t(bh) * p.out_stride_bh + (p.out_row0 + size_t(slot) * kChunkTokens + r) * kHeadDim;
  const long long token = (long long)id * kChunkTokens + r;

#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
    uint32_t lo[32], hi[32];
    tcgen05_ld32(taddr + half * 32, lo);            // columns [32*half, 32*half+32)
    tcgen05_ld32(taddr + 64 + half * 32, hi);       // columns [64+32*half, ...)
    tcgen05_wait_ld();
    if (!write) continue;
    uint32_t olo[16], ohi[16];
    if (p.rope == 1) {
      const __nv_bfloat16 *cs = p.cos_sin + token * kHeadDim;
      const uint4 *c4 = reinterpret_cast<const uint4 *>(cs + half * 32);
      const uint4 *s4 = reinterpret_cast<const uint4 *>(cs + 64 + half * 32);
      uint32_t cw[16], sw[16];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint4 a = c4[q], bq = s4[q];
        cw[4 * q] = a.x; cw[4 * q + 1] = a.y; cw[4 * q + 2] = a.z; cw[4 * q + 3] = a.w;
        sw[4 * q] = bq.x; sw[4 * q + 1] = bq.y; sw[4 * q + 2] = bq.z; sw[4 * q + 3] = bq.w;
      }
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const __nv_bfloat162 x1 = __floats2bfloat162_rn(__uint_as_float(lo[2 * e]), __uint_as_float(lo[2 * e + 1]));
        const __nv_bfloat162 x2 = __floats2bfloat162_rn(__uint_as_float(hi[2 * e]), __uint_as_float(hi[2 * e + 1]));
        const __nv_bfloat162 c = *reinterpret_cast<const __nv_bfloat162 *>(&cw[e]);
        const __nv_bfloat162 s = *reinterpret_cast<const __nv_bfloat162 *>(&sw[e]);
        const __nv_bfloat162 o1 = __hfma2(x1, c, __hmul2(__hneg2(x2), s));
        const __nv_bfloat162 o2 = __hfma2(x2, c, __hmul2(x1, s));
        olo[e] = *reinterpret_cast<const uint32_t *>(&o1);
        ohi[e] = *reinterpret_cast<const uint32_t *>(&o2);
      }
    } else if (p.rope == 2 ) {
      // GLM: interleaved pairs over dims [0,64) with [cos32|sin32]; dims [64,128) copied.
      // lo holds dims [32*half, 32*half+32) = pairs [16*half, 16*half+16)
      const __nv_bfloat16 *cs = p.cos_sin + token * (kHeadDim / 2);
      const uint4 *c4 = reinterpret_cast<const uint4 *>(cs + half * 16);
      const uint4 *s4 = reinterpret_cast<const uint4 *>(cs + 32 + half * 16);
      uint32_t cw[8], sw[8];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const uint4 a = c4[q], bq = s4[q];
        cw[4 * q] = a.x; cw[4 * q + 1] = a.y; cw[4 * q + 2] = a.z; cw[4 * q + 3] = a.w;
        sw[4 * q] = bq.x; sw[4 * q + 1] = bq.y; sw[4 * q + 2] = bq.z; sw[4 * q + 3] = bq.w;
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {         // two pairs per iteration: elements 4e..4e+3
        const __nv_bfloat162 xe = __floats2bfloat162_rn(__uint_as_float(lo[4 * e]), __uint_as_float(lo[4 * e + 2]));
        const __nv_bfloat162 xo = __floats2bfloat162_rn(__uint_as_float(lo[4 * e + 1]), __uint_as_float(lo[4 * e + 3]));
        const __nv_bfloat162 c = *reinterpret_cast<const __nv_bfloat162 *>(&cw[e]);
        const __nv_bfloat162 s = *reinterpret_cast<const __nv_bfloat162 *>(&sw[e]);
        const __nv_bfloat162 oe = __hfma2(xe, c, __hmul2(__hneg2(xo), s));
        const __nv_bfloat162 oo = __hfma2(xo, c, __hmul2(xe, s));
        olo[2 * e] = pack_bf16(__low2bfloat16(oe), __low2bfloat16(oo));
        olo[2 * e + 1] = pack_bf16(__high2bfloat16(oe), __high2bfloat16(oo));
      }
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const __nv_bfloat162 x2 = __floats2bfloat162_rn(__uint_as_float(hi[2 * e]), __uint_as_float(hi[2 * e + 1]));
        ohi[e] = *reinterpret_cast<const uint32_t *>(&x2);
      }
    } else {
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const __nv_bfloat162 x1 = __floats2bfloat162_rn(__uint_as_float(lo[2 * e]), __uint_as_float(lo[2 * e + 1]));
        const __nv_bfloat162 x2 = __floats2bfloat162_rn(__uint_as_float(hi[2 * e]), __uint_as_float(hi[2 * e + 1]));
        olo[e] = *reinterpret_cast<const uint32_t *>(&x1);
        ohi[e] = *reinterpret_cast<const uint32_t *>(&x2);
      }
    }
    uint4 *d1 = reinterpret_cast<uint4 *>(orow + half * 32);
    uint4 *d2 = reinterpret_cast<uint4 *>(orow + 64 + half * 32);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      d1[q] = make_uint4(olo[4 * q], olo[4 * q + 1], olo[4 * q + 2], olo[4 * q + 3]);
      d2[q] = make_uint4(ohi[4 * q], ohi[4 * q + 1], ohi[4 * q + 2], ohi[4 * q + 3]);
    }
  }
  tcgen05_fence_before();
  tile_sync();
  if (warp == 0) tcgen05_dealloc(tmem_base, 128);
}

__global__ void __launch_bounds__(128, 2)
    rebuild_tcgen05_kernel(const __grid_constant__ CUtensorMap map_sv, RebuildParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  rebuild_tile(&map_sv, p, blockIdx.x, blockIdx.y, smem);
}

// ------------------------------------------------------------------------------------------------
// The two transfers a missed chunk needs, in ONE launch: its V rows from the pinned host store (PCIe, or the HBM
// chunk store) and its K rows rebuilt from the low-rank factors.  A CTA owns 16 entries of the miss list (one rebuild
// tile): warps 0-3 rebuild their K rows exactly as rebuild_tcgen05_kernel does, warps 4-7 fetch their V chunks, four
// chunks per warp with all 16 x 16-byte loads of a lane in flight before the first store.  Grid, shared memory and
// TMEM footprint are those of the rebuild kernel alone; the fetch warps only add registers.
// Why one launch: with two launches on two streams the attention that follows has two predecessors and cannot be
// launched programmatically; here the chain select -> fetch+rebuild -> attention -> next scoring has single
// predecessors all the way (reference stream choreography: models/base.py:326-338).
// ------------------------------------------------------------------------------------------------
struct FetchParams {
  const skv_bf16 *src;       // pinned host (or device) chunk-major store
  int64_t src_stride_bh;     // elements
  skv_bf16 *dst;             // V buffer
  int64_t dst_stride_bh, dst_base_off;
  const int32_t *offsets;    // reference table: chunk id of slot i >= cnt; in-place table: slot of the i-th miss
  const int64_t *resident;   // in-place table only: chunk id per slot
  skv_vstore vs;
};

constexpr int kFetchPerWarp = 4;

__device__ __forceinline__ void fetch_entries(const FetchParams &f, const int32_t *cnts, const int S, const int block,
                                              const int bh, const int warp) {
  const int lane = threadIdx.x & 31;
  const int i = (block * 4 + warp) * kFetchPerWarp;   // i + kFetchPerWarp <= S: the grid covers S entries exactly
  const int32_t *off = f.offsets + size_t(bh) * S;
  // The index chain (hit count | miss list -> slot -> chunk id -> store tag) for all four chunks at once: every level
  // is one batch of independent loads, so its latency is paid once per warp and not once per chunk.  The list entries
  // are read before the hit count is known (entries beyond the misses hold stale but in-range values).
  int raw[kFetchPerWarp];
#pragma unroll
  for (int u = 0; u < kFetchPerWarp; ++u) raw[u] = off[i + u];
  const int cnt = cnts[bh];
  if (f.resident) {
    if (i >= S - cnt) return;
  } else if (i + kFetchPerWarp <= cnt) {
    return;
  }
  const int4 *sbase = reinterpret_cast<const int4 *>(f.src + size_t(bh) * f.src_stride_bh);
  int4 *dbase = reinterpret_cast<int4 *>(f.dst + size_t(bh) * f.dst_stride_bh + f.dst_base_off);
  bool on[kFetchPerWarp];
  int slot[kFetchPerWarp];
  long long chunk[kFetchPerWarp];
  if (f.resident) {
#pragma unroll
    for (int u = 0; u < kFetchPerWarp; ++u) {
      on[u] = (i + u) < S - cnt;
      slot[u] = raw[u] & (S - 1);                     // S is a power of two; a no-op for the entries that count
      chunk[u] = f.resident[size_t(bh) * S + slot[u]];
    }
  } else {
#pragma unroll
    for (int u = 0; u < kFetchPerWarp; ++u) {
      on[u] = (i + u) >= cnt;
      slot[u] = i + u;
      chunk[u] = raw[u];
    }
  }
  const int4 *src[kFetchPerWarp];
#pragma unroll
  for (int u = 0; u < kFetchPerWarp; ++u) src[u] = sbase + size_t(chunk[u]) * (kChunkBytes / 16) + lane;
  if (f.vs.tags) {                                    // HBM chunk store: read-only here (skv_vstore_fill updates it)
    int tag[kFetchPerWarp];
#pragma unroll
    for (int u = 0; u < kFetchPerWarp; ++u) {
      const size_t w = size_t(bh) * f.vs.ways + (static_cast<int>(chunk[u]) & (f.vs.ways - 1));
      tag[u] = on[u] ? f.vs.tags[w] : -1;
    }
#pragma unroll
    for (int u = 0; u < kFetchPerWarp; ++u) {
      if (!on[u]) continue;
      const size_t w = size_t(bh) * f.vs.ways + (static_cast<int>(chunk[u]) & (f.vs.ways - 1));
      if (tag[u] == static_cast<int>(chunk[u]))       // warp uniform
        src[u] = reinterpret_cast<const int4 *>(f.vs.data) + w * (kChunkBytes / 16) + lane;
      else if (f.vs.pcie_count && lane == 0)
        atomicAdd(f.vs.pcie_count + bh, 1);
    }
  }
  int4 v[kFetchPerWarp][4];
#pragma unroll
  for (int u = 0; u < kFetchPerWarp; ++u) {
    if (on[u]) {
#pragma unroll
      for (int k = 0; k < 4; ++k) v[u][k] = ld_stream(src[u] + k * 32);
    }
  }
#pragma unroll
  for (int u = 0; u < kFetchPerWarp; ++u) {
    if (on[u]) {
      int4 *d = dbase + size_t(slot[u]) * (kChunkBytes / 16) + lane;
#pragma unroll
      for (int k = 0; k < 4; ++k) d[k * 32] = v[u][k];
    }
  }
}

__global__ void __launch_bounds__(256, 2)
    fetch_rebuild_kernel(const __grid_constant__ CUtensorMap map_sv, RebuildParams p, FetchParams f) {
  extern __shared__ __align__(1024) unsigned char smem[];
  pdl_launch_dependents();
  pdl_wait();                                         // offsets / cnts / resident come from the selection kernel
  if (threadIdx.x < 128) rebuild_tile(&map_sv, p, blockIdx.x, blockIdx.y, smem);
  else fetch_entries(f, p.cnts, p.S, blockIdx.x, blockIdx.y, int(threadIdx.x >> 5) - 4);
}

// ------------------------------------------------------------------------------------------------
// SIMT cross-check of the same computation (tests only; impl = 1).  One CTA per missed chunk.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
    rebuild_simt_kernel(const __nv_bfloat16 *__restrict__ U, const __nv_bfloat16 *__restrict__ SV, RebuildParams p) {
  __shared__ float s_u[kChunkTokens][256];
  __shared__ __nv_bfloat16 s_k[kChunkTokens][kHeadDim];
  const int bh = blockIdx.y, b = bh / p.H, d = threadIdx.x;
  const int cnt = p.cnts ? p.cnts[bh] : 0;
  int slot = blockIdx.x;
  if (p.miss_slots) {
    if (slot >= p.S - cnt) return;
    slot = p.miss_slots[size_t(bh) * p.S + slot];
  } else if (p.tile_skip ? (slot / kTileChunks + 1) * kTileChunks <= cnt : slot < cnt) {
    return;
  }
  const int id = p.resident ? static_cast<int>(p.resident[size_t(bh) * p.S + slot]) : p.ids32[size_t(bh) * p.S + slot];
  const __nv_bfloat16 *u = U + (size_t(b) * p.L + size_t(id) * kChunkTokens) * p.R;
  for (int i = threadIdx.x; i < kChunkTokens * p.R; i += 128) s_u[i / p.R][i % p.R] = __bfloat162float(u[i]);
  __syncthreads();
  const __nv_bfloat16 *sv = SV + (size_t(bh) * kHeadDim + d) * p.R;
  float acc[kChunkTokens] = {};
  for (int k = 0; k < p.R; ++k) {
    const float w = __bfloat162float(sv[k]);
#pragma unroll
    for (int r = 0; r < kChunkTokens; ++r) acc[r] = fmaf(s_u[r][k], w, acc[r]);
  }
#pragma unroll
  for (int r = 0; r < kChunkTokens; ++r) s_k[r][d] = __float2bfloat16_rn(acc[r]);
  __syncthreads();
#pragma unroll 1
  for (int r = 0; r < kChunkTokens; ++r) {
    const long long token = (long long)id * kChunkTokens + r;
    __nv_bfloat16 o = s_k[r][d];
    if (p.rope == 1) {
      const __nv_bfloat16 *cs = p.cos_sin + token * kHeadDim;
      __nv_bfloat16 o1, o2;
      if (d < 64) { rope_pair_bf16(s_k[r][d], s_k[r][d + 64], cs[d], cs[64 + d], o1, o2); o = o1; }
      else        { rope_pair_bf16(s_k[r][d - 64], s_k[r][d], cs[d - 64], cs[d], o1, o2); o = o2; }
    } else if (p.rope == 2 && d < 64) {
      const __nv_bfloat16 *cs = p.cos_sin + token * (kHeadDim / 2);
      const int pr = d >> 1;
      __nv_bfloat16 o1, o2;
      rope_pair_bf16(s_k[r][2 * pr], s_k[r][2 * pr + 1], cs[pr], cs[32 + pr], o1, o2);
      o = (d & 1) ? o2 : o1;
    }
    p.out[size_t(bh) * p.out_stride_bh + (p.out_row0 + size_t(slot) * kChunkTokens + r) * kHeadDim + d] = o;
  }
}

// ------------------------------------------------------------------------------------------------
// Host side: tensor maps (cached per pointer/shape) and launch
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void *sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  });
  return fn;
}

// 2-D bf16 row-major [rows, cols] tensor, box {64 cols, box_rows}, 128-byte swizzle, zero OOB fill.
int make_tensor_map_2d(CUtensorMap *out, const void *base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                       uint32_t box_cols) {
  typedef std::tuple<const void *, uint64_t, uint64_t, uint32_t, uint32_t> Key;
  static std::map<Key, CUtensorMap> cache;
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  Key key(base, rows, cols, box_rows, box_cols);
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return SKV_OK; }
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return SKV_ERR_DRIVER; }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", int(r)); return SKV_ERR_DRIVER; }
  if (cache.size() > 4096) cache.clear();
  cache[key] = m;
  *out = m;
  return SKV_OK;
}

// 4-D map over a K / V cache view for the attention kernel (attention_mma.cu): dims {128, rows, heads, batch},
// strides in elements, box {64 dims, 64 rows, 1, 1}, 128-byte swizzle, zero OOB fill.
int make_tensor_map_kv(CUtensorMap *out, const void *base, uint64_t rows, uint64_t heads, uint64_t batch, int64_t stride_row,
                       int64_t stride_h, int64_t stride_b) {
  typedef std::tuple<const void *, uint64_t, uint64_t, uint64_t, int64_t, int64_t, int64_t> Key;
  static std::map<Key, CUtensorMap> cache;
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  Key key(base, rows, heads, batch, stride_row, stride_h, stride_b);
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return SKV_OK; }
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return SKV_ERR_DRIVER; }
  cuuint64_t dims[4] = {cuuint64_t(kHeadDim), rows, heads, batch};
  cuuint64_t strides[3] = {cuuint64_t(stride_row) * 2, cuuint64_t(stride_h) * 2, cuuint64_t(stride_b) * 2};
  cuuint32_t box[4] = {64, 64, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (kv view) failed (%d)", int(r)); return SKV_ERR_DRIVER; }
  if (cache.size() > 4096) cache.clear();
  cache[key] = m;
  *out = m;
  return SKV_OK;
}

static int launch_rebuild(const skv_bf16 *U, const skv_bf16 *SV, RebuildParams p, int B, int impl, cudaStream_t st) {
  const int BH = B * p.H;
  if (p.R % 8 || p.R < 8 || p.R > kMaxSlabs * kSlabCols) { set_error("rebuild: rank %d unsupported", p.R); return SKV_ERR_UNSUPPORTED; }
  p.n_slabs = (p.R + kSlabCols - 1) / kSlabCols;
  p.k_steps = (p.R + 15) / 16;
  if (impl == 1) {
    dim3 grid(p.S, BH);
    rebuild_simt_kernel<<<grid, 128, 0, st>>>(reinterpret_cast<const __nv_bfloat16 *>(U),
                                              reinterpret_cast<const __nv_bfloat16 *>(SV), p);
    return launch_status("rebuild_simt_kernel");
  }
  CUtensorMap map_sv;
  int rc = make_tensor_map_2d(&map_sv, SV, uint64_t(BH) * kHeadDim, p.R, kHeadDim, kSlabCols);
  if (rc) return rc;
  p.U = reinterpret_cast<const __nv_bfloat16 *>(U);
  const int smem = p.n_slabs * (kSlabBytesA + kSlabBytesB) + (kMaxSlabs + 1) * 8 + 8 + 2 * 16 * 4 + 64;
  if (int rc2 = ensure_dynamic_smem(reinterpret_cast<const void *>(rebuild_tcgen05_kernel), smem, "rebuild_tcgen05_kernel")) return rc2;
  dim3 grid((p.S + kTileChunks - 1) / kTileChunks, BH);
  rebuild_tcgen05_kernel<<<grid, 128, smem, st>>>(map_sv, p);
  return launch_status("rebuild_tcgen05_kernel");
}

static int launch_fetch_rebuild(const skv_bf16 *U, const skv_bf16 *SV, RebuildParams p, FetchParams f, int B, cudaStream_t st) {
  const int BH = B * p.H;
  if (p.R % 8 || p.R < 8 || p.R > kMaxSlabs * kSlabCols) { set_error("fetch_and_rebuild: rank %d unsupported", p.R); return SKV_ERR_UNSUPPORTED; }
  p.n_slabs = (p.R + kSlabCols - 1) / kSlabCols;
  p.k_steps = (p.R + 15) / 16;
  CUtensorMap map_sv;
  int rc = make_tensor_map_2d(&map_sv, SV, uint64_t(BH) * kHeadDim, p.R, kHeadDim, kSlabCols);
  if (rc) return rc;
  p.U = reinterpret_cast<const __nv_bfloat16 *>(U);
  const int smem = p.n_slabs * (kSlabBytesA + kSlabBytesB) + (kMaxSlabs + 1) * 8 + 8 + 2 * 16 * 4 + 64;
  if (int rc2 = ensure_dynamic_smem(reinterpret_cast<const void *>(fetch_rebuild_kernel), smem, "fetch_rebuild_kernel")) return rc2;
  static_assert(kTileChunks == 4 * kFetchPerWarp, "a CTA fetches the V chunks of the tile it rebuilds");
  const int n_tiles = (p.S + kTileChunks - 1) / kTileChunks;
  cudaError_t e = launch_kernel(fetch_rebuild_kernel, dim3(n_tiles, BH), dim3(256), smem, st, /*pdl=*/true, map_sv, p, f);
  if (e != cudaSuccess) { set_error("fetch_rebuild_kernel: %s", cudaGetErrorString(e)); return int(e); }
  return launch_status("fetch_rebuild_kernel");
}

}  // namespace skv

using namespace skv;

extern "C" {

int skv_fetch_and_rebuild(const skv_geom *g, const skv_bf16 *v_host, int64_t host_stride, skv_bf16 *v_buf,
                          const skv_bf16 *U, const skv_bf16 *SV, const skv_bf16 *cos_sin, skv_bf16 *k_buf,
                          const int64_t *resident, const int32_t *cnts, const int32_t *offsets,
                          const skv_vstore *vstore, skv_stream_t stream) {
  if (!g || !v_host || !v_buf || !U || !SV || !cos_sin || !k_buf || !resident || !cnts || !offsets) {
    set_error("fetch_and_rebuild: null pointer"); return SKV_ERR_ARGUMENT;
  }
  if (vstore && (!vstore->tags || !vstore->data || vstore->ways < 1 || (vstore->ways & (vstore->ways - 1)))) {
    set_error("fetch_and_rebuild: chunk store needs tags, data and a power-of-two way count"); return SKV_ERR_ARGUMENT;
  }
  if (g->head_dim != kHeadDim || g->chunk != kChunkTokens || !valid_map_size(g->select)) {
    set_error("fetch_and_rebuild: unsupported geometry"); return SKV_ERR_UNSUPPORTED;
  }
  RebuildParams p{};
  p.resident = resident; p.cnts = cnts; p.miss_slots = g->table_mode ? offsets : nullptr;
  p.cos_sin = reinterpret_cast<const __nv_bfloat16 *>(cos_sin);
  p.out = reinterpret_cast<__nv_bfloat16 *>(k_buf);
  p.out_stride_bh = int64_t(g->buf_len) * g->head_dim;
  p.out_row0 = g->sparse_start;
  p.H = g->kv_heads; p.S = g->select; p.L = g->seq_len; p.R = g->rank;
  p.rope = g->rope_glm ? 2 : 1;
  p.tile_skip = 0;
  FetchParams f{};
  f.src = v_host; f.src_stride_bh = host_stride; f.dst = v_buf;
  f.dst_stride_bh = int64_t(g->buf_len) * g->head_dim;
  f.dst_base_off = int64_t(g->sparse_start) * g->head_dim;
  f.offsets = offsets; f.resident = g->table_mode ? resident : nullptr;
  if (vstore) f.vs = *vstore;
  return launch_fetch_rebuild(U, SV, p, f, g->batch, as_stream(stream));
}

int skv_rebuild_keys(const skv_geom *g, const skv_bf16 *U, const skv_bf16 *SV, const skv_bf16 *cos_sin,
                     const int64_t *resident, const int32_t *cnts, const int32_t *offsets, skv_bf16 *k_buf, int impl,
                     skv_stream_t stream) {
  if (!g || !U || !SV || !cos_sin || !resident || !cnts || !k_buf || (g->table_mode && !offsets)) {
    set_error("rebuild_keys: null pointer"); return SKV_ERR_ARGUMENT;
  }
  if (g->head_dim != kHeadDim || g->chunk != kChunkTokens || !valid_map_size(g->select)) {
    set_error("rebuild_keys: unsupported geometry"); return SKV_ERR_UNSUPPORTED;
  }
  RebuildParams p{};
  p.resident = resident; p.cnts = cnts; p.miss_slots = g->table_mode ? offsets : nullptr;
  p.cos_sin = reinterpret_cast<const __nv_bfloat16 *>(cos_sin);
  p.out = reinterpret_cast<__nv_bfloat16 *>(k_buf);
  p.out_stride_bh = int64_t(g->buf_len) * g->head_dim;
  p.out_row0 = g->sparse_start;
  p.H = g->kv_heads; p.S = g->select; p.L = g->seq_len; p.R = g->rank;
  p.rope = g->rope_glm ? 2 : 1;
  p.tile_skip = 0;
  return launch_rebuild(U, SV, p, g->batch, impl, as_stream(stream));
}

int skv_batch_gather_gemm(const skv_bf16 *a, const skv_bf16 *b, const skv_bf16 *cos, const skv_bf16 *sin,
                          const int32_t *position_ids, skv_bf16 *output, int batch_size, int heads, int seq_len,
                          int embed_dim, int rank, int sparse_budget, int max_seq_len, int chunk_size,
                          const int32_t *offset_array, skv_stream_t stream) {
  (void)cos; (void)sin; (void)max_seq_len;
  if (!a || !b || !position_ids || !output) { set_error("batch_gather_gemm: null pointer"); return SKV_ERR_ARGUMENT; }
  if (embed_dim != kHeadDim || chunk_size != kChunkTokens || sparse_budget % (kChunkTokens * kTileChunks)) {
    set_error("batch_gather_gemm: need embed_dim=128, chunk_size=8, budget %% 128 == 0"); return SKV_ERR_UNSUPPORTED;
  }
  RebuildParams p{};
  p.ids32 = position_ids; p.cnts = offset_array;
  p.out = reinterpret_cast<__nv_bfloat16 *>(output);
  p.out_stride_bh = int64_t(sparse_budget) * embed_dim;
  p.out_row0 = 0;
  p.H = heads; p.S = sparse_budget / chunk_size; p.L = seq_len; p.R = rank;
  p.rope = 0;
  p.tile_skip = 1;
  const char *impl_env = getenv("SKV_REBUILD_IMPL");
  return launch_rebuild(a, b, p, batch_size, impl_env && impl_env[0] == '1' ? 1 : 0, as_stream(stream));
}

}  // extern "C"
