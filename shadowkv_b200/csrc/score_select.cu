// Landmark scoring AND chunk selection in ONE launch (stage a3-a5 of the decode step): one thread-block cluster per
// (b, kv head) row streams the row's landmark matrix once through TMA rings (HBM bound: N*D*2 bytes), takes the
// G x 128-landmark products on tcgen05 (fp32 accumulators in TMEM), keeps the bf16 scores IN SHARED MEMORY, exchanges
// the softmax statistics and the radix-select histograms through distributed shared memory, and goes straight into the
// top-S selection, the chunk-id gather and the chunk-table update.
//
// Replaces, on the product path, score_tcgen05_kernel + select_kernel (score_mma.cu, score.cu): no score round trip
// through L2 (1 MB written and re-read per sequence-layer), no per-tile partial statistics in global memory, no launch
// boundary between the stream and the selection.  Same rounding points and the same selection rule, so the result is
// bit-identical to the two-launch path: bf16(alpha * q.k) from fp32 accumulation, softmax max / sum over the ROUNDED
// scores, p = bf16(exp(s - max) / sum), key = max over the G heads, top-S with ties to the lowest landmark index
// (reference: kernels/batch_gemm_softmax.cu:229-293, models/kv_cache.py:1023-1042, kernels/map.cuh:754-796).
//
// Cluster size: 16 CTAs (non-portable size, one cluster per GPC) while B*H <= 9 rows -- 128 CTAs keep HBM busy at batch
// 1 -- else 8.  Every CTA owns a contiguous range of 128-landmark tiles.  Warp roles while streaming as in
// score_mma.cu: warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2..9 = two epilogue warpgroups; the
// 256 epilogue threads then run the selection.  S <= 256 (sparse budget <= 2048); larger tables keep the two launches.
#include <cooperative_groups.h>

#include <cstdlib>
#include <mutex>

#include "chunk_table.cuh"
#include "sm100_ptx.cuh"

namespace cg = cooperative_groups;

namespace skv {

constexpr int kSsTileRows = 128;
constexpr int kSsSlabBytes = kSsTileRows * 128;
constexpr int kSsTileBytes = 2 * kSsSlabBytes;          // 32 KiB per tile
constexpr int kSsQSlabBytes = 16 * 128;
constexpr int kSsThreads = 64 + 256;
constexpr int kSsWorkers = 256, kSsWarps = 8;
constexpr int kSsAccCols = 16;
constexpr int kSsMaxCluster = 16;
constexpr int kSsMaxStages = 6;
constexpr int kSsMaxTilesPerCta = 128;                  // per-tile partial statistics kept in shared memory

struct ScoreSelectParams {
  const skv_bf16 *q;              // [BH, G, 128]
  const int64_t *landmark_idx;    // [BH, N]
  int64_t *resident;              // [BH, S] in/out
  int32_t *offsets, *cnts;
  int64_t *selected_out;          // [BH, S] or null
  skv_bf16 *score_out;            // [BH, N] or null
  int G, N, T, tiles_per_cta, stages, dynamic_lm, ld_sc;
  float alpha;
  skv_append ap;
  int64_t buf_rows;
  // byte offsets inside dynamic shared memory
  int off_q, off_bars, off_stat, off_parts, off_new, off_sc, off_keys, off_pm, off_small;
};

__device__ __forceinline__ void ss_hist_add(int *hist, unsigned bin) {       // warp-private histogram, no atomics
  const unsigned peers = __match_any_sync(0xffffffffu, bin);
  if (bin != 0xFFFFu && (threadIdx.x & 31) == (__ffs(peers) - 1)) hist[bin] += __popc(peers);
  __syncwarp();
}

// Worker warp 0: highest bin b with (count above b) < want <= (count of bins >= b).  res[0] = b, res[1] = count above b.
__device__ __forceinline__ void ss_find_bin(const int *hist, int want, int *res) {
  const int lane = threadIdx.x & 31;
  int local[8], sum = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { local[i] = hist[255 - (lane * 8 + i)]; sum += local[i]; }
  const int incl = warp_incl_scan(sum);
  const unsigned hitmask = __ballot_sync(0xffffffffu, incl >= want);
  const int first = __ffs(hitmask) - 1;
  if (lane == first) {
    int above = incl - sum;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (above + local[i] >= want) { res[0] = 255 - (lane * 8 + i); res[1] = above; break; }
      above += local[i];
    }
  }
}

template <int S, bool kInPlace>
__global__ void __launch_bounds__(kSsThreads, 1)
    score_select_kernel(const __grid_constant__ CUtensorMap map_lm, ScoreSelectParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char *sA = smem;                                                      // p.stages x 32 KiB ring
  unsigned char *sQ = smem + p.off_q;                                            // 2 slabs x 2 KiB
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + p.off_bars);
  uint64_t *empty_bar = full_bar + kSsMaxStages;
  uint64_t *acc_full = empty_bar + kSsMaxStages;                                 // [2]
  uint64_t *acc_empty = acc_full + 2;                                            // [2]
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);
  // written by OTHER CTAs of the cluster (never aliased with the ring)
  float *s_stat = reinterpret_cast<float *>(smem + p.off_stat);                  // [16 ranks][16 heads][2]
  int *s_parts1 = reinterpret_cast<int *>(smem + p.off_parts);                   // [16 ranks][256]
  int *s_parts2 = s_parts1 + kSsMaxCluster * 256;                                // [16 ranks][256]
  int *s_new = reinterpret_cast<int *>(smem + p.off_new);                        // [S] (rank 0 receives the selection)
  // local
  uint16_t *s_sc = reinterpret_cast<uint16_t *>(smem + p.off_sc);                // [G][ld_sc] bf16 scores of this CTA
  uint16_t *keys = reinterpret_cast<uint16_t *>(smem + p.off_keys);              // [ld_sc]
  float *s_pm = reinterpret_cast<float *>(smem + p.off_pm);                      // [16 heads][tiles_per_cta] tile max
  float *s_ps = s_pm + 16 * p.tiles_per_cta;                                     // [16 heads][tiles_per_cta] tile sum
  float *s_red = reinterpret_cast<float *>(smem + p.off_small);                  // [2 acc][4 warps][16 heads][2]
  float *s_m = s_red + 2 * 4 * 32;                                               // [16]
  float *s_inv = s_m + 16;                                                       // [16]
  int *s_tot = reinterpret_cast<int *>(s_inv + 16);                              // [256]
  int *s_res = s_tot + 256;                                                      // [4]
  int *s_wa = s_res + 4, *s_wb = s_wa + kSsWarps + 1;                            // [9] each
  int *s_sorted = s_wb + kSsWarps + 1, *s_slot = s_sorted + S, *s_warp = s_slot + S;   // chunk table scratch
  // warp-private histograms: over the ring, once this CTA's stream is finished
  int *whist1 = reinterpret_cast<int *>(sA);                                     // [8 warps][256]
  int *whist2 = whist1 + kSsWarps * 256;

  cg::cluster_group cluster = cg::this_cluster();
  const int csz = int(cluster.num_blocks());
  const int rank = int(cluster.block_rank());
  const int bh = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wt = tid - 64;                                                        // worker thread id 0..255 (warps 2..9)
  const int wwarp = warp - 2;
  const int G = p.G, N = p.N;
  const int tile0 = rank * p.tiles_per_cta;                                       // this CTA's contiguous tile range
  const int n_my = max(0, min(p.T - tile0, p.tiles_per_cta));
  const int n_begin = min(N, tile0 * kSsTileRows);
  const int count = min(N, (tile0 + n_my) * kSsTileRows) - n_begin;               // landmarks of this CTA

  const int tr = bh == 0 ? rank : -1;               // SKV_TRACE builds: timeline row
  if (tid == 0) SKV_STAMP_NS(tr, 0);
  pdl_launch_dependents();
  // every CTA of the cluster must be running before its shared memory is written remotely: arrive now, wait in front
  // of the first exchange (by then the barrier has long completed)
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) { mbarrier_init(&full_bar[s], 1); mbarrier_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbarrier_init(&acc_full[a], 1); mbarrier_init(&acc_empty[a], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // the landmarks are static state: under a programmatic launch the ring fills while the previous kernel finishes
    const uint64_t pol = l2_policy_evict_first();
    for (int i = 0; i < p.stages && i < n_my && !p.dynamic_lm; ++i) {
      const int row0 = bh * N + (tile0 + i) * kSsTileRows;
      mbarrier_expect_tx(&full_bar[i], kSsTileBytes);
      tma_load_2d_hint(sA + i * kSsTileBytes, &map_lm, 0, row0, &full_bar[i], pol);
      tma_load_2d_hint(sA + i * kSsTileBytes + kSsSlabBytes, &map_lm, 64, row0, &full_bar[i], pol);
    }
  }
  if (warp == 1) tcgen05_alloc(tmem_slot, 32);
  // static too: the landmark -> chunk id table of this thread's first 8 landmarks
  const int64_t *lidx = p.landmark_idx + size_t(bh) * N + n_begin;
  int ids0[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) ids0[e] = (wt >= 0 && wt * 8 + e < count) ? static_cast<int>(lidx[wt * 8 + e]) : 0;
  pdl_wait();                                    // q, the K / V buffers and the chunk table are the predecessors' outputs

  int old_id = 0;
  if (rank == 0 && wt >= 0 && wt < S) old_id = static_cast<int>(p.resident[size_t(bh) * S + wt]);          // map.cuh:771
  // fused append (a2): the first CTA of every row writes the new token's K and V rows
  if (p.ap.k_new && rank == 0 && warp == 2) {
    const int64_t pos = p.ap.pos + (p.ap.pos_dev ? *p.ap.pos_dev : 0);
    if (pos < p.buf_rows) {
      const int c = lane & 15;
      const int4 *src = reinterpret_cast<const int4 *>(lane < 16 ? p.ap.k_new : p.ap.v_new) + size_t(bh) * 16 + c;
      int4 *dst = reinterpret_cast<int4 *>(lane < 16 ? p.ap.k_buf : p.ap.v_buf) + (size_t(bh) * p.buf_rows + pos) * 16 + c;
      *dst = *src;
    }
  }
  // B operand: the G query rows, zero padded to 16, 128-byte-swizzled K-major
  for (int i = tid; i < 2 * 16 * 8; i += kSsThreads) {
    const int slab = i >> 7, r = (i >> 3) & 15, c = i & 7;
    int4 v = make_int4(0, 0, 0, 0);
    if (r < G) v = *reinterpret_cast<const int4 *>(p.q + (size_t(bh) * G + r) * kHeadDim + slab * 64 + c * 8);
    *reinterpret_cast<int4 *>(sQ + slab * kSsQSlabBytes + (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)) = v;
  }
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0) SKV_STAMP_NS(tr, 1);

  // =========================== the stream: scores of this CTA's tiles into shared memory ===========================
  if (warp == 0) {
    if (lane == 0) {
      const uint64_t pol = l2_policy_evict_first();
      for (int i = p.dynamic_lm ? 0 : p.stages; i < n_my; ++i) {
        const int stage = i % p.stages, ph = (i / p.stages) & 1;
        mbarrier_wait(&empty_bar[stage], ph ^ 1);
        mbarrier_expect_tx(&full_bar[stage], kSsTileBytes);
        const int row0 = bh * N + (tile0 + i) * kSsTileRows;     // rows past N belong to the next (b,h) or are zero
        tma_load_2d_hint(sA + stage * kSsTileBytes, &map_lm, 0, row0, &full_bar[stage], pol);
        tma_load_2d_hint(sA + stage * kSsTileBytes + kSsSlabBytes, &map_lm, 64, row0, &full_bar[stage], pol);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16_m128(kSsAccCols);
      const uint64_t db0 = umma_desc_k_sw128(smem_addr(sQ)), db1 = umma_desc_k_sw128(smem_addr(sQ + kSsQSlabBytes));
      for (int i = 0; i < n_my; ++i) {
        const int stage = i % p.stages, ph = (i / p.stages) & 1;
        const int acc = i & 1, aph = (i >> 1) & 1;
        mbarrier_wait(&acc_empty[acc], aph ^ 1);
        mbarrier_wait(&full_bar[stage], ph);
        tcgen05_fence_after();
        const uint64_t da0 = umma_desc_k_sw128(smem_addr(sA + stage * kSsTileBytes));
        const uint64_t da1 = umma_desc_k_sw128(smem_addr(sA + stage * kSsTileBytes + kSsSlabBytes));
        const uint32_t d = tmem_base + acc * kSsAccCols;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) tcgen05_mma_bf16(d, da0 + uint64_t(kk * 2), db0 + uint64_t(kk * 2), idesc, kk > 0);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) tcgen05_mma_bf16(d, da1 + uint64_t(kk * 2), db1 + uint64_t(kk * 2), idesc, 1);
        tcgen05_commit(&empty_bar[stage]);
        tcgen05_commit(&acc_full[acc]);
      }
    }
  } else {
    const int lg = warp & 3;                                 // TMEM lane group this warp may read
    const int grp = wwarp >> 2;                              // epilogue warpgroup 0 / 1 <-> accumulator buffer
    const int ew = wwarp & 3;
    for (int i = grp; i < n_my; i += 2) {
      const int acc = i & 1, aph = (i >> 1) & 1;
      mbarrier_wait(&acc_full[acc], aph);
      tcgen05_fence_after();
      uint32_t r[16];
      tcgen05_ld16(tmem_base + (uint32_t(lg * 32) << 16) + acc * kSsAccCols, r);
      tcgen05_wait_ld();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbarrier_arrive(&acc_empty[acc]);
      const int nl = i * kSsTileRows + lg * 32 + lane;       // landmark index inside this CTA
      const bool valid = nl < count;
      float *red = s_red + (acc * 4 + ew) * 32;
#pragma unroll
      for (int g0 = 0; g0 < 16; g0 += 4) {
        if (g0 < G) {                                        // warp uniform
          float x[4], m[4], e[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint16_t bits = f32_to_bf16_bits(__uint_as_float(r[g0 + k]) * p.alpha);
            const bool on = valid && g0 + k < G;
            if (g0 + k < G) s_sc[(g0 + k) * p.ld_sc + nl] = on ? bits : uint16_t(0xFF80u);   // -inf beyond the slice
            x[k] = on ? bf16_bits_to_f32(bits) : -INFINITY;
            m[k] = x[k];
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int k = 0; k < 4; ++k) m[k] = fmaxf(m[k], __shfl_xor_sync(0xffffffffu, m[k], o));
#pragma unroll
          for (int k = 0; k < 4; ++k) e[k] = (x[k] == -INFINITY) ? 0.f : expf(x[k] - m[k]);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int k = 0; k < 4; ++k) e[k] += __shfl_xor_sync(0xffffffffu, e[k], o);
          if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) { red[(g0 + k) * 2] = m[k]; red[(g0 + k) * 2 + 1] = e[k]; }
          }
        }
      }
      asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
      if (ew == 0 && lane < G) {
        const float *rb = s_red + acc * 4 * 32 + lane * 2;
        const float m = fmaxf(fmaxf(rb[0], rb[32]), fmaxf(rb[64], rb[96]));
        float sum = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          const float mw = rb[w * 32];
          sum += (mw == -INFINITY) ? 0.f : rb[w * 32 + 1] * expf(mw - m);
        }
        s_pm[lane * p.tiles_per_cta + i] = m;
        s_ps[lane * p.tiles_per_cta + i] = sum;
      }
      asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
    }
  }
  tcgen05_fence_before();
  __syncthreads();                               // the ring is idle from here on (every tile landed and was consumed)
  if (warp == 1) tcgen05_dealloc(tmem_base, 32);
  if (tid == 0) SKV_STAMP_NS(tr, 2);

  // =========================== softmax statistics of the whole row (exchange 1) ====================================
  // this CTA's (max, sum) per head, pushed to every CTA of the cluster
  asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
  if (wt >= 0) {
    for (int i = wt; i < 2 * kSsWarps * 256; i += kSsWorkers) whist1[i] = 0;      // both histogram levels
    for (int g = wwarp; g < G; g += kSsWarps) {
      float m = -INFINITY;
      for (int t = lane; t < n_my; t += 32) m = fmaxf(m, s_pm[g * p.tiles_per_cta + t]);
      m = warp_max(m);
      float s = 0.f;
      for (int t = lane; t < n_my; t += 32) {
        const float pm = s_pm[g * p.tiles_per_cta + t];
        s += (pm == -INFINITY) ? 0.f : s_ps[g * p.tiles_per_cta + t] * expf(pm - m);
      }
      s = warp_sum(s);
      if (lane < csz) {
        float *dst = cluster.map_shared_rank(s_stat, lane) + (rank * 16 + g) * 2;
        dst[0] = m;
        dst[1] = s;
      }
    }
  }
  cluster.sync();
  if (wt >= 0 && wt < G) {
    float m = -INFINITY;
    for (int r = 0; r < csz; ++r) m = fmaxf(m, s_stat[(r * 16 + wt) * 2]);
    float s = 0.f;
    for (int r = 0; r < csz; ++r) {
      const float mr = s_stat[(r * 16 + wt) * 2];
      s += (mr == -INFINITY) ? 0.f : s_stat[(r * 16 + wt) * 2 + 1] * expf(mr - m);
    }
    s_m[wt] = m;
    s_inv[wt] = 1.0f / s;
  }
  __syncthreads();
  if (tid == 0) SKV_STAMP_NS(tr, 3);

  // =========================== keys + level-1 histogram (exchange 2) ================================================
  int *my_h1 = whist1 + (wwarp & 7) * 256, *my_h2 = whist2 + (wwarp & 7) * 256;
  if (wt >= 0) {
    for (int base = 0; base < count; base += kSsWorkers * 8) {
      const int off = base + wt * 8;
      const bool on = off < count;
      uint16_t best[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      if (on) {
        for (int g = 0; g < G; ++g) {
          const int4 w = *reinterpret_cast<const int4 *>(s_sc + g * p.ld_sc + off);
          const uint32_t u[4] = {(uint32_t)w.x, (uint32_t)w.y, (uint32_t)w.z, (uint32_t)w.w};
          const float m = s_m[g], inv = s_inv[g];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const uint16_t b0 = f32_to_bf16_bits(expf(bf16_lo(u[e]) - m) * inv);
            const uint16_t b1 = f32_to_bf16_bits(expf(bf16_hi(u[e]) - m) * inv);
            best[2 * e] = b0 > best[2 * e] ? b0 : best[2 * e];
            best[2 * e + 1] = b1 > best[2 * e + 1] ? b1 : best[2 * e + 1];
          }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (off + e >= count) best[e] = 0;                    // ragged tail: never selected (p > 0 for real keys)
        const int4 packed = make_int4(int(best[0] | (uint32_t(best[1]) << 16)), int(best[2] | (uint32_t(best[3]) << 16)),
                                      int(best[4] | (uint32_t(best[5]) << 16)), int(best[6] | (uint32_t(best[7]) << 16)));
        *reinterpret_cast<int4 *>(keys + off) = packed;
        if (p.score_out) {
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if (off + e < count) p.score_out[size_t(bh) * N + n_begin + off + e] = best[e];
        }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) ss_hist_add(my_h1, (on && off + e < count) ? unsigned(best[e] >> 8) : 0xFFFFu);
    }
  }
  __syncthreads();
  if (wt >= 0) {
    int t = 0;
#pragma unroll
    for (int w = 0; w < kSsWarps; ++w) t += whist1[w * 256 + wt];
    for (int r = 0; r < csz; ++r) cluster.map_shared_rank(s_parts1, r)[rank * 256 + wt] = t;
  }
  if (tid == 0) SKV_STAMP_NS(tr, 4);
  cluster.sync();
  if (tid == 0) SKV_STAMP_NS(tr, 5);
  if (wt >= 0) {
    int t = 0;
    for (int r = 0; r < csz; ++r) t += s_parts1[r * 256 + wt];
    s_tot[wt] = t;
  }
  __syncthreads();
  if (wwarp == 0) ss_find_bin(s_tot, S, s_res);
  __syncthreads();
  const int b1 = s_res[0], above1 = s_res[1];
  if (tid == 0) SKV_STAMP_NS(tr, 6);

  // =========================== level 2 inside the threshold bin (exchange 3) =======================================
  if (wt >= 0) {
    for (int base = 0; base < count; base += kSsWorkers * 8) {
      const int off = base + wt * 8;
      const bool on = off < count;
      uint32_t kw[4] = {0, 0, 0, 0};
      if (on) {
        const int4 w = *reinterpret_cast<const int4 *>(keys + off);
        kw[0] = w.x; kw[1] = w.y; kw[2] = w.z; kw[3] = w.w;
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const unsigned k = (kw[e >> 1] >> ((e & 1) * 16)) & 0xFFFFu;
        ss_hist_add(my_h2, (on && off + e < count && int(k >> 8) == b1) ? (k & 0xFFu) : 0xFFFFu);
      }
    }
  }
  __syncthreads();
  if (wt >= 0) {
    int t = 0;
#pragma unroll
    for (int w = 0; w < kSsWarps; ++w) t += whist2[w * 256 + wt];
    for (int r = 0; r < csz; ++r) cluster.map_shared_rank(s_parts2, r)[rank * 256 + wt] = t;
  }
  if (tid == 0) SKV_STAMP_NS(tr, 7);
  cluster.sync();
  if (tid == 0) SKV_STAMP_NS(tr, 8);
  if (wt >= 0) {
    int t = 0;
    for (int r = 0; r < csz; ++r) t += s_parts2[r * 256 + wt];
    s_tot[wt] = t;
  }
  __syncthreads();
  if (wwarp == 0) ss_find_bin(s_tot, S - above1, s_res + 2);
  __syncthreads();
  const unsigned thr = (unsigned(b1) << 8) | unsigned(s_res[2]);
  const int need_eq = S - (above1 + s_res[3]);      // how many keys equal to thr are taken (>= 1)
  if (tid == 0) SKV_STAMP_NS(tr, 9);

  // (greater, equal) counts of the lower-ranked CTAs: every CTA already holds all partial histograms
  int gt_before = 0, eq_before = 0;
  {
    const int b2 = s_res[2];
    int a = 0, e = 0;
    if (wt >= 0) {
      for (int r = 0; r < rank; ++r) {
        if (wt > b1) a += s_parts1[r * 256 + wt];
        if (wt > b2) a += s_parts2[r * 256 + wt];
        if (wt == b2) e += s_parts2[r * 256 + wt];
      }
    }
    a = warp_sum_int(a); e = warp_sum_int(e);
    if (wt >= 0 && lane == 0) { s_wa[wwarp] = a; s_wb[wwarp] = e; }
    __syncthreads();
    for (int w = 0; w < kSsWarps; ++w) { gt_before += s_wa[w]; eq_before += s_wb[w]; }
    __syncthreads();
  }

  // =========================== ordered emission into rank 0's shared memory (exchange 4) ===========================
  {
    int *dst = cluster.map_shared_rank(s_new, 0);
    for (int base = 0; base < count; base += kSsWorkers * 8) {
      const int off = base + max(wt, 0) * 8;
      const bool on = wt >= 0 && off < count;
      uint32_t kw[4] = {0, 0, 0, 0};
      if (on) {
        const int4 w = *reinterpret_cast<const int4 *>(keys + off);
        kw[0] = w.x; kw[1] = w.y; kw[2] = w.z; kw[3] = w.w;
      }
      int gt = 0, eq = 0;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const unsigned k = (kw[e >> 1] >> ((e & 1) * 16)) & 0xFFFFu;
        const bool real = on && off + e < count;
        gt += real && k > thr; eq += real && k == thr;
      }
      // exclusive scan of (gt, eq) over the 256 workers
      const int ia = warp_incl_scan(gt), ib = warp_incl_scan(eq);
      if (wt >= 0 && lane == 31) { s_wa[wwarp] = ia; s_wb[wwarp] = ib; }
      __syncthreads();
      if (wwarp == 0) {
        const int xa = lane < kSsWarps ? s_wa[lane] : 0, xb = lane < kSsWarps ? s_wb[lane] : 0;
        const int sa = warp_incl_scan(xa), sb = warp_incl_scan(xb);
        if (lane < kSsWarps) { s_wa[lane] = sa - xa; s_wb[lane] = sb - xb; }
        if (lane == 31) { s_wa[kSsWarps] = sa; s_wb[kSsWarps] = sb; }
      }
      __syncthreads();
      if (wt >= 0) {
        int gt_seen = gt_before + s_wa[wwarp] + ia - gt, eq_seen = eq_before + s_wb[wwarp] + ib - eq;
        if (gt | eq) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const unsigned k = (kw[e >> 1] >> ((e & 1) * 16)) & 0xFFFFu;
            const bool real = off + e < count;
            const bool take = real && (k > thr || (k == thr && eq_seen < need_eq));
            if (take) dst[gt_seen + min(eq_seen, need_eq)] = base == 0 ? ids0[e] : static_cast<int>(lidx[off + e]);
            gt_seen += real && k > thr; eq_seen += real && k == thr;
          }
        }
      }
      gt_before += s_wa[kSsWarps]; eq_before += s_wb[kSsWarps];
      __syncthreads();
    }
  }
  if (tid == 0) SKV_STAMP_NS(tr, 10);
  cluster.sync();
  if (tid == 0) SKV_STAMP_NS(tr, 11);
  if (rank != 0) return;

  // =========================== chunk-table update (rank 0; threads beyond S carry no element) ======================
  int new_id = 0;
  const int ct = wt >= 0 ? wt : S + tid;                     // warps 0-1 take ids >= S: inactive in the update
  if (wt >= 0 && wt < S) {
    new_id = s_new[wt];
    if (p.selected_out) p.selected_out[size_t(bh) * S + wt] = static_cast<int64_t>(new_id);
  }
  ChunkTableSmem sm{s_sorted, s_slot, s_warp};
  chunk_table_update<S, kInPlace>(ct, old_id, new_id, sm, p.resident + size_t(bh) * S, p.offsets + size_t(bh) * S,
                                  p.cnts + bh);
  if (tid == 0) SKV_STAMP_NS(tr, 12);
}

int make_tensor_map_2d(CUtensorMap *out, const void *base, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_cols);

static int ss_align(int x, int a) { return (x + a - 1) / a * a; }

template <int S, bool kInPlace>
static int launch_score_select_t(const CUtensorMap &map, ScoreSelectParams p, int BH, int csz, int smem, cudaStream_t st) {
  const void *fn = reinterpret_cast<const void *>(score_select_kernel<S, kInPlace>);
  if (int rc = ensure_dynamic_smem(fn, smem, "score_select_kernel")) return rc;
  if (csz > 8) {
    static std::once_flag once;
    std::call_once(once, [fn] { cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1); });
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(csz, BH);
  cfg.blockDim = dim3(kSsThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = csz; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, score_select_kernel<S, kInPlace>, map, p);
  if (e != cudaSuccess) { set_error("score_select_kernel: %s", cudaGetErrorString(e)); return int(e); }
  return launch_status("score_select_kernel");
}

// Returns SKV_OK, an error, or SKV_ERR_UNSUPPORTED when the shape does not fit this kernel (the caller then takes the
// two-launch path).
int launch_score_select(const skv_geom *g, const skv_bf16 *q, const skv_bf16 *landmarks, const int64_t *landmark_idx,
                        int64_t *resident, int32_t *offsets, int32_t *cnts, int64_t *selected_out, skv_bf16 *score_out,
                        const skv_append *append, cudaStream_t st) {
  const int BH = g->batch * g->kv_heads, G = g->q_groups, N = g->n_landmarks, S = g->select;
  if (S > 256 || G > 16) return SKV_ERR_UNSUPPORTED;
  const int T = (N + kSsTileRows - 1) / kSsTileRows;
  // 16-CTA clusters (one per GPC) while the rows are few: 128 CTAs at batch 1 x 8 kv heads keep HBM busy
  static const int force = [] { const char *e = getenv("SKV_SELECT_CLUSTER"); return (e && *e) ? atoi(e) : 0; }();
  int csz = force ? force : (BH <= 9 ? 16 : 8);
  if (csz != 8 && csz != 16) csz = 8;
  const int tpc = (T + csz - 1) / csz;
  if (tpc > kSsMaxTilesPerCta) return SKV_ERR_UNSUPPORTED;
  ScoreSelectParams p{};
  p.q = q; p.landmark_idx = landmark_idx; p.resident = resident; p.offsets = offsets; p.cnts = cnts;
  p.selected_out = selected_out; p.score_out = score_out;
  p.G = G; p.N = N; p.T = T; p.tiles_per_cta = tpc; p.alpha = 0.08838834764831845f;   // 1/sqrt(128), kv_cache.py:1017
  p.dynamic_lm = (g->flags & SKV_GEOM_DYNAMIC_LANDMARKS) ? 1 : 0;
  p.buf_rows = g->buf_len;
  if (append) p.ap = *append;
  p.ld_sc = ss_align(tpc * kSsTileRows, 8);
  // shared-memory plan: [ring | q | barriers | remote-written: stats, histograms, selection | scores | keys | partials | small]
  const int fixed_after_ring =
      2 * kSsQSlabBytes + 256 /*barriers*/ + kSsMaxCluster * 16 * 2 * 4 + 2 * kSsMaxCluster * 256 * 4 + 256 * 4 +
      ss_align(G * p.ld_sc * 2, 16) + ss_align(p.ld_sc * 2, 16) + ss_align(2 * 16 * tpc * 4, 16) +
      (2 * 4 * 32 + 32) * 4 + (256 + 4 + 2 * (kSsWarps + 1)) * 4 + (2 * 256 + 33) * 4 + 256;
  int stages = kSsMaxStages;
  while (stages > 2 && stages * kSsTileBytes + fixed_after_ring > 227 * 1024) --stages;
  if (stages * kSsTileBytes + fixed_after_ring > 227 * 1024 || stages * kSsTileBytes < 2 * kSsWarps * 256 * 4)
    return SKV_ERR_UNSUPPORTED;
  p.stages = stages;
  int off = stages * kSsTileBytes;
  p.off_q = off; off += 2 * kSsQSlabBytes;
  p.off_bars = off; off += 256;
  p.off_stat = off; off += kSsMaxCluster * 16 * 2 * 4;
  p.off_parts = off; off += 2 * kSsMaxCluster * 256 * 4;
  p.off_new = off; off += 256 * 4;
  p.off_sc = off; off += ss_align(G * p.ld_sc * 2, 16);
  p.off_keys = off; off += ss_align(p.ld_sc * 2, 16);
  p.off_pm = off; off += ss_align(2 * 16 * tpc * 4, 16);
  p.off_small = off; off += (2 * 4 * 32 + 32) * 4 + (256 + 4 + 2 * (kSsWarps + 1)) * 4 + (2 * 256 + 33) * 4 + 256;
  const int smem = off;
  CUtensorMap map;
  if (int rc = make_tensor_map_2d(&map, landmarks, uint64_t(BH) * N, kHeadDim, kSsTileRows, 64)) return rc;
  const bool ip = g->table_mode != 0;
  switch (S) {
    case 128: return ip ? launch_score_select_t<128, true>(map, p, BH, csz, smem, st) : launch_score_select_t<128, false>(map, p, BH, csz, smem, st);
    case 256: return ip ? launch_score_select_t<256, true>(map, p, BH, csz, smem, st) : launch_score_select_t<256, false>(map, p, BH, csz, smem, st);
    default: return SKV_ERR_UNSUPPORTED;
  }
}

}  // namespace skv

SKV_TRACE_READER(score_select)

// Tuning aid: how many clusters of `csz` CTAs with `smem` bytes of dynamic shared memory the device holds at once.
extern "C" int skv_debug_max_active_clusters(int csz, int smem) {
  using namespace skv;
  const void *fn = reinterpret_cast<const void *>(score_select_kernel<256, true>);
  cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (csz > 8) cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(csz, 64);
  cfg.blockDim = dim3(kSsThreads);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = csz; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = -1;
  cudaError_t e = cudaOccupancyMaxActiveClusters(&n, score_select_kernel<256, true>, &cfg);
  return e == cudaSuccess ? n : -int(e);
}

