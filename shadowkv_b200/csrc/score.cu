// Landmark scoring and chunk selection.
//
//  score_tiles_kernel : streams the landmark matrix once (HBM bound: N*D*2 bytes per (b,h)), computes the
//                       G query-head dot products per landmark, scales, rounds to bf16 (the reference's
//                       rounding point, batch_gemm_softmax.cu:260-281) and leaves per-tile softmax partials.
//  select_kernel      : per (b,h): final softmax + bf16 round (batch_gemm_softmax.h:289-292), max over the G
//                       heads (kv_cache.py:1024-1027), top-S by 2-level radix select with lowest-index tie
//                       break (kv_cache.py:1031-1035), chunk-id gather (:1039-1042) and the chunk-table update
//                       (map.cuh:754-796) -- replaces 3 CUTLASS launches + torch.max + torch.topk + gather +
//                       reorder of the reference with two launches.
//  softmax_apply_kernel: only for the drop-in batch_gemm_softmax entry point (materialises Softmax).
#include <algorithm>
#include <cooperative_groups.h>

#include "chunk_table.cuh"
#include "common.cuh"

namespace cg = cooperative_groups;

namespace skv {

// score_select.cu: scoring + selection in one cluster launch (experiment, SKV_FUSED_SELECT=1; SKV_ERR_UNSUPPORTED: shape not taken)
int launch_score_select(const skv_geom *g, const skv_bf16 *q, const skv_bf16 *landmarks, const int64_t *landmark_idx,
                        int64_t *resident, int32_t *offsets, int32_t *cnts, int64_t *selected_out, skv_bf16 *score_out,
                        const skv_append *append, cudaStream_t st);

// score_mma.cu: TMA + tcgen05 scoring (product path)
int launch_scores_tcgen05(const skv_bf16 *q, const skv_bf16 *lm, skv_bf16 *scores, float *pmax, float *psum, int BH, int G,
                          int N, int ldn, int T, float alpha, cudaStream_t st, const skv_append *append,
                          int64_t buf_rows, bool dynamic_landmarks, const int32_t *wait_flags = nullptr,
                          const int32_t *wait_epoch_dev = nullptr, int wait_count = 0, bool wait_only = false);

constexpr int kScoreThreads = 256;   // 8 warps

// One warp handles kRowsPerWarp landmarks in groups of 16 (8 x 512 B loads in flight per lane group).
// GT query heads per pass (4 or 8); heads beyond G are zero rows of q.
template <int GT, int kRowsPerWarp>
__global__ void __launch_bounds__(kScoreThreads)
    score_tiles_kernel(const skv_bf16 *__restrict__ q,          // [BH, G, 128]
                       const skv_bf16 *__restrict__ landmarks,  // [BH, N, 128]
                       skv_bf16 *__restrict__ scores,           // [BH, G, N]
                       float *__restrict__ part_max,            // [BH, G, T]
                       float *__restrict__ part_sum,            // [BH, G, T]
                       int G, int N, int ldn, int T, float alpha,       // ldn: row stride of `scores` (>= N)
                       skv_append ap, int64_t buf_rows) {
  constexpr int kTileRows = kRowsPerWarp * (kScoreThreads / 32);
  // fused append (a2): the first CTA of every (b,h) writes the new token's K and V rows
  if (ap.k_new && blockIdx.x == 0 && blockIdx.z == 0 && threadIdx.x < 32) {
    const int64_t pos = ap.pos + (ap.pos_dev ? *ap.pos_dev : 0);
    if (pos < buf_rows) {
      const int c = threadIdx.x & 15;
      const int4 *src = reinterpret_cast<const int4 *>(threadIdx.x < 16 ? ap.k_new : ap.v_new) + size_t(blockIdx.y) * 16 + c;
      int4 *dst = reinterpret_cast<int4 *>(threadIdx.x < 16 ? ap.k_buf : ap.v_buf) + (size_t(blockIdx.y) * buf_rows + pos) * 16 + c;
      *dst = *src;
    }
  }
  __shared__ float s_tile[GT][kTileRows];
  const int tile = blockIdx.x, bh = blockIdx.y, g0 = blockIdx.z * GT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane & 15, hw = lane >> 4;

  // q slice of this lane: dims [sub*8, sub*8+8) of GT heads, as f32
  float qf[GT][8];
#pragma unroll
  for (int g = 0; g < GT; ++g) {
    if (g0 + g < G) {
      const int4 w = *reinterpret_cast<const int4 *>(q + (size_t(bh) * G + g0 + g) * kHeadDim + sub * 8);
      const uint32_t u[4] = {(uint32_t)w.x, (uint32_t)w.y, (uint32_t)w.z, (uint32_t)w.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) { qf[g][2 * e] = bf16_lo(u[e]); qf[g][2 * e + 1] = bf16_hi(u[e]); }
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) qf[g][e] = 0.f;
    }
  }

  const int row_base = tile * kTileRows + warp * kRowsPerWarp;
  const int4 *lm = reinterpret_cast<const int4 *>(landmarks + size_t(bh) * N * kHeadDim);

#pragma unroll 1
  for (int grp = 0; grp < kRowsPerWarp / 16; ++grp) {
    const int r0 = row_base + grp * 16;
    int4 ld[8];
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int row = r0 + it * 2 + hw;
      ld[it] = row < N ? ld_stream(lm + size_t(row) * 16 + sub) : make_int4(0, 0, 0, 0);
    }
    float v[8][GT];
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const uint32_t u[4] = {(uint32_t)ld[it].x, (uint32_t)ld[it].y, (uint32_t)ld[it].z, (uint32_t)ld[it].w};
      float x[8];
#pragma unroll
      for (int e = 0; e < 4; ++e) { x[2 * e] = bf16_lo(u[e]); x[2 * e + 1] = bf16_hi(u[e]); }
#pragma unroll
      for (int g = 0; g < GT; ++g) {
        float a = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) a = fmaf(x[e], qf[g][e], a);
        v[it][g] = a;
      }
    }
    // transposing butterfly over the 16 lanes of a half-warp: 7.5*GT shuffles instead of 32*GT
    float a4[4][GT];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int g = 0; g < GT; ++g) {
        const float keep = (sub & 8) ? v[i + 4][g] : v[i][g];
        const float send = (sub & 8) ? v[i][g] : v[i + 4][g];
        a4[i][g] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
      }
    float a2[2][GT];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int g = 0; g < GT; ++g) {
        const float keep = (sub & 4) ? a4[i + 2][g] : a4[i][g];
        const float send = (sub & 4) ? a4[i][g] : a4[i + 2][g];
        a2[i][g] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
      }
    float a1[GT];
#pragma unroll
    for (int g = 0; g < GT; ++g) {
      const float keep = (sub & 2) ? a2[1][g] : a2[0][g];
      const float send = (sub & 2) ? a2[0][g] : a2[1][g];
      a1[g] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    float fin[GT / 2];
#pragma unroll
    for (int g = 0; g < GT / 2; ++g) {
      const float keep = (sub & 1) ? a1[g + GT / 2] : a1[g];
      const float send = (sub & 1) ? a1[g] : a1[g + GT / 2];
      fin[g] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
    }
    // this lane now owns row `it`, heads gsel..gsel+GT/2
    const int it = ((sub >> 3) & 1) * 4 + ((sub >> 2) & 1) * 2 + ((sub >> 1) & 1);
    const int row = r0 + it * 2 + hw;
    const int lrow = row - tile * kTileRows;
    const int gsel = (sub & 1) * (GT / 2);
#pragma unroll
    for (int k = 0; k < GT / 2; ++k) {
      const int g = gsel + k;
      const uint16_t bits = f32_to_bf16_bits(fin[k] * alpha);
      float sv = -INFINITY;
      if (row < N && g0 + g < G) {
        scores[(size_t(bh) * G + g0 + g) * ldn + row] = bits;
        sv = bf16_bits_to_f32(bits);
      }
      s_tile[g][lrow] = sv;
    }
  }
  __syncthreads();
  // per-tile softmax partials on the bf16-rounded scores: warp g reduces head g
  for (int g = warp; g < GT; g += kScoreThreads / 32) {
    if (g0 + g >= G) continue;
    float m = -INFINITY;
    for (int i = lane; i < kTileRows; i += 32) m = fmaxf(m, s_tile[g][i]);
    m = warp_max(m);
    float s = 0.f;
    for (int i = lane; i < kTileRows; i += 32) {
      const float x = s_tile[g][i];
      s += (x == -INFINITY) ? 0.f : expf(x - m);
    }
    s = warp_sum(s);
    if (lane == 0) {
      part_max[(size_t(bh) * G + g0 + g) * T + tile] = m;
      part_sum[(size_t(bh) * G + g0 + g) * T + tile] = s;
    }
  }
}

// Final (max, 1/sum) of one softmax row from its tile partials; executed by one warp.
__device__ __forceinline__ void reduce_partials(const float *pm, const float *ps, int T, int lane, float &m_out,
                                                float &inv_out) {
  float m = -INFINITY;
  for (int t = lane; t < T; t += 32) m = fmaxf(m, pm[t]);
  m = warp_max(m);
  float s = 0.f;
  for (int t = lane; t < T; t += 32) s += ps[t] * expf(pm[t] - m);
  s = warp_sum(s);
  m_out = m;
  inv_out = 1.0f / s;
}

__global__ void __launch_bounds__(256)
    softmax_apply_kernel(const skv_bf16 *__restrict__ scores, const float *__restrict__ part_max,
                         const float *__restrict__ part_sum, skv_bf16 *__restrict__ softmax_o, int M, int N, int T) {
  __shared__ float s_m[16], s_inv[16];
  const int bc = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int g = warp; g < M; g += 8) {
    float m, inv;
    reduce_partials(part_max + (size_t(bc) * M + g) * T, part_sum + (size_t(bc) * M + g) * T, T, lane, m, inv);
    if (lane == 0) { s_m[g] = m; s_inv[g] = inv; }
  }
  __syncthreads();
  const int n0 = blockIdx.x * 2048;
  for (int g = 0; g < M; ++g) {
    const float m = s_m[g], inv = s_inv[g];
    const skv_bf16 *src = scores + (size_t(bc) * M + g) * N;
    skv_bf16 *dst = softmax_o + (size_t(bc) * M + g) * N;
    for (int n = n0 + threadIdx.x; n < min(N, n0 + 2048); n += 256)
      dst[n] = f32_to_bf16_bits(expf(bf16_bits_to_f32(src[n]) - m) * inv);
  }
}

// ------------------------------------------------------------------------------------------------
// select_kernel: one cluster of 8 CTAs per (b,h).  Every CTA owns a contiguous eighth of the landmarks:
// it turns the bf16 scores of the G heads into the bf16 softmax value, keeps the max over the heads as a
// 16-bit key in shared memory, and builds warp-private 256-bin histograms (no shared-memory atomics; a
// softmax row has only a handful of distinct exponents, so a shared histogram serialises on 3-4 bins).
// The 8 partial histograms are summed through distributed shared memory, so every CTA finds the same
// two-level radix threshold; the selected landmarks are then emitted in ascending index order straight
// into CTA 0's shared memory, and CTA 0 runs the chunk-table update.
// ------------------------------------------------------------------------------------------------
constexpr int kSelCluster = 8;
constexpr int kSelSpreadBytes = 100 * 1024;   // dynamic shared memory of a "spread" launch: + 15.3 KiB static + 1 KiB reserved
                                              // = 116.3 KiB per CTA: more than half of an SM's 228 KiB, and a CTA of the
                                              // fused fetch launch (109 KiB + 1 at rank 160) still fits next to it

// warp-private histogram: after __match_any_sync the leaders of distinct bins touch distinct words
__device__ __forceinline__ void hist_add_private(int *hist, unsigned bin) {
  const unsigned peers = __match_any_sync(0xffffffffu, bin);
  if (bin != 0xFFFFu && (threadIdx.x & 31) == (__ffs(peers) - 1)) hist[bin] += __popc(peers);
  __syncwarp();
}

// Warp 0: find the highest bin b with (count of bins above b) < want <= (count of bins >= b).
// Writes res[0] = b, res[1] = count strictly above b.
__device__ __forceinline__ void find_threshold_bin(const int *hist, int want, int *res) {
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

// The same search over the SUM of the cluster's partial histograms, by all threads of the CTA (kT >= 256): warp w < 8 owns
// the 32 bins 255 - 32 w - lane (descending, bank-conflict free), adds the partials, scans; after one barrier the warp
// whose segment holds the threshold writes res[0] = bin, res[1] = count strictly above it.  Ends with a barrier.
__device__ __forceinline__ int hist_count(const uint32_t *row, int bin) { return int((row[bin >> 1] >> ((bin & 1) * 16)) & 0xFFFFu); }
__device__ __forceinline__ void find_threshold_parts(const uint32_t (*parts)[128], int want, int *res, int *s_seg) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int v = 0, incl = 0;
  if (warp < 8) {
    const int bin = 255 - (warp * 32 + lane);
#pragma unroll
    for (int r = 0; r < kSelCluster; ++r) v += hist_count(parts[r], bin);
    incl = warp_incl_scan(v);
    if (lane == 31) s_seg[warp] = incl;
  }
  __syncthreads();
  if (warp < 8) {
    int carry = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) carry += (w < warp) ? s_seg[w] : 0;
    if (carry < want && want <= carry + s_seg[warp]) {       // warp uniform: the threshold lies in this warp's segment
      const unsigned hit = __ballot_sync(0xffffffffu, carry + incl >= want);
      if (lane == __ffs(hit) - 1) { res[0] = 255 - (warp * 32 + lane); res[1] = carry + incl - v; }
    }
  }
  __syncthreads();
}

// Final (max, 1/sum) with both partial arrays fetched in one round trip (<= 8 x 32 tiles per lane in
// registers; longer rows fall back to reduce_partials).
__device__ __forceinline__ void reduce_partials_fast(const float *pm, const float *ps, int T, int lane, float &m_out,
                                                     float &inv_out) {
  if (T > 256) { reduce_partials(pm, ps, T, lane, m_out, inv_out); return; }
  float vm[8], vs[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int t = lane + 32 * i;
    vm[i] = t < T ? pm[t] : -INFINITY;
    vs[i] = t < T ? ps[t] : 0.f;
  }
  float m = vm[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) m = fmaxf(m, vm[i]);
  m = warp_max(m);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) sum += vs[i] * ((vm[i] == -INFINITY) ? 0.f : expf(vm[i] - m));
  sum = warp_sum(sum);
  m_out = m;
  inv_out = 1.0f / sum;
}

// exclusive block scan of two ints at once; returns the exclusive prefixes, `tot_*` the block sums
template <int kWarps>
__device__ __forceinline__ void block_excl_scan2(int a, int b, int *wa, int *wb, int &pa, int &pb, int &tot_a,
                                                 int &tot_b) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ia = warp_incl_scan(a), ib = warp_incl_scan(b);
  __syncthreads();                        // wa / wb may still be read from a previous round
  if (lane == 31) { wa[warp] = ia; wb[warp] = ib; }
  __syncthreads();
  if (warp == 0) {
    const int xa = lane < kWarps ? wa[lane] : 0, xb = lane < kWarps ? wb[lane] : 0;
    const int sa = warp_incl_scan(xa), sb = warp_incl_scan(xb);
    if (lane < kWarps) { wa[lane] = sa - xa; wb[lane] = sb - xb; }
    if (lane == 31) { wa[kWarps] = sa; wb[kWarps] = sb; }
  }
  __syncthreads();
  pa = wa[warp] + ia - a;
  pb = wb[warp] + ib - b;
  tot_a = wa[kWarps];
  tot_b = wb[kWarps];
}

// kT threads per CTA: max(S, 256) normally; 512 when a CTA's slice is long (contexts of 256K tokens and more), where the
// key pass (G exponentials per landmark) is what the kernel spends its time on.
// kCap: 64-register cap, so that 1024 threads of this kernel fit an SM (4 CTAs of 256).  It pays when the grid has more
// CTAs than the GPU has SMs (batch 8: 20.4 -> 15.4 us); with fewer (batch 1: 64 CTAs) the uncapped build is the faster
// one (the score + select pair 18.4 against 20.5 us), so the launcher picks per grid size.
template <int S, bool kInPlace, int kT, bool kCap>
__global__ void __cluster_dims__(kSelCluster, 1, 1) __launch_bounds__(kT, kCap ? 1024 / kT : 1)
    select_kernel(const skv_bf16 *__restrict__ scores,        // [BH, G, ldn] bf16 scaled dot products
                  const float *__restrict__ part_max, const float *__restrict__ part_sum,   // [BH, G, T]
                  const int64_t *__restrict__ landmark_idx,   // [BH, N]
                  int64_t *__restrict__ resident,             // [BH, S] in/out
                  int32_t *__restrict__ offsets, int32_t *__restrict__ cnts,
                  int64_t *__restrict__ selected_out,         // [BH, S] or null
                  skv_bf16 *__restrict__ score_out,           // [BH, N] or null
                  int32_t *row_flags,                         // [BH] or null: per-row hand-over to the fetch launch
                  int32_t *hint,                              // [BH] or null: the row's previous threshold key (in/out)
                  const int32_t *fast_status,                 // [BH] or null: 1 = select_fast_kernel (launched in front of
                                                              // this grid) has done the row; it also cleared row_flags
                  int G, int N, int ldn, int T, int per, int filter_on) {
  constexpr int kThreads = kT;
  static_assert(kT >= S && kT >= 256 && kT % 32 == 0, "one thread per table entry and per histogram bin");
  constexpr int kWarps = kThreads / 32;
  extern __shared__ __align__(16) unsigned char sel_smem[];
  int *whist1 = reinterpret_cast<int *>(sel_smem);                                 // [kWarps][256] level 1
  int *whist2 = whist1 + kWarps * 256;                                             // [kWarps][256] level 2
  uint16_t *keys = reinterpret_cast<uint16_t *>(sel_smem + 2 * kWarps * 256 * 4);  // [per rounded to 8]
  __shared__ float s_m[16], s_inv[16], s_linv[16];
  // level-1 / level-2 histograms of all 8 CTAs (each CTA stores its own into every CTA), two 16-bit counts per word:
  // a CTA's slice has fewer than 65536 landmarks (check_select_geom), and the exchange is 4 KiB per CTA instead of 8
  __shared__ uint32_t s_parts1[kSelCluster][128];
  __shared__ uint32_t s_parts2[kSelCluster][128];
  __shared__ uint32_t s_parts0[kSelCluster][128];   // the window histogram (threshold hint), same layout
  __shared__ int s_res[4];
  __shared__ int s_wa[kWarps + 1], s_wb[kWarps + 1];
  __shared__ int s_new[S];                      // CTA 0: selected chunk ids in ascending landmark order
  __shared__ int s_sorted[S], s_slot[S], s_warp[33];

  cg::cluster_group cluster = cg::this_cluster();
  const int rank = blockIdx.x;                  // the cluster spans blockIdx.x
  const int bh = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_begin = min(N, rank * per), n_end = min(N, n_begin + per);
  const int count = n_end - n_begin;            // only the last slice can be ragged (N % 8 != 0)
  SKV_STAMP(bh == 0 ? rank : 999, 0);
  if (tid == 0 && rank == 0) SKV_STAMP_NS(bh == 0 ? 17 : -1, 0);

  // early, independent global loads: the (static) landmark ids of this thread's first 8 landmarks.  Under a
  // programmatic launch this and the histogram clearing overlap the tail of the scoring kernel.
  if (!row_flags || fast_status) pdl_launch_dependents();
  const int64_t *lidx = landmark_idx + size_t(bh) * N + n_begin;
  int ids0[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) ids0[e] = (tid * 8 + e < count) ? static_cast<int>(lidx[tid * 8 + e]) : 0;   // map.cuh:776
  for (int i = tid; i < 2 * kWarps * 256; i += kThreads) whist1[i] = 0;
  pdl_wait();                                   // scores, partials and the chunk table are the predecessors' outputs
  // Rows the single-CTA fast selection has settled (its launch sits between the scoring launch and this one) are done,
  // flag raised and all: the whole cluster leaves (uniform: one word per row), before any cluster barrier.
  if (fast_status && __ldcg(fast_status + bh) == 1) return;
  if (tid == 0 && rank == 0) SKV_STAMP_NS(bh == 0 ? 0 : -1, 21);
  // (the loads below are requested BEFORE the row word is cleared: the fence behind the clear then overlaps their L2 round
  // trip instead of preceding it in CTA 0, the CTA that ends the chain)
  int old_id = 0;
  if (rank == 0 && tid < S) old_id = static_cast<int>(resident[size_t(bh) * S + tid]);                       // map.cuh:771
  // the scores of this thread's first 8 landmarks for the first four heads, requested together with the partial
  // statistics below: one L2 round trip instead of two in front of the first histogram pass
  const skv_bf16 *sc = scores + size_t(bh) * G * ldn;
  // Threshold hint: the key this row's selection ended on for the previous token.  The first histogram then has ONE KEY
  // VALUE per bin in a window around it (bins 1..254 = keys win0 .. win0 + 253, bin 0 = everything below, bin 255 =
  // everything above), so when the new threshold lies inside the window -- consecutive tokens move it by a few bf16
  // steps, the window spans a factor of two either way -- ONE histogram exchange settles it exactly; otherwise the
  // two-level search below runs as if there had been no hint.  The result never depends on the hint.
  const int hint_key = hint ? __ldcg(hint + bh) : 0;
  const int win0 = hint_key > 0 ? max(1, hint_key - 126) : 0;          // 0: no window
  int4 w0[4] = {make_int4(0, 0, 0, 0), make_int4(0, 0, 0, 0), make_int4(0, 0, 0, 0), make_int4(0, 0, 0, 0)};
  if (tid * 8 < count) {
#pragma unroll
    for (int g = 0; g < 4; ++g)
      if (g < G) w0[g] = *reinterpret_cast<const int4 *>(sc + size_t(g) * ldn + n_begin + tid * 8);
  }

  if (row_flags && !fast_status) {
    // Row hand-over (skv_geom.row_flags): the dependents of this launch wait for their ROW's word, not for this grid.
    // The word still carries the previous token's 1; every consumer of that value has finished (this kernel's wait
    // covers, transitively, the previous launch that read it), and no consumer of the new value exists before the
    // trigger below: clear it, make the store visible, and only then let the dependent grid be scheduled.
    // The clear and its fence are done by the first lane of the LAST warp, which has no head to reduce below when
    // G < 8: the fence overlaps the L2 round trip of the statistics instead of standing in front of it; the trigger
    // follows the barrier that ends the statistics (the clearing thread is past its fence by then).
    if (rank == 0 && tid == kThreads - 32) {
      asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(row_flags + bh), "r"(0) : "memory");
      __threadfence();
    }
  }

  // 0. final softmax statistics per head (every CTA redundantly: G x T floats)
  for (int g = warp; g < G; g += kWarps) {
    float m, inv;
    reduce_partials_fast(part_max + (size_t(bh) * G + g) * T, part_sum + (size_t(bh) * G + g) * T, T, lane, m, inv);
    if (lane == 0) { s_m[g] = m; s_inv[g] = inv; s_linv[g] = logf(inv); }
  }
  __syncthreads();
  if (row_flags && !fast_status) pdl_launch_dependents();
  SKV_STAMP(bh == 0 ? rank : 999, 1);
  if (tid == 0 && rank == 0) SKV_STAMP_NS(bh == 0 ? 17 : -1, 1);

  // 1. p = bf16(exp(s - max) * inv); key = max over heads (bf16 bits are monotone for p >= 0);
  //    first-level histogram on the high byte.  8 consecutive landmarks per thread: 16-byte loads.
  int *my_h1 = whist1 + warp * 256, *my_h2 = whist2 + warp * 256;
  // all keys of round `base` exactly (G exponentials per landmark) into `keys`; with_hist: also the first histogram
  auto full_keys_round = [&](const int base, const bool first_regs, const bool with_hist) {
    const int off = base + tid * 8;
    const bool on = off < count;
    uint16_t best[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (on) {
      const int n = n_begin + off;
      auto fold = [&](const int4 &w, const int g) {
        const uint32_t u[4] = {(uint32_t)w.x, (uint32_t)w.y, (uint32_t)w.z, (uint32_t)w.w};
        const float m = s_m[g], inv = s_inv[g];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          // expf, not __expf: the oracle (and the tile partials of the scoring kernel) use the IEEE-accurate exponential,
          // and a 2-ulp fp32 difference here flips bf16 roundings at the selection threshold
          const uint16_t b0 = f32_to_bf16_bits(expf(bf16_lo(u[e]) - m) * inv);
          const uint16_t b1 = f32_to_bf16_bits(expf(bf16_hi(u[e]) - m) * inv);
          best[2 * e] = b0 > best[2 * e] ? b0 : best[2 * e];
          best[2 * e + 1] = b1 > best[2 * e + 1] ? b1 : best[2 * e + 1];
        }
      };
#pragma unroll
      for (int g = 0; g < 4; ++g)                          // the first round's loads are already in registers
        if (g < G) fold(first_regs ? w0[g] : *reinterpret_cast<const int4 *>(sc + size_t(g) * ldn + n), g);
#pragma unroll 4
      for (int g = 4; g < G; ++g) fold(*reinterpret_cast<const int4 *>(sc + size_t(g) * ldn + n), g);
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (off + e >= count) best[e] = 0;                    // ragged tail: never selected (p > 0 for real keys)
      const int4 packed = make_int4(int(best[0] | (uint32_t(best[1]) << 16)), int(best[2] | (uint32_t(best[3]) << 16)),
                                    int(best[4] | (uint32_t(best[5]) << 16)), int(best[6] | (uint32_t(best[7]) << 16)));
      *reinterpret_cast<int4 *>(keys + off) = packed;
      if (score_out) {
        if (ldn == N) {
          *reinterpret_cast<int4 *>(score_out + size_t(bh) * N + n) = packed;
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if (off + e < count) score_out[size_t(bh) * N + n + e] = best[e];
        }
      }
    }
    if (with_hist) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int k = best[e];
        const unsigned bin = win0 ? (k < win0 ? 0u : unsigned(min(255, k - win0 + 1))) : unsigned(k >> 8);
        hist_add_private(my_h1, (on && off + e < count) ? bin : 0xFFFFu);
      }
    }
  };
  // With a usable hint the exponentials are only taken for CANDIDATES: log p_g = (s_g - m_g) + log(1 / Z_g) costs one
  // add per score, and a landmark whose max_g log p_g lies below log(value of key win0 - 1) - 1e-3 provably has a key
  // below the window (the margin covers the fp32 rounding of log p_g, of logf and of the product expf(.) * inv by two
  // orders of magnitude; a product below the bf16 number win0 - 1 cannot round up to win0; the window stays clear of
  // products that could be denormal).  Such a landmark can neither be selected nor move a threshold that lies inside the
  // window: it is counted in bin 0 with key 0.  The few hundred candidates of the row (everything within about a factor
  // of two of the previous threshold) are compacted per warp and get their exact key.  If the threshold turns out to lie
  // outside the window, all keys are formed after all (below) and the two-level search runs as without a hint.
  const bool filt = filter_on && win0 >= 0x1000 && score_out == nullptr;
  if (!filt) {
    for (int base = 0; base < count; base += kThreads * 8) full_keys_round(base, base == 0, true);
  } else {
    const float cut = logf(bf16_bits_to_f32(uint16_t(win0 - 1))) - 1e-3f;
    uint8_t *wl = reinterpret_cast<uint8_t *>(my_h2);       // this warp's candidate list (whist2 is re-cleared before use)
    uint16_t *sraw = keys + per;                            // [G][per] raw scores of the slice (per is a multiple of 8)
    for (int base = 0; base < count; base += kThreads * 8) {
      const int off = base + tid * 8;
      const bool on = off < count;
      float umax[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) umax[e] = -INFINITY;
      if (on) {
        const int n = n_begin + off;
        auto fold_u = [&](const int4 &w, const int g) {
          const uint32_t u[4] = {(uint32_t)w.x, (uint32_t)w.y, (uint32_t)w.z, (uint32_t)w.w};
          const float m = s_m[g], li = s_linv[g];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            umax[2 * e] = fmaxf(umax[2 * e], (bf16_lo(u[e]) - m) + li);
            umax[2 * e + 1] = fmaxf(umax[2 * e + 1], (bf16_hi(u[e]) - m) + li);
          }
        };
#pragma unroll
        for (int g = 0; g < 4; ++g)
          if (g < G) {
            const int4 w = base == 0 ? w0[g] : *reinterpret_cast<const int4 *>(sc + size_t(g) * ldn + n);
            fold_u(w, g);
            *reinterpret_cast<int4 *>(sraw + g * per + off) = w;     // raw scores for the candidates' exact keys
          }
#pragma unroll 4
        for (int g = 4; g < G; ++g) {
          const int4 w = *reinterpret_cast<const int4 *>(sc + size_t(g) * ldn + n);
          fold_u(w, g);
          *reinterpret_cast<int4 *>(sraw + g * per + off) = w;
        }
        *reinterpret_cast<int4 *>(keys + off) = make_int4(0, 0, 0, 0);
      }
      __syncwarp();                                           // the zeroed keys precede the candidates' keys below
      int n_c = 0;                                            // warp uniform: candidates of this warp's 256 landmarks
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const bool c = on && off + e < count && umax[e] >= cut;
        const unsigned b = __ballot_sync(0xffffffffu, c);
        if (c) wl[n_c + __popc(b & ((1u << lane) - 1u))] = uint8_t(lane * 8 + e);
        n_c += __popc(b);
      }
      __syncwarp();
      const int wbase = base + warp * 256;                    // first landmark (slice-local) of this warp in this round
      for (int i0 = 0; i0 < n_c; i0 += 32) {
        const bool have = i0 + lane < n_c;
        unsigned bin = 0xFFFFu;
        if (have) {
          const int o = wbase + wl[i0 + lane];
          uint16_t key = 0;
          for (int g = 0; g < G; ++g) {
            const uint16_t k = f32_to_bf16_bits(expf(bf16_bits_to_f32(sraw[g * per + o]) - s_m[g]) * s_inv[g]);
            key = k > key ? k : key;
          }
          keys[o] = key;
          bin = key < win0 ? 0u : unsigned(min(255, int(key) - win0 + 1));
        }
        hist_add_private(my_h1, bin);
      }
      if (lane == 0) my_h1[0] += max(0, min(256, count - wbase)) - n_c;   // everything below the window
      __syncwarp();
    }
  }
  SKV_STAMP(bh == 0 ? rank : 999, 2);
  if (tid == 0 && rank == 0) SKV_STAMP_NS(bh == 0 ? 17 : -1, 2);
  // sum of the warp-private histograms `wh`, two 16-bit counts per word, into row `rank` of `parts` of every CTA
  auto push_hist = [&](const int *wh, uint32_t (*parts)[128]) {
    __syncthreads();
    if (tid < 128) {
      int2 t = make_int2(0, 0);
#pragma unroll
      for (int w = 0; w < kWarps; ++w) {
        const int2 x = *reinterpret_cast<const int2 *>(wh + w * 256 + 2 * tid);
        t.x += x.x; t.y += x.y;
      }
      const uint32_t pk = uint32_t(t.x) | (uint32_t(t.y) << 16);
#pragma unroll
      for (int r = 0; r < kSelCluster; ++r) cluster.map_shared_rank(&parts[rank][tid], r)[0] = pk;
    }
    cluster.sync();
  };
  unsigned thr = 0;
  int need_eq = 0, gt_before = 0, eq_before = 0;
  // (greater, equal) counts of the lower-ranked CTAs from the exchanged partial histograms: bins above `ba` of `pa` and
  // (two-level search) above `bb` of `pb` hold greater keys, bin `bb` of `pb` the keys equal to the threshold
  auto counts_before = [&](const uint32_t (*pa)[128], int ba, const uint32_t (*pb)[128], int bb) {
    int a = 0, e = 0;
    if (tid < 256) {
#pragma unroll
      for (int r = 0; r < kSelCluster; ++r) {
        if (r < rank) {
          if (pa != nullptr && tid > ba) a += hist_count(pa[r], tid);
          const int c2 = hist_count(pb[r], tid);
          if (tid > bb) a += c2;
          if (tid == bb) e += c2;
        }
      }
    }
    a = warp_sum_int(a); e = warp_sum_int(e);
    if (lane == 0) { s_wa[warp] = a; s_wb[warp] = e; }
    __syncthreads();
    for (int w = 0; w < (kWarps < 8 ? kWarps : 8); ++w) { gt_before += s_wa[w]; eq_before += s_wb[w]; }
  };
  bool settled = false;
  if (win0) {                                   // cluster uniform (the hint word is the same for all eight CTAs)
    push_hist(whist1, s_parts0);
    SKV_STAMP(bh == 0 ? rank : 999, 3);
  if (tid == 0 && rank == 0) SKV_STAMP_NS(bh == 0 ? 17 : -1, 3);
    find_threshold_parts(s_parts0, S, s_res, s_wa);
    const int b0 = s_res[0], above0 = s_res[1];
    if (b0 >= 1 && b0 <= 254) {                 // the threshold is a key of the window: done after ONE exchange
      thr = unsigned(win0 + b0 - 1);
      need_eq = S - above0;
      counts_before(nullptr, 0, s_parts0, b0);
      settled = true;
    } else {                                    // outside: the level-1 histogram of the two-level search, from the keys
      __syncthreads();
      if (filt) {                               // only the candidates have their key so far: form all of them now
        for (int base = 0; base < count; base += kThreads * 8) full_keys_round(base, false, false);
      }
      for (int i = tid; i < kWarps * 256; i += kThreads) { whist1[i] = 0; whist2[i] = 0; }   // whist2 held the lists
      __syncthreads();
      for (int base = 0; base < count; base += kThreads * 8) {
        const int off = base + tid * 8;
        const bool on = off < count;
        uint32_t kw[4] = {0, 0, 0, 0};
        if (on) {
          const int4 w = *reinterpret_cast<const int4 *>(keys + off);
          kw[0] = w.x; kw[1] = w.y; kw[2] = w.z; kw[3] = w.w;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const unsigned k = (kw[e >> 1] >> ((e & 1) * 16)) & 0xFFFFu;
          hist_add_private(my_h1, (on && off + e < count) ? (k >> 8) : 0xFFFFu);
        }
      }
    }
  }
  if (!settled) {
  push_hist(whist1, s_parts1);
  SKV_STAMP(bh == 0 ? rank : 999, 3);
  if (tid == 0 && rank == 0) SKV_STAMP_NS(bh == 0 ? 17 : -1, 3);
  find_threshold_parts(s_parts1, S, s_res, s_wa);
  const int b1 = s_res[0], above1 = s_res[1];
  SKV_STAMP(bh == 0 ? rank : 999, 4);
  if (tid == 0 && rank == 0) SKV_STAMP_NS(bh == 0 ? 17 : -1, 4);

  // 2. second level: low byte of the keys inside the threshold bin
  for (int base = 0; base < count; base += kThreads * 8) {
    const int off = base + tid * 8;
    const bool on = off < count;
    uint32_t kw[4] = {0, 0, 0, 0};
    if (on) {
      const int4 w = *reinterpret_cast<const int4 *>(keys + off);
      kw[0] = w.x; kw[1] = w.y; kw[2] = w.z; kw[3] = w.w;
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const unsigned k = (kw[e >> 1] >> ((e & 1) * 16)) & 0xFFFFu;
      hist_add_private(my_h2, (on && off + e < count && int(k >> 8) == b1) ? (k & 0xFFu) : 0xFFFFu);
    }
  }
  SKV_STAMP(bh == 0 ? rank : 999, 5);
  if (tid == 0 && rank == 0) SKV_STAMP_NS(bh == 0 ? 17 : -1, 5);
  push_hist(whist2, s_parts2);
  SKV_STAMP(bh == 0 ? rank : 999, 6);
  if (tid == 0 && rank == 0) SKV_STAMP_NS(bh == 0 ? 17 : -1, 6);
  find_threshold_parts(s_parts2, S - above1, s_res + 2, s_wa);
  thr = (unsigned(b1) << 8) | unsigned(s_res[2]);
  need_eq = S - (above1 + s_res[3]);                // how many keys equal to thr are taken (>= 1)
  SKV_STAMP(bh == 0 ? rank : 999, 7);
  if (tid == 0 && rank == 0) SKV_STAMP_NS(bh == 0 ? 17 : -1, 7);

  // 3. (greater, equal) counts of the lower-ranked CTAs: every CTA already holds all eight partial
  //    histograms, so no further exchange is needed
  counts_before(s_parts1, b1, s_parts2, s_res[2]);
  }
  SKV_STAMP(bh == 0 ? rank : 999, 10);
  if (tid == 0 && rank == 0) SKV_STAMP_NS(bh == 0 ? 17 : -1, 10);

  // 4. ordered emission (ascending landmark index): thread t owns 8 consecutive landmarks per round, so the
  //    position of a taken key = (#greater before it) + min(#equal before it, need_eq), by a block scan.
  {
    int *dst = cluster.map_shared_rank(s_new, 0);
    for (int base = 0; base < count; base += kThreads * 8) {
      const int off = base + tid * 8;
      const bool on = off < count;
      uint32_t kw[4] = {0, 0, 0, 0};
      if (on) {
        const int4 w = *reinterpret_cast<const int4 *>(keys + off);
        kw[0] = w.x; kw[1] = w.y; kw[2] = w.z; kw[3] = w.w;
      }
      int gt = 0, eq = 0;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const unsigned k = (kw[e >> 1] >> ((e & 1) * 16)) & 0xFFFFu;
        const bool real = off + e < count;                   // the ragged tail of the last slice never counts
        gt += real && k > thr; eq += real && k == thr;
      }
      int pg, pe, tg, te;
      block_excl_scan2<kWarps>(gt, eq, s_wa, s_wb, pg, pe, tg, te);
      int gt_seen = gt_before + pg, eq_seen = eq_before + pe;
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
      gt_before += tg; eq_before += te;
    }
  }
  SKV_STAMP(bh == 0 ? rank : 999, 11);
  if (tid == 0 && rank == 0) SKV_STAMP_NS(bh == 0 ? 17 : -1, 11);
  cluster.sync();
  SKV_STAMP(bh == 0 ? rank : 999, 12);
  if (tid == 0 && rank == 0) SKV_STAMP_NS(bh == 0 ? 17 : -1, 12);
  if (rank != 0) return;

  // 5. chunk-table update (CTA 0)
  int new_id = 0;
  if (tid < S) {
    new_id = s_new[tid];
    if (selected_out) selected_out[size_t(bh) * S + tid] = static_cast<int64_t>(new_id);
  }
  SKV_STAMP(bh == 0 ? rank : 999, 13);
  if (tid == 0 && rank == 0) SKV_STAMP_NS(bh == 0 ? 17 : -1, 13);
  ChunkTableSmem sm{s_sorted, s_slot, s_warp};
  chunk_table_update<S, kInPlace>(tid, old_id, new_id, sm, resident + size_t(bh) * S, offsets + size_t(bh) * S,
                                  cnts + bh);
  SKV_STAMP(bh == 0 ? rank : 999, 14);
  if (tid == 0 && rank == 0) SKV_STAMP_NS(bh == 0 ? 17 : -1, 14);
  if (hint && tid == 0) hint[bh] = int(thr);      // read by this row's selection of the next token, behind its wait
  if (row_flags) {                              // the row's chunk table (resident, offsets, cnt) is written: release it
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(row_flags + bh), "r"(1) : "memory");
    }
  }
  if (tid == 0) SKV_STAMP_NS(bh == 0 ? 0 : -1, 22);
}

// ------------------------------------------------------------------------------------------------
// select_fast_kernel: the selection of a row by ONE CTA, without a cluster, when the previous token's threshold is a good
// guess (skv_geom.select_hint).  The cluster kernel above spends its time in exchanges (three cluster barriers, DSMEM
// histograms) and in G exact exponentials per landmark.  Here:
//   1. every thread takes 16 consecutive landmarks and only forms u_g = (s_g - m_g) + log(1 / Z_g) = log p_g per head (one
//      add per score, no exponential): a landmark whose max_g u_g lies below log(value of key win0 - 1) - 1e-3 provably
//      has a key below the window [win0, win0 + 253] around the previous threshold (the margin covers the fp32 rounding
//      of u_g and of the product expf(.) * inv by two orders of magnitude, and a value below the bf16 number win0 - 1
//      cannot round up to win0), so it can neither be selected nor move the threshold when that lies in the window;
//   2. the other landmarks (a few hundred: everything within a factor of two of the threshold) are compacted in landmark
//      order and get their exact key, the same expression as in select_kernel; their window histogram lives in shared
//      memory, one warp finds the threshold bin;
//   3. ordered emission, chunk-table update, hint, row hand-over exactly as in select_kernel.
// When the row has no hint, the threshold falls outside the window, the window reaches down into keys whose products
// would be denormal, or the candidates overflow the list, the kernel touches nothing and leaves status[bh] = 0: the
// cluster kernel, launched right behind this one, does the row.  status[bh] = 1: the row is done, flag raised.
// The result never depends on which kernel produced it (test_select_fast_path_equals_cluster_kernel).
// EXPERIMENT, not the product path (SKV_SELECT_FAST=1): measured slower than the cluster kernel alone, see the launcher.
// ------------------------------------------------------------------------------------------------
constexpr int kFsThreads = 1024;
constexpr int kFsPer = 16;                       // landmarks per thread: N <= 16384
constexpr int kFsCap = 2048;                     // candidate list entries

template <int S, bool kInPlace>
__global__ void __launch_bounds__(kFsThreads, 1)
    select_fast_kernel(const skv_bf16 *__restrict__ scores, const float *__restrict__ part_max,
                       const float *__restrict__ part_sum, const int64_t *__restrict__ landmark_idx,
                       int64_t *__restrict__ resident, int32_t *__restrict__ offsets, int32_t *__restrict__ cnts,
                       int64_t *__restrict__ selected_out, int32_t *row_flags, int32_t *hint, int32_t *status, int G, int N,
                       int ldn, int T) {
  constexpr int kWarps = kFsThreads / 32;
  __shared__ float s_m[16], s_inv[16], s_linv[16];
  __shared__ int s_hist[256];
  __shared__ uint16_t s_cidx[kFsCap], s_ckey[kFsCap];     // candidate: landmark (row-local), exact key
  __shared__ int s_cval[kFsCap];                          // ... its chunk id (landmark_idx)
  __shared__ uint2 s_cs[kFsCap];                          // ... its scores of heads 0..3 (G <= 4: no second trip to L2)
  __shared__ int s_wa[kWarps + 1], s_wb[kWarps + 1], s_res[2];
  __shared__ int s_new[S], s_sorted[S], s_slot[S], s_warp[33];
  const int bh = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid < 256) s_hist[tid] = 0;
  if (tid == 0) { s_res[0] = 0; s_res[1] = 0; }
  if (tid == 0) SKV_STAMP_NS(bh == 0 ? 16 : -1, 0);
  pdl_wait();                                   // scores, partials: the scoring launch; chunk table, hint: earlier launches
  if (tid == 0) SKV_STAMP_NS(bh == 0 ? 16 : -1, 1);
  if (row_flags) {                              // see select_kernel: clear the row's word, then let the dependents come
    if (tid == 0) {
      asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(row_flags + bh), "r"(0) : "memory");
      __threadfence();
    }
    __syncthreads();
  }
  pdl_launch_dependents();
  const int hint_key = __ldcg(hint + bh);
  const int win0 = hint_key > 0 ? max(1, hint_key - 126) : 0;
  // the window must stay clear of keys whose fp32 product could be denormal (relative error bounds end there)
  const bool usable = win0 >= 0x1000;           // bf16 0x1000 = 2^-95
  int old_id = 0;
  if (usable && tid < S) old_id = static_cast<int>(resident[size_t(bh) * S + tid]);
  const skv_bf16 *sc = scores + size_t(bh) * G * ldn;
  const int n0 = tid * kFsPer;
  // the first four heads' scores of this thread's landmarks, requested together with the partial statistics
  int4 w[4][2];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    w[g][0] = make_int4(0, 0, 0, 0); w[g][1] = make_int4(0, 0, 0, 0);
    if (usable && g < G && n0 < N) {
      w[g][0] = *reinterpret_cast<const int4 *>(sc + size_t(g) * ldn + n0);
      if (n0 + 8 < ldn) w[g][1] = *reinterpret_cast<const int4 *>(sc + size_t(g) * ldn + n0 + 8);
    }
  }
  for (int g = warp; g < G; g += kWarps) {
    float m, inv;
    reduce_partials_fast(part_max + (size_t(bh) * G + g) * T, part_sum + (size_t(bh) * G + g) * T, T, lane, m, inv);
    if (lane == 0) { s_m[g] = m; s_inv[g] = inv; s_linv[g] = logf(inv); }
  }
  __syncthreads();
  if (tid == 0) SKV_STAMP_NS(bh == 0 ? 16 : -1, 2);
  if (!usable) {                                // CTA uniform
    if (tid == 0) status[bh] = -1;              // <= 0: not done (the code says why; tools/exp_select_fast.py)
    return;
  }
  // 1. candidates: max_g log p_g against the lower edge of the window
  const float cut = logf(bf16_bits_to_f32(uint16_t(win0 - 1))) - 1e-3f;
  float umax[kFsPer];
#pragma unroll
  for (int e = 0; e < kFsPer; ++e) umax[e] = -INFINITY;
  auto fold = [&](const int4 &a, const int4 &b, const int g) {
    const uint32_t u[8] = {(uint32_t)a.x, (uint32_t)a.y, (uint32_t)a.z, (uint32_t)a.w,
                           (uint32_t)b.x, (uint32_t)b.y, (uint32_t)b.z, (uint32_t)b.w};
    const float m = s_m[g], li = s_linv[g];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      umax[2 * e] = fmaxf(umax[2 * e], (bf16_lo(u[e]) - m) + li);
      umax[2 * e + 1] = fmaxf(umax[2 * e + 1], (bf16_hi(u[e]) - m) + li);
    }
  };
  if (n0 < N) {
#pragma unroll
    for (int g = 0; g < 4; ++g)
      if (g < G) fold(w[g][0], w[g][1], g);
    for (int g = 4; g < G; ++g) {
      const int4 a = *reinterpret_cast<const int4 *>(sc + size_t(g) * ldn + n0);
      const int4 b = n0 + 8 < ldn ? *reinterpret_cast<const int4 *>(sc + size_t(g) * ldn + n0 + 8) : make_int4(0, 0, 0, 0);
      fold(a, b, g);
    }
  }
  unsigned mask = 0;
#pragma unroll
  for (int e = 0; e < kFsPer; ++e) mask |= (n0 + e < N && umax[e] >= cut) ? (1u << e) : 0u;
  int total;
  int off = block_excl_scan(__popc(mask), s_warp, total);
  if (tid == 0) SKV_STAMP_NS(bh == 0 ? 16 : -1, 3);
#ifdef SKV_TRACE
  if (tid == 0 && bh == 0) g_skv_trace[16 * 32 + 12] = total;
#endif
  if (total > kFsCap || total < S) {            // CTA uniform.  Fewer than S candidates: the threshold is below the window
    if (tid == 0) status[bh] = total > kFsCap ? -2 : -3;
    return;
  }
  if (mask) {                                   // ascending landmark order
#pragma unroll
    for (int e = 0; e < kFsPer; ++e) {
      if (mask & (1u << e)) {
        s_cidx[off] = uint16_t(n0 + e);
        uint32_t b[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int4 &q4 = w[g][e >> 3];
          const int c = (e & 7) >> 1;
          const uint32_t word = uint32_t(c == 0 ? q4.x : c == 1 ? q4.y : c == 2 ? q4.z : q4.w);
          b[g] = (e & 1) ? (word >> 16) : (word & 0xFFFFu);
        }
        s_cs[off] = make_uint2(b[0] | (b[1] << 16), b[2] | (b[3] << 16));
        ++off;
      }
    }
  }
  __syncthreads();
  if (tid == 0) SKV_STAMP_NS(bh == 0 ? 16 : -1, 4);
  // 2. exact keys of the candidates (the expression of select_kernel), window histogram in shared memory
  const int64_t *lidx = landmark_idx + size_t(bh) * N;
  for (int i = tid; i < total; i += kFsThreads) {
    const int n = s_cidx[i];
    const int64_t id = lidx[n];                 // requested here, needed at the emission
    uint16_t key = 0;
    if (G <= 4) {
      const uint2 cs = s_cs[i];
      const uint16_t sb[4] = {uint16_t(cs.x & 0xFFFFu), uint16_t(cs.x >> 16), uint16_t(cs.y & 0xFFFFu), uint16_t(cs.y >> 16)};
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        if (g < G) {
          const uint16_t k = f32_to_bf16_bits(expf(bf16_bits_to_f32(sb[g]) - s_m[g]) * s_inv[g]);
          key = k > key ? k : key;
        }
      }
    } else {
      for (int g = 0; g < G; ++g) {
        const uint16_t k = f32_to_bf16_bits(expf(bf16_bits_to_f32(sc[size_t(g) * ldn + n]) - s_m[g]) * s_inv[g]);
        key = k > key ? k : key;
      }
    }
    s_ckey[i] = key;
    s_cval[i] = static_cast<int>(id);
    atomicAdd(&s_hist[key < win0 ? 0 : min(255, int(key) - win0 + 1)], 1);
  }
  __syncthreads();
  if (tid == 0) SKV_STAMP_NS(bh == 0 ? 16 : -1, 5);
  if (warp == 0) {                              // threshold bin: the largest b with count(bins >= b) >= S
    int c[8], tot = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { c[j] = s_hist[lane * 8 + j]; tot += c[j]; }
    int suf = tot;                              // sum over lanes >= lane
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_down_sync(0xffffffffu, suf, o);
      if (lane + o < 32) suf += v;
    }
    const int above_lane = suf - tot;
    if (above_lane < S && suf >= S) {
      int acc = above_lane;
#pragma unroll
      for (int j = 7; j >= 0; --j) {
        if (acc < S && acc + c[j] >= S) { s_res[0] = lane * 8 + j; s_res[1] = acc; }
        acc += c[j];
      }
    }
  }
  __syncthreads();
  const int b0 = s_res[0], above = s_res[1];
  if (tid == 0) SKV_STAMP_NS(bh == 0 ? 16 : -1, 6);
  if (b0 < 1 || b0 > 254) {                     // CTA uniform: the threshold left the window
    if (tid == 0) status[bh] = b0 < 1 ? -4 : -5;
    return;
  }
  const unsigned thr = unsigned(win0 + b0 - 1);
  const int need_eq = S - above;
  // 3. ordered emission: thread t owns the candidates [t * seg, (t + 1) * seg)
  {
    const int seg = (total + kFsThreads - 1) / kFsThreads;
    const int i0 = min(total, tid * seg), i1 = min(total, i0 + seg);
    int gt = 0, eq = 0;
    for (int i = i0; i < i1; ++i) { const unsigned k = s_ckey[i]; gt += k > thr; eq += k == thr; }
    int pg, pe, tg, te;
    block_excl_scan2<kWarps>(gt, eq, s_wa, s_wb, pg, pe, tg, te);
    for (int i = i0; i < i1; ++i) {
      const unsigned k = s_ckey[i];
      if (k > thr || (k == thr && pe < need_eq)) s_new[pg + min(pe, need_eq)] = s_cval[i];
      pg += k > thr; pe += k == thr;
    }
  }
  __syncthreads();
  if (tid == 0) SKV_STAMP_NS(bh == 0 ? 16 : -1, 7);
  int new_id = 0;
  if (tid < S) {
    new_id = s_new[tid];
    if (selected_out) selected_out[size_t(bh) * S + tid] = static_cast<int64_t>(new_id);
  }
  ChunkTableSmem sm{s_sorted, s_slot, s_warp};
  chunk_table_update<S, kInPlace>(tid, old_id, new_id, sm, resident + size_t(bh) * S, offsets + size_t(bh) * S,
                                  cnts + bh);
  if (tid == 0) { hint[bh] = int(thr); status[bh] = 1; }
  if (tid == 0) SKV_STAMP_NS(bh == 0 ? 16 : -1, 8);
  if (row_flags) {                              // the row's chunk table is written: release it
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(row_flags + bh), "r"(1) : "memory");
    }
  }
}

static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

struct SelectWorkspace {
  skv_bf16 *scores;      // [BH, G, ldn], ldn = N rounded up to 8 so every row is 16-byte aligned
  float *pmax, *psum;
  int32_t *fast_status;  // [BH] select_fast_kernel -> select_kernel: 1 = row done
  int T, ldn;
};

static SelectWorkspace carve(const skv_geom *g, void *ws) {
  SelectWorkspace w;
  const int64_t BH = int64_t(g->batch) * g->kv_heads;
  w.T = (g->n_landmarks + 127) / 128;
  w.ldn = (g->n_landmarks + 7) / 8 * 8;
  unsigned char *p = static_cast<unsigned char *>(ws);
  w.scores = reinterpret_cast<skv_bf16 *>(p);
  p += align_up(BH * g->q_groups * w.ldn * 2, 256);
  w.pmax = reinterpret_cast<float *>(p);
  p += align_up(BH * g->q_groups * w.T * 4, 256);
  w.psum = reinterpret_cast<float *>(p);
  p += align_up(BH * g->q_groups * w.T * 4, 256);
  w.fast_status = reinterpret_cast<int32_t *>(p);
  return w;
}

template <int kRowsPerWarp>
static int launch_scores(const skv_bf16 *q, const skv_bf16 *lm, skv_bf16 *scores, float *pmax, float *psum, int BH,
                         int G, int N, int ldn, int T, float alpha, cudaStream_t st, const skv_append *append = nullptr,
                         int64_t buf_rows = 0) {
  skv_append ap{};
  if (append) ap = *append;
  if (G <= 4) {
    dim3 grid(T, BH, 1);
    score_tiles_kernel<4, kRowsPerWarp><<<grid, kScoreThreads, 0, st>>>(q, lm, scores, pmax, psum, G, N, ldn, T, alpha, ap, buf_rows);
  } else {
    dim3 grid(T, BH, (G + 7) / 8);
    score_tiles_kernel<8, kRowsPerWarp><<<grid, kScoreThreads, 0, st>>>(q, lm, scores, pmax, psum, G, N, ldn, T, alpha, ap, buf_rows);
  }
  return launch_status("score_tiles_kernel");
}

}  // namespace skv

using namespace skv;

SKV_TRACE_READER(score)

extern "C" {

int64_t skv_select_workspace_bytes(const skv_geom *g) {
  if (!g) return 0;
  const int64_t BH = int64_t(g->batch) * g->kv_heads;
  const int64_t T = (g->n_landmarks + 127) / 128;
  const int64_t ldn = (g->n_landmarks + 7) / 8 * 8;
  return align_up(BH * g->q_groups * ldn * 2, 256) + 2 * align_up(BH * g->q_groups * T * 4, 256) + align_up(BH * 4, 256);
}

static int check_select_geom(const skv_geom *g, const char *who) {
  if (g->head_dim != kHeadDim || g->q_groups < 1 || g->q_groups > 16 || !valid_map_size(g->select) ||
      g->n_landmarks < g->select || g->n_landmarks > 524280) {   // a select CTA's slice (N / 8) fits 16-bit counts
    set_error("%s: unsupported geometry (D=%d G=%d S=%d N=%d)", who, g->head_dim, g->q_groups, g->select,
              g->n_landmarks);
    return SKV_ERR_UNSUPPORTED;
  }
  return SKV_OK;
}

int skv_score_landmarks(const skv_geom *g, const skv_bf16 *q, const skv_bf16 *landmarks, void *workspace,
                        const skv_append *append, skv_stream_t stream) {
  if (!g || !q || !landmarks || !workspace) { set_error("score_landmarks: null pointer"); return SKV_ERR_ARGUMENT; }
  if (int rc = check_select_geom(g, "score_landmarks")) return rc;
  if (append && (!append->k_new || !append->v_new || !append->k_buf || !append->v_buf || append->pos < 0 ||
                 append->pos >= g->buf_len)) {
    set_error("score_landmarks: bad fused append (pos %d, buffer %d rows)", append->pos, g->buf_len);
    return SKV_ERR_ARGUMENT;
  }
  SelectWorkspace w = carve(g, workspace);
  // scale = 1/sqrt(128), the constant the reference hard-codes (kv_cache.py:1017)
  if (g->wait_flags && (!g->wait_epoch_dev || g->wait_count < 1 || g->wait_count > 4096)) {
    set_error("score_landmarks: wait_flags needs wait_epoch_dev and 1..4096 flags"); return SKV_ERR_ARGUMENT;
  }
  static const bool simt = [] { const char *e = getenv("SKV_SCORE_IMPL"); return e && e[0] == '1'; }();
  if (!simt)   // product path: TMA ring + tcgen05 (score_mma.cu); SKV_SCORE_IMPL=1 selects the SIMT cross-check
    return launch_scores_tcgen05(q, landmarks, w.scores, w.pmax, w.psum, g->batch * g->kv_heads, g->q_groups,
                                 g->n_landmarks, w.ldn, w.T, 0.08838834764831845f, as_stream(stream), append,
                                 g->buf_len, (g->flags & SKV_GEOM_DYNAMIC_LANDMARKS) != 0, g->wait_flags,
                                 g->wait_epoch_dev, g->wait_count,
                                 (g->flags & SKV_GEOM_WAIT_FLAGS_ONLY) != 0 && !(g->flags & SKV_GEOM_DYNAMIC_LANDMARKS));
  if (g->wait_flags) { set_error("score_landmarks: the SIMT cross-check kernel has no flag wait"); return SKV_ERR_UNSUPPORTED; }
  return launch_scores<16>(q, landmarks, w.scores, w.pmax, w.psum, g->batch * g->kv_heads, g->q_groups,
                           g->n_landmarks, w.ldn, w.T, 0.08838834764831845f, as_stream(stream), append, g->buf_len);
}

int skv_select_from_scores(const skv_geom *g, const int64_t *landmark_idx, int64_t *resident, int32_t *offsets,
                           int32_t *cnts, int64_t *selected_out, skv_bf16 *score_out, void *workspace,
                           skv_stream_t stream) {
  if (!g || !landmark_idx || !resident || !offsets || !cnts || !workspace) {
    set_error("select_from_scores: null pointer"); return SKV_ERR_ARGUMENT;
  }
  if (int rc = check_select_geom(g, "select_from_scores")) return rc;
  cudaStream_t st = as_stream(stream);
  const int BH = g->batch * g->kv_heads, G = g->q_groups, N = g->n_landmarks;
  SelectWorkspace w = carve(g, workspace);
  const int per = ((N + 8 * kSelCluster - 1) / (8 * kSelCluster)) * 8;       // landmarks per CTA, multiple of 8
#define SKV_SEL_LAUNCH2(SV, IP, KT, CAP)                                                                          \
  {                                                                                                              \
    const int smem0 = 2 * (KT / 32) * 1024 + (1 + (filter_on ? G : 0)) * int(align_up(int64_t(per) * 2, 16));      \
    const int smem = smem0 + (pad >= 0 ? pad : (spread ? std::max(0, (kSelSpreadBytes - smem0) / 1024 * 1024) : 0)); \
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void *>(select_kernel<SV, IP, KT, CAP>), smem, "select_kernel")) return rc; \
    cudaError_t e = launch_kernel(select_kernel<SV, IP, KT, CAP>, dim3(kSelCluster, BH), dim3(KT), smem, st, true,  \
                                  w.scores, w.pmax, w.psum, landmark_idx, resident, offsets, cnts, selected_out,   \
                                  score_out, g->row_flags, sel_hint, fast_status, G, N, w.ldn, w.T, per, filter_on); \
    if (e != cudaSuccess) { set_error("select_kernel: %s", cudaGetErrorString(e)); return int(e); }               \
  }
#define SKV_SEL_LAUNCH(SV, IP, KT)                                                                                \
  { if (cap) SKV_SEL_LAUNCH2(SV, IP, KT, true) else SKV_SEL_LAUNCH2(SV, IP, KT, false) }
#define SKV_SEL(SV)                                                                                              \
  case SV: {                                                                                                     \
    constexpr int kBase = SV > 256 ? SV : 256, kWide = SV > 512 ? SV : 512;                                      \
    if (wide2 && kBase != 1024) {                                                                                \
      if (g->table_mode) SKV_SEL_LAUNCH(SV, true, 1024) else SKV_SEL_LAUNCH(SV, false, 1024)                     \
    } else if (wide && kWide != kBase) {                                                                         \
      if (g->table_mode) SKV_SEL_LAUNCH(SV, true, kWide) else SKV_SEL_LAUNCH(SV, false, kWide)                   \
    } else {                                                                                                     \
      if (g->table_mode) SKV_SEL_LAUNCH(SV, true, kBase) else SKV_SEL_LAUNCH(SV, false, kBase)                   \
    }                                                                                                            \
  } break;
  // register cap (4 x 256 threads per SM) only when the grid has more CTAs than the GPU has SMs; SKV_SELECT_CAP=0/1 and
  // SKV_SELECT_PAD_KB (extra dynamic shared memory = fewer CTAs of this kernel per SM) are A/B switches
  const char *he = getenv("SKV_SELECT_HINT");               // A/B switch: 0 ignores skv_geom.select_hint
  int32_t *sel_hint = (he && he[0] == '0') ? nullptr : g->select_hint;
  const char *fle = getenv("SKV_SELECT_FILTER");            // A/B switch: 0 = exact keys for every landmark even with a hint
  const int filter_on = (fle && fle[0] == '0') ? 0 : 1;
  const char *ce = getenv("SKV_SELECT_CAP"), *pe = getenv("SKV_SELECT_PAD_KB");
  const bool cap = (ce && *ce) ? ce[0] == '1' : BH * kSelCluster > kNumSMs;
  // One CTA of this kernel per SM while the grid covers at most half of the SMs (batch 1): the dynamic shared memory is
  // padded so that two of its CTAs do not fit an SM, but one CTA of the fused fetch launch (109 KiB) still fits next to
  // it and can set itself up ahead of its row's hand-over.  Measured at batch 1 (122K): score + select 16.8 -> 14.2 us,
  // step 62.5 -> 61.9 us (V on the host) and 32.4 -> 31.9 us (V in HBM); a larger pad that keeps the fused launch's CTAs
  // out (100 KiB) costs the HBM configuration 1.2 us, and at batch 2 (128 CTAs) the padding is a loss (9.2 -> 14.5 us).
  const bool spread = BH * kSelCluster <= kNumSMs / 2;
  const int pad = (pe && *pe) ? atoi(pe) * 1024 : -1;
  const char *wf = getenv("SKV_SELECT_WIDE_FROM");          // read per call: tests and A/B runs flip it
  const bool wide = per >= ((wf && *wf) ? atoi(wf) : 3072);  // landmarks per CTA from which 512 threads pay
  const char *wf2 = getenv("SKV_SELECT_WIDE2_FROM");
  // ... and 1024 threads.  With the candidate filter the key pass is cheap and the 1024-thread barriers are not: at 512K
  // (8192 landmarks per CTA) 512 threads take 12.4 us against 13.4 us, so 1024 threads are left to 1M-token slices
  const bool wide2 = per >= ((wf2 && *wf2) ? atoi(wf2) : 12288);
  // Single-CTA hinted selection in front of the cluster kernel (select_fast_kernel): rows whose threshold stays inside the
  // hint window are settled there, the cluster kernel only does the others.  A measured experiment, OFF by default
  // (SKV_SELECT_FAST=1 enables it): bit-identical results, but one SM is too slow for a row -- 14.4 us behind its wait at
  // 122K (%globaltimer stamps: 124 KB of scores into one SM 4.7, filter + scan 2.6, compaction 2.2, exact keys of 1 600
  // candidates 1.7, emission 1.4, table 1.5; 1024-thread barriers cost 0.2-0.3 us each) against 7.5 us for the 8-CTA
  // cluster kernel, and its 1024 x 64 registers keep every other CTA off its SM: 67.7 against 61.1 us per layer-step
  const char *fse = getenv("SKV_SELECT_FAST");
  const int32_t *fast_status = nullptr;
  if ((fse && fse[0] == '1') && sel_hint && !score_out && N <= kFsThreads * kFsPer && N >= g->select) {
#define SKV_FAST(SV)                                                                                              \
  case SV: {                                                                                                      \
    cudaError_t e = g->table_mode                                                                                 \
        ? launch_kernel(select_fast_kernel<SV, true>, dim3(BH), dim3(kFsThreads), 0, st, true, w.scores, w.pmax,    \
                        w.psum, landmark_idx, resident, offsets, cnts, selected_out, g->row_flags, sel_hint,        \
                        w.fast_status, G, N, w.ldn, w.T)                                                            \
        : launch_kernel(select_fast_kernel<SV, false>, dim3(BH), dim3(kFsThreads), 0, st, true, w.scores, w.pmax,   \
                        w.psum, landmark_idx, resident, offsets, cnts, selected_out, g->row_flags, sel_hint,        \
                        w.fast_status, G, N, w.ldn, w.T);                                                           \
    if (e != cudaSuccess) { set_error("select_fast_kernel: %s", cudaGetErrorString(e)); return int(e); }          \
  } break;
    switch (g->select) {
      SKV_FAST(128) SKV_FAST(256) SKV_FAST(512) SKV_FAST(1024)
      default: return SKV_ERR_UNSUPPORTED;
    }
#undef SKV_FAST
    fast_status = w.fast_status;
  }
  switch (g->select) {
    SKV_SEL(128) SKV_SEL(256) SKV_SEL(512) SKV_SEL(1024)
    default: return SKV_ERR_UNSUPPORTED;
  }
#undef SKV_SEL_LAUNCH
#undef SKV_SEL_LAUNCH2
#undef SKV_SEL
  return launch_status("select_kernel");
}

int skv_select_chunks(const skv_geom *g, const skv_bf16 *q, const skv_bf16 *landmarks, const int64_t *landmark_idx,
                      int64_t *resident, int32_t *offsets, int32_t *cnts, int64_t *selected_out,
                      skv_bf16 *score_out, void *workspace, const skv_append *append, skv_stream_t stream) {
  if (!g || !q || !landmarks || !landmark_idx || !resident || !offsets || !cnts || !workspace) {
    set_error("select_chunks: null pointer"); return SKV_ERR_ARGUMENT;
  }
  if (int rc = check_select_geom(g, "select_chunks")) return rc;
  if (append && (!append->k_new || !append->v_new || !append->k_buf || !append->v_buf || append->pos < 0 ||
                 append->pos >= g->buf_len)) {
    set_error("select_chunks: bad fused append (pos %d, buffer %d rows)", append->pos, g->buf_len);
    return SKV_ERR_ARGUMENT;
  }
  // experiment (SKV_FUSED_SELECT=1): scoring + selection as ONE cluster launch (score_select.cu).  Measured slower than
  // the two launches below at every batch (a row's cluster cannot span enough SMs to keep HBM busy),
  // so it is NOT the product path; shapes it does not take fall through
  const char *fe = getenv("SKV_FUSED_SELECT"), *se = getenv("SKV_SCORE_IMPL");      // read per call: tests flip it
  const bool fused = (fe && fe[0] == '1') && !(se && se[0] == '1') && !g->row_flags && !g->wait_flags;   // the experiment has neither hand-over
  if (fused) {
    const int rc = launch_score_select(g, q, landmarks, landmark_idx, resident, offsets, cnts, selected_out, score_out,
                                       append, as_stream(stream));
    if (rc != SKV_ERR_UNSUPPORTED) return rc;
  }
  if (int rc = skv_score_landmarks(g, q, landmarks, workspace, append, stream)) return rc;
  return skv_select_from_scores(g, landmark_idx, resident, offsets, cnts, selected_out, score_out, workspace, stream);
}

int skv_batch_gemm_softmax(const skv_bf16 *A, const skv_bf16 *B, skv_bf16 *D, float *Norm, float *Sum,
                           skv_bf16 *Softmax, int batch_count, int m, int n, int k, float alpha, float beta,
                           skv_stream_t stream) {
  if (!A || !B || !D || !Norm || !Sum || !Softmax) { set_error("batch_gemm_softmax: null pointer"); return SKV_ERR_ARGUMENT; }
  if (k != kHeadDim || m < 1 || m > 16 || beta != 0.f || n < 1) {
    set_error("batch_gemm_softmax: need k=128, 1<=m<=16, beta=0 (got k=%d m=%d beta=%g)", k, m, (double)beta);
    return SKV_ERR_UNSUPPORTED;
  }
  cudaStream_t st = as_stream(stream);
  const int T = (n + 255) / 256;     // the reference's partial layout [batch, m, ceil(n/256)] (kv_cache.py:778-780)
  int rc = launch_scores<32>(A, B, D, Norm, Sum, batch_count, m, n, n, T, alpha, st);
  if (rc) return rc;
  dim3 grid((n + 2047) / 2048, batch_count);
  softmax_apply_kernel<<<grid, 256, 0, st>>>(D, Norm, Sum, Softmax, m, n, T);
  return launch_status("softmax_apply_kernel");
}

}  // extern "C"
