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
#include "chunk_table.cuh"
#include "common.cuh"

namespace skv {

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
                       int G, int N, int T, float alpha) {
  constexpr int kTileRows = kRowsPerWarp * (kScoreThreads / 32);
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
        scores[(size_t(bh) * G + g0 + g) * N + row] = bits;
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
// select_kernel
// ------------------------------------------------------------------------------------------------
constexpr int kSelThreads = 1024;

// histogram of `bin` (0..255; 0xFFFF = no element) with one atomic per distinct value per warp
__device__ __forceinline__ void hist_add(int *hist, unsigned bin) {
  const unsigned peers = __match_any_sync(0xffffffffu, bin);
  if (bin != 0xFFFFu && (threadIdx.x & 31) == (__ffs(peers) - 1)) atomicAdd(&hist[bin], __popc(peers));
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

template <int S>
__global__ void __launch_bounds__(kSelThreads)
    select_kernel(const skv_bf16 *__restrict__ scores,        // [BH, G, N] bf16 scaled dot products
                  const float *__restrict__ part_max, const float *__restrict__ part_sum,   // [BH, G, T]
                  const int64_t *__restrict__ landmark_idx,   // [BH, N]
                  int64_t *__restrict__ resident,             // [BH, S] in/out
                  int32_t *__restrict__ offsets, int32_t *__restrict__ cnts,
                  int64_t *__restrict__ selected_out,         // [BH, S] or null
                  skv_bf16 *__restrict__ score_out,           // [BH, N] or null
                  int G, int N, int T) {
  extern __shared__ __align__(16) unsigned char sel_smem[];
  uint16_t *keys = reinterpret_cast<uint16_t *>(sel_smem);                   // [N rounded to 8]
  __shared__ float s_m[16], s_inv[16];
  __shared__ int s_hist[256];
  __shared__ int s_res[4];
  __shared__ int s_wgt[32], s_weq[32];
  __shared__ int s_sel[S];
  __shared__ int s_sorted[S], s_slot[S], s_warp[33];

  const int bh = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // 1. final softmax statistics per head
  for (int g = warp; g < G; g += kSelThreads / 32) {
    float m, inv;
    reduce_partials(part_max + (size_t(bh) * G + g) * T, part_sum + (size_t(bh) * G + g) * T, T, lane, m, inv);
    if (lane == 0) { s_m[g] = m; s_inv[g] = inv; }
  }
  if (tid < 256) s_hist[tid] = 0;
  __syncthreads();

  // 2. p = bf16(exp(s - max) * inv); key = max over heads (bf16 bits are monotone for p >= 0);
  //    first-level histogram on the high byte
  const skv_bf16 *sc = scores + size_t(bh) * G * N;
  const int n_iter = (N + kSelThreads - 1) / kSelThreads;
  for (int i = 0; i < n_iter; ++i) {
    const int n = i * kSelThreads + tid;
    unsigned bin = 0xFFFFu;
    if (n < N) {
      uint16_t best = 0;
      for (int g = 0; g < G; ++g) {
        const float p = expf(bf16_bits_to_f32(sc[size_t(g) * N + n]) - s_m[g]) * s_inv[g];
        const uint16_t b = f32_to_bf16_bits(p);
        best = b > best ? b : best;
      }
      keys[n] = best;
      if (score_out) score_out[size_t(bh) * N + n] = best;
      bin = best >> 8;
    }
    hist_add(s_hist, bin);
  }
  __syncthreads();
  if (warp == 0) find_threshold_bin(s_hist, S, s_res);
  __syncthreads();
  const int b1 = s_res[0], above1 = s_res[1];
  if (tid < 256) s_hist[tid] = 0;
  __syncthreads();
  for (int i = 0; i < n_iter; ++i) {
    const int n = i * kSelThreads + tid;
    unsigned bin = 0xFFFFu;
    if (n < N) {
      const uint16_t k = keys[n];
      if ((k >> 8) == b1) bin = k & 0xFF;
    }
    hist_add(s_hist, bin);
  }
  __syncthreads();
  if (warp == 0) find_threshold_bin(s_hist, S - above1, s_res + 2);
  __syncthreads();
  const unsigned thr = (unsigned(b1) << 8) | unsigned(s_res[2]);
  const int need_eq = S - (above1 + s_res[3]);      // how many keys equal to thr are taken (>= 1)

  // 3. ordered emission (ascending landmark index).  Warp w owns a contiguous range.
  const int per_warp = (((N + 31) / 32 + 31) / 32) * 32;      // multiple of 32
  const int w_begin = warp * per_warp, w_end = min(N, w_begin + per_warp);
  int wgt = 0, weq = 0;
  for (int n0 = w_begin; n0 < w_end; n0 += 32) {
    const int n = n0 + lane;
    const unsigned k = n < w_end ? keys[n] : 0u;
    const bool valid = n < w_end;
    wgt += __popc(__ballot_sync(0xffffffffu, valid && k > thr));
    weq += __popc(__ballot_sync(0xffffffffu, valid && k == thr));
  }
  if (lane == 0) { s_wgt[warp] = wgt; s_weq[warp] = weq; }
  __syncthreads();
  if (warp == 0) {
    const int g_incl = warp_incl_scan(s_wgt[lane]);
    const int e_incl = warp_incl_scan(s_weq[lane]);
    __syncwarp();
    s_wgt[lane] = g_incl - s_wgt[lane];
    s_weq[lane] = e_incl - s_weq[lane];
  }
  __syncthreads();
  {
    int eq_seen = s_weq[warp];
    int pos = s_wgt[warp] + min(eq_seen, need_eq);
    for (int n0 = w_begin; n0 < w_end; n0 += 32) {
      const int n = n0 + lane;
      const bool valid = n < w_end;
      const unsigned k = valid ? keys[n] : 0u;
      const unsigned eqm = __ballot_sync(0xffffffffu, valid && k == thr);
      const int my_eq_rank = eq_seen + __popc(eqm & ((1u << lane) - 1u));
      const bool take = valid && (k > thr || (k == thr && my_eq_rank < need_eq));
      const unsigned tm = __ballot_sync(0xffffffffu, take);
      if (take) s_sel[pos + __popc(tm & ((1u << lane) - 1u))] = n;
      pos += __popc(tm);
      eq_seen += __popc(eqm);
    }
  }
  __syncthreads();

  // 4. chunk ids of the selection, then the chunk-table update
  int new_id = 0, old_id = 0;
  if (tid < S) {
    const int64_t cid = landmark_idx[size_t(bh) * N + s_sel[tid]];
    if (selected_out) selected_out[size_t(bh) * S + tid] = cid;
    new_id = static_cast<int>(cid);
    old_id = static_cast<int>(resident[size_t(bh) * S + tid]);
  }
  ChunkTableSmem sm{s_sorted, s_slot, s_warp};
  chunk_table_update<S>(tid, old_id, new_id, sm, resident + size_t(bh) * S, offsets + size_t(bh) * S, cnts + bh);
}

static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

struct SelectWorkspace {
  skv_bf16 *scores;
  float *pmax, *psum;
  int T;
};

static SelectWorkspace carve(const skv_geom *g, void *ws) {
  SelectWorkspace w;
  const int64_t BH = int64_t(g->batch) * g->kv_heads;
  w.T = (g->n_landmarks + 127) / 128;
  unsigned char *p = static_cast<unsigned char *>(ws);
  w.scores = reinterpret_cast<skv_bf16 *>(p);
  p += align_up(BH * g->q_groups * g->n_landmarks * 2, 256);
  w.pmax = reinterpret_cast<float *>(p);
  p += align_up(BH * g->q_groups * w.T * 4, 256);
  w.psum = reinterpret_cast<float *>(p);
  return w;
}

template <int kRowsPerWarp>
static int launch_scores(const skv_bf16 *q, const skv_bf16 *lm, skv_bf16 *scores, float *pmax, float *psum, int BH,
                         int G, int N, int T, float alpha, cudaStream_t st) {
  if (G <= 4) {
    dim3 grid(T, BH, 1);
    score_tiles_kernel<4, kRowsPerWarp><<<grid, kScoreThreads, 0, st>>>(q, lm, scores, pmax, psum, G, N, T, alpha);
  } else {
    dim3 grid(T, BH, (G + 7) / 8);
    score_tiles_kernel<8, kRowsPerWarp><<<grid, kScoreThreads, 0, st>>>(q, lm, scores, pmax, psum, G, N, T, alpha);
  }
  return launch_status("score_tiles_kernel");
}

}  // namespace skv

using namespace skv;

extern "C" {

int64_t skv_select_workspace_bytes(const skv_geom *g) {
  if (!g) return 0;
  const int64_t BH = int64_t(g->batch) * g->kv_heads;
  const int64_t T = (g->n_landmarks + 127) / 128;
  return align_up(BH * g->q_groups * g->n_landmarks * 2, 256) + 2 * align_up(BH * g->q_groups * T * 4, 256);
}

static int check_select_geom(const skv_geom *g, const char *who) {
  if (g->head_dim != kHeadDim || g->q_groups < 1 || g->q_groups > 16 || !valid_map_size(g->select) ||
      g->n_landmarks < g->select || g->n_landmarks > 100000) {
    set_error("%s: unsupported geometry (D=%d G=%d S=%d N=%d)", who, g->head_dim, g->q_groups, g->select,
              g->n_landmarks);
    return SKV_ERR_UNSUPPORTED;
  }
  return SKV_OK;
}

int skv_score_landmarks(const skv_geom *g, const skv_bf16 *q, const skv_bf16 *landmarks, void *workspace,
                        skv_stream_t stream) {
  if (!g || !q || !landmarks || !workspace) { set_error("score_landmarks: null pointer"); return SKV_ERR_ARGUMENT; }
  if (int rc = check_select_geom(g, "score_landmarks")) return rc;
  SelectWorkspace w = carve(g, workspace);
  // scale = 1/sqrt(128), the constant the reference hard-codes (kv_cache.py:1017)
  return launch_scores<16>(q, landmarks, w.scores, w.pmax, w.psum, g->batch * g->kv_heads, g->q_groups,
                           g->n_landmarks, w.T, 0.08838834764831845f, as_stream(stream));
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
  const int smem = int(align_up(int64_t(N) * 2, 16));
#define SKV_SEL(SV)                                                                                              \
  case SV: {                                                                                                     \
    static int cur = 0;                                                                                          \
    if (smem > cur) { cudaFuncSetAttribute(select_kernel<SV>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); cur = smem; } \
    select_kernel<SV><<<BH, kSelThreads, smem, st>>>(w.scores, w.pmax, w.psum, landmark_idx, resident, offsets,  \
                                                     cnts, selected_out, score_out, G, N, w.T);                  \
  } break;
  switch (g->select) {
    SKV_SEL(128) SKV_SEL(256) SKV_SEL(512) SKV_SEL(1024)
    default: return SKV_ERR_UNSUPPORTED;
  }
#undef SKV_SEL
  return launch_status("select_kernel");
}

int skv_select_chunks(const skv_geom *g, const skv_bf16 *q, const skv_bf16 *landmarks, const int64_t *landmark_idx,
                      int64_t *resident, int32_t *offsets, int32_t *cnts, int64_t *selected_out,
                      skv_bf16 *score_out, void *workspace, skv_stream_t stream) {
  if (int rc = skv_score_landmarks(g, q, landmarks, workspace, stream)) return rc;
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
  int rc = launch_scores<32>(A, B, D, Norm, Sum, batch_count, m, n, T, alpha, st);
  if (rc) return rc;
  dim3 grid((n + 2047) / 2048, batch_count);
  softmax_apply_kernel<<<grid, 256, 0, st>>>(D, Norm, Sum, Softmax, m, n, T);
  return launch_status("softmax_apply_kernel");
}

}  // extern "C"
