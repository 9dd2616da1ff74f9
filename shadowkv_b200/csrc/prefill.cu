// Prefill-side state build of ShadowKVCache_CPU.prefill_kv_cache (reference models/kv_cache.py:788-980; SURVEY.md
// section 8 f1): the pieces that stream the prompt's post-RoPE keys once and decide which chunks become exact
// "outlier" chunks and which become landmarks.  HBM bound (read K once: H*L*D*2 bytes per sequence-layer).
//
//  chunk_stats_kernel   : per chunk of 8 keys, the bf16 mean (landmark candidate, :856) and the minimum over the
//                         8 keys of cosine_similarity(mean, key) (:859-866) with the bf16 rounding points of the
//                         torch ops the reference calls on bf16 tensors (torch 2.3 ATen/native/Distance.cpp:
//                         norms via linalg_vector_norm -> clamp_min(eps) -> x/|x| elementwise -> product -> sum).
//  outlier_select_kernel: per (b,h), the `n_outliers` chunks with the SMALLEST min-cosine (:868; ties: lowest chunk
//                         index, torch.topk leaves them unspecified), emitted in ascending chunk order, plus the
//                         ascending list of all other chunks = the landmark index (:909-915).
//  gather_rows_kernel   : landmarks = means[rest_idx] (:917-921).
// The outlier / first-selection K and V chunks are moved with the chunk gather of gather.cu (skv_gather_chunks).
#include "common.cuh"

namespace skv {

__device__ __forceinline__ float bf16r(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

__global__ void __launch_bounds__(256)
    chunk_stats_kernel(const skv_bf16 *__restrict__ k, int64_t sb, int64_t sh, int64_t sr, int H, int chunks,
                       skv_bf16 *__restrict__ means, skv_bf16 *__restrict__ mincos) {
  const int bh = blockIdx.y, b = bh / H, h = bh % H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, sub = lane & 15, hw = lane >> 4;
  const int chunk = blockIdx.x * 8 + warp;
  if (chunk >= chunks) return;
  const skv_bf16 *base = k + size_t(b) * sb + size_t(h) * sh + size_t(chunk) * kChunkTokens * sr + sub * 8;
  float x[4][8];                                   // this lane's 8 dims of rows 2j + hw
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int4 w = ld_stream(reinterpret_cast<const int4 *>(base + size_t(2 * j + hw) * sr));
    const uint32_t u[4] = {(uint32_t)w.x, (uint32_t)w.y, (uint32_t)w.z, (uint32_t)w.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) { x[j][2 * e] = bf16_lo(u[e]); x[j][2 * e + 1] = bf16_hi(u[e]); }
  }
  // mean over the 8 rows: fp32 sum, / 8, round to bf16 (torch.mean on a bf16 tensor)
  float m[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    float s = (x[0][e] + x[1][e]) + (x[2][e] + x[3][e]);
    s += __shfl_xor_sync(0xffffffffu, s, 16);
    m[e] = bf16r(s * 0.125f);
  }
  if (hw == 0) {
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) o[e] = uint32_t(f32_to_bf16_bits(m[2 * e])) | (uint32_t(f32_to_bf16_bits(m[2 * e + 1])) << 16);
    *reinterpret_cast<int4 *>(means + (size_t(bh) * chunks + chunk) * kHeadDim + sub * 8) = make_int4(o[0], o[1], o[2], o[3]);
  }
  // |mean| : fp32 accumulate, sqrt, round to bf16, clamp_min(eps)
  float n1 = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) n1 = fmaf(m[e], m[e], n1);
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) n1 += __shfl_xor_sync(0xffffffffu, n1, o);
  n1 = fmaxf(bf16r(sqrtf(n1)), 1e-8f);
  float a[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) a[e] = bf16r(m[e] / n1);
  float cmin = INFINITY;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float n2 = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) n2 = fmaf(x[j][e], x[j][e], n2);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, o);      // the 16 lanes of one row
    n2 = fmaxf(bf16r(sqrtf(n2)), 1e-8f);
    float c = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) c += bf16r(a[e] * bf16r(x[j][e] / n2));
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    cmin = fminf(cmin, bf16r(c));
  }
  cmin = fminf(cmin, __shfl_xor_sync(0xffffffffu, cmin, 16));
  if (lane == 0) mincos[size_t(bh) * chunks + chunk] = f32_to_bf16_bits(cmin);
}

// bf16 bits -> 16-bit key that ascends with the value
__device__ __forceinline__ unsigned ascending_key(uint16_t bits) {
  return (bits & 0x8000u) ? (~unsigned(bits) & 0xFFFFu) : (unsigned(bits) | 0x8000u);
}

constexpr int kOutThreads = 1024;

__global__ void __launch_bounds__(kOutThreads)
    outlier_select_kernel(const skv_bf16 *__restrict__ mincos, int chunks, int n_out, int64_t *__restrict__ outliers,
                          int64_t *__restrict__ rest_idx) {
  extern __shared__ __align__(16) unsigned char osm[];
  uint16_t *keys = reinterpret_cast<uint16_t *>(osm);                   // [chunks]
  __shared__ int s_hist[256];
  __shared__ int s_res[2];
  __shared__ int s_wa[33], s_wb[33];
  const int row = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const skv_bf16 *src = mincos + size_t(row) * chunks;
  // two-level radix select of the n_out SMALLEST keys
  int thr_hi = 0, below_hi = 0, thr = 0, need_eq = 0;
  for (int level = 0; level < 2; ++level) {
    if (tid < 256) s_hist[tid] = 0;
    __syncthreads();
    for (int i0 = 0; i0 < chunks; i0 += kOutThreads) {
      const int i = i0 + tid;
      unsigned bin = 0xFFFFu;
      if (i < chunks) {
        unsigned key;
        if (level == 0) { key = ascending_key(src[i]); keys[i] = uint16_t(key); bin = key >> 8; }
        else { key = keys[i]; if (int(key >> 8) == thr_hi) bin = key & 0xFFu; }
      }
      const unsigned peers = __match_any_sync(0xffffffffu, bin);
      if (bin != 0xFFFFu && lane == (__ffs(peers) - 1)) atomicAdd(&s_hist[bin], __popc(peers));
    }
    __syncthreads();
    if (tid == 0) {                                                  // 256 bins: a serial scan is fine here
      const int want = level == 0 ? n_out : n_out - below_hi;
      int acc = 0, bsel = 255;
      for (int bidx = 0; bidx < 256; ++bidx) {
        if (acc + s_hist[bidx] >= want) { bsel = bidx; break; }
        acc += s_hist[bidx];
      }
      s_res[0] = bsel; s_res[1] = acc;
    }
    __syncthreads();
    if (level == 0) { thr_hi = s_res[0]; below_hi = s_res[1]; }
    else { thr = (thr_hi << 8) | s_res[0]; need_eq = n_out - below_hi - s_res[1]; }
    __syncthreads();
  }
  // ordered emission: warp w owns a contiguous range; position of a chunk = (#selected before it) in the outlier
  // list, (#unselected before it) in the landmark index
  const int per_warp = (((chunks + 31) / 32 + 31) / 32) * 32;
  const int w_begin = min(chunks, warp * per_warp), w_end = min(chunks, w_begin + per_warp);
  int lt = 0, eq = 0;
  for (int i0 = w_begin; i0 < w_end; i0 += 32) {
    const int i = i0 + lane;
    const bool valid = i < w_end;
    const int key = valid ? keys[i] : 0xFFFFFF;
    lt += __popc(__ballot_sync(0xffffffffu, valid && key < thr));
    eq += __popc(__ballot_sync(0xffffffffu, valid && key == thr));
  }
  if (lane == 0) { s_wa[warp] = lt; s_wb[warp] = eq; }
  __syncthreads();
  if (warp == 0) {
    const int a = s_wa[lane], e = s_wb[lane];
    const int ia = warp_incl_scan(a), ie = warp_incl_scan(e);
    __syncwarp();
    s_wa[lane] = ia - a; s_wb[lane] = ie - e;
  }
  __syncthreads();
  int lt_seen = s_wa[warp], eq_seen = s_wb[warp];
  int64_t *out_o = outliers + size_t(row) * n_out;
  int64_t *out_r = rest_idx + size_t(row) * (chunks - n_out);
  for (int i0 = w_begin; i0 < w_end; i0 += 32) {
    const int i = i0 + lane;
    const bool valid = i < w_end;
    const int key = valid ? keys[i] : 0xFFFFFF;
    const unsigned ltm = __ballot_sync(0xffffffffu, valid && key < thr);
    const unsigned eqm = __ballot_sync(0xffffffffu, valid && key == thr);
    const unsigned below = (1u << lane) - 1u;
    const int my_lt = lt_seen + __popc(ltm & below), my_eq = eq_seen + __popc(eqm & below);
    if (valid) {
      const bool take = key < thr || (key == thr && my_eq < need_eq);
      const int sel_before = my_lt + min(my_eq, need_eq);
      if (take) out_o[sel_before] = i;
      else out_r[i - sel_before] = i;
    }
    lt_seen += __popc(ltm); eq_seen += __popc(eqm);
  }
}

__global__ void __launch_bounds__(256)
    gather_rows_kernel(const skv_bf16 *__restrict__ src, const int64_t *__restrict__ idx, skv_bf16 *__restrict__ dst, int n_src,
                       int n) {
  const int row = blockIdx.y;
  const int i = blockIdx.x * 16 + (threadIdx.x >> 4), c = threadIdx.x & 15;     // 16 lanes move one 256-byte row
  if (i >= n) return;
  const int64_t s = idx[size_t(row) * n + i];
  const int4 v = ld_stream(reinterpret_cast<const int4 *>(src + (size_t(row) * n_src + s) * kHeadDim) + c);
  reinterpret_cast<int4 *>(dst + (size_t(row) * n + i) * kHeadDim)[c] = v;
}

}  // namespace skv

using namespace skv;

extern "C" {

int skv_chunk_stats(const skv_bf16 *k, int batch, int kv_heads, int chunks, int64_t stride_b, int64_t stride_h,
                    int64_t stride_row, skv_bf16 *means, skv_bf16 *mincos, skv_stream_t stream) {
  if (!k || !means || !mincos) { set_error("chunk_stats: null pointer"); return SKV_ERR_ARGUMENT; }
  if (chunks < 1 || batch < 1 || kv_heads < 1 || (stride_row % 8) || (stride_h % 8) || (stride_b % 8)) {
    set_error("chunk_stats: need 16-byte aligned strides and at least one chunk"); return SKV_ERR_UNSUPPORTED;
  }
  dim3 grid((chunks + 7) / 8, batch * kv_heads);
  chunk_stats_kernel<<<grid, 256, 0, as_stream(stream)>>>(k, stride_b, stride_h, stride_row, kv_heads, chunks, means, mincos);
  return launch_status("chunk_stats_kernel");
}

int skv_select_outliers(const skv_bf16 *mincos, int rows, int chunks, int n_outliers, int64_t *outliers, int64_t *rest_idx,
                        skv_stream_t stream) {
  if (!mincos || !outliers || !rest_idx) { set_error("select_outliers: null pointer"); return SKV_ERR_ARGUMENT; }
  if (rows < 1 || n_outliers < 1 || n_outliers >= chunks || chunks > 100000) {
    set_error("select_outliers: need 1 <= n_outliers < chunks <= 100000 (got %d of %d)", n_outliers, chunks);
    return SKV_ERR_UNSUPPORTED;
  }
  const int smem = (chunks * 2 + 15) / 16 * 16;
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void *>(outlier_select_kernel), smem, "outlier_select_kernel")) return rc;
  outlier_select_kernel<<<rows, kOutThreads, smem, as_stream(stream)>>>(mincos, chunks, n_outliers, outliers, rest_idx);
  return launch_status("outlier_select_kernel");
}

int skv_gather_rows(const skv_bf16 *src, const int64_t *idx, skv_bf16 *dst, int rows, int n_src, int n, skv_stream_t stream) {
  if (!src || !idx || !dst) { set_error("gather_rows: null pointer"); return SKV_ERR_ARGUMENT; }
  if (rows < 1 || n < 1) return SKV_OK;
  gather_rows_kernel<<<dim3((n + 15) / 16, rows), 256, 0, as_stream(stream)>>>(src, idx, dst, n_src, n);
  return launch_status("gather_rows_kernel");
}

}  // extern "C"
