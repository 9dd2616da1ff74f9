// Split-K flash-decoding attention of ONE query token over the assembled K/V buffer
// [local | outliers | sparse | generated]; replaces flash_attn_with_kvcache at reference
// models/base.py:341.  HBM bound: every K and V row is read once (2 * H * L' * D * 2 bytes per sequence);
// 4 query heads share each loaded row.  One launch: the last CTA to arrive per (b, head group) merges the
// split partials (arrival counter in the workspace, left at zero for the next call).
#include "common.cuh"

namespace skv {

// attention_mma.cu: TMA ring + tensor-core attention (product path)
int launch_attention_mma(const skv_bf16 *q, const skv_bf16 *k, const skv_bf16 *v, skv_bf16 *out, int batch, int kv_heads,
                         int q_groups, int kv_len, int64_t stride_b, int64_t stride_h, int64_t stride_row, float scale,
                         int num_splits, void *workspace, const int32_t *kv_len_dev, bool pdl, cudaStream_t st);

constexpr int kAttnMinWarps = 8;              // 256 threads: the combine maps 128 dims x 2 head pairs onto them
constexpr int kAttnMaxWarps = 12;             // block size is chosen per launch (see plan_splits)
constexpr int kAttnGT = 4;                   // query heads per CTA
constexpr int kAttnGroupKeys = 16;           // keys per warp iteration
constexpr int kAttnMaxSplits = 64;

// barrier over the first 256 threads only (the extra key-streaming warps have left by then)
__device__ __forceinline__ void bar256() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

struct AttnParams {
  const skv_bf16 *q;     // [B, H*G, 128]
  const skv_bf16 *k, *v;
  skv_bf16 *out;         // [B, H*G, 128]
  float *part_o;         // [rows, max_splits, GT, 128]
  float *part_ml;        // [rows, max_splits, GT, 2]
  int *counters;         // [rows]
  const int32_t *kv_len_dev;   // optional device-side length
  int B, H, G, ngz, kv_len, keys_per_split, num_splits;
  int64_t sb, sh, sr;
  float scale_log2;      // scale * log2(e)
};

__global__ void __launch_bounds__(kAttnMaxWarps * 32, 1)
    attention_decode_kernel(AttnParams p) {
  __shared__ __align__(16) float s_p[kAttnMaxWarps][kAttnGroupKeys][kAttnGT];
  __shared__ float s_o[kAttnMaxWarps][kAttnGT][kHeadDim];
  __shared__ float s_ml[kAttnMaxWarps][kAttnGT][2];
  __shared__ float s_f[kAttnMaxSplits][kAttnGT];
  __shared__ float s_L[kAttnGT];
  __shared__ int s_last;

  pdl_launch_dependents();      // the next layer's scoring kernel may start streaming its landmarks behind us
  const int split = blockIdx.x;
  const int h = blockIdx.y / p.ngz, gz = blockIdx.y % p.ngz, b = blockIdx.z;
  const int g0 = gz * kAttnGT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, sub = lane & 15, hw = lane >> 4;
  const int nwarps = blockDim.x >> 5;
  const int row_id = (b * p.H + h) * p.ngz + gz;
  SKV_STAMP(row_id == 0 ? split : 999, 0);

  // q stays packed (4 registers per head); it is unpacked after the K/V loads of an iteration are in flight,
  // so the first K/V round trip does not wait behind the q round trip
  int4 qraw[kAttnGT];
#pragma unroll
  for (int g = 0; g < kAttnGT; ++g)
    qraw[g] = (g0 + g < p.G)
                  ? *reinterpret_cast<const int4 *>(p.q + (size_t(b) * p.H * p.G + size_t(h) * p.G + g0 + g) * kHeadDim + sub * 8)
                  : make_int4(0, 0, 0, 0);

  const int kv_len = p.kv_len_dev ? min(p.kv_len, *p.kv_len_dev) : p.kv_len;
  SKV_STAMP(row_id == 0 ? split : 999, 1);
  const int k_begin = split * p.keys_per_split;
  const int k_end = min(kv_len, k_begin + p.keys_per_split);
  const int4 *kbase = reinterpret_cast<const int4 *>(p.k + size_t(b) * p.sb + size_t(h) * p.sh);
  const int4 *vbase = reinterpret_cast<const int4 *>(p.v + size_t(b) * p.sb + size_t(h) * p.sh);
  const size_t row_i4 = size_t(p.sr) / 8;          // int4 per row stride

  float m_run[kAttnGT], l_run[kAttnGT], acc[kAttnGT][8];
#pragma unroll
  for (int g = 0; g < kAttnGT; ++g) {
    m_run[g] = -INFINITY; l_run[g] = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[g][e] = 0.f;
  }

#pragma unroll 1
  for (int kg = k_begin + warp * kAttnGroupKeys; kg < k_end; kg += nwarps * kAttnGroupKeys) {
    int4 kl[8], vl[8];
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int key = kg + it * 2 + hw;
      const bool ok = key < k_end;
      kl[it] = ok ? ld_stream(kbase + size_t(key) * row_i4 + sub) : make_int4(0, 0, 0, 0);
      vl[it] = ok ? ld_stream(vbase + size_t(key) * row_i4 + sub) : make_int4(0, 0, 0, 0);
    }
    float qf[kAttnGT][8];
#pragma unroll
    for (int g = 0; g < kAttnGT; ++g) {
      const uint32_t u[4] = {(uint32_t)qraw[g].x, (uint32_t)qraw[g].y, (uint32_t)qraw[g].z, (uint32_t)qraw[g].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) { qf[g][2 * e] = bf16_lo(u[e]) * p.scale_log2; qf[g][2 * e + 1] = bf16_hi(u[e]) * p.scale_log2; }
    }
    float v[8][kAttnGT];
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const uint32_t u[4] = {(uint32_t)kl[it].x, (uint32_t)kl[it].y, (uint32_t)kl[it].z, (uint32_t)kl[it].w};
      float x[8];
#pragma unroll
      for (int e = 0; e < 4; ++e) { x[2 * e] = bf16_lo(u[e]); x[2 * e + 1] = bf16_hi(u[e]); }
#pragma unroll
      for (int g = 0; g < kAttnGT; ++g) {
        float a = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) a = fmaf(x[e], qf[g][e], a);
        v[it][g] = a;
      }
    }
    // transposing butterfly (see score.cu): lane ends with key `it`, heads gsel, gsel+1
    float a4[4][kAttnGT];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int g = 0; g < kAttnGT; ++g) {
        const float keep = (sub & 8) ? v[i + 4][g] : v[i][g];
        const float send = (sub & 8) ? v[i][g] : v[i + 4][g];
        a4[i][g] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
      }
    float a2[2][kAttnGT];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int g = 0; g < kAttnGT; ++g) {
        const float keep = (sub & 4) ? a4[i + 2][g] : a4[i][g];
        const float send = (sub & 4) ? a4[i][g] : a4[i + 2][g];
        a2[i][g] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
      }
    float a1[kAttnGT];
#pragma unroll
    for (int g = 0; g < kAttnGT; ++g) {
      const float keep = (sub & 2) ? a2[1][g] : a2[0][g];
      const float send = (sub & 2) ? a2[0][g] : a2[1][g];
      a1[g] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    float fin[2];
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const float keep = (sub & 1) ? a1[g + 2] : a1[g];
      const float send = (sub & 1) ? a1[g] : a1[g + 2];
      fin[g] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
    }
    const int it = ((sub >> 3) & 1) * 4 + ((sub >> 2) & 1) * 2 + ((sub >> 1) & 1);
    const int lkey = it * 2 + hw;                 // 0..15 within the group
    const int gsel = (sub & 1) * 2;
    const bool key_ok = kg + lkey < k_end;
    float s0 = key_ok ? fin[0] : -INFINITY, s1 = key_ok ? fin[1] : -INFINITY;   // already in log2 units
    // group max per head: reduce over the 16 lanes that share (sub&1), then swap halves
    float mx0 = s0, mx1 = s1;
#pragma unroll
    for (int o = 2; o <= 16; o <<= 1) {
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, o));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, o));
    }
    const float ox0 = __shfl_xor_sync(0xffffffffu, mx0, 1), ox1 = __shfl_xor_sync(0xffffffffu, mx1, 1);
    const bool odd = (sub & 1) != 0;              // odd lanes hold heads 2,3; even lanes heads 0,1
    float gmax[kAttnGT];
    gmax[0] = odd ? ox0 : mx0; gmax[1] = odd ? ox1 : mx1; gmax[2] = odd ? mx0 : ox0; gmax[3] = odd ? mx1 : ox1;
    float corr[kAttnGT], m_new[kAttnGT];
#pragma unroll
    for (int g = 0; g < kAttnGT; ++g) {
      m_new[g] = fmaxf(m_run[g], gmax[g]);
      corr[g] = exp2f(m_run[g] - m_new[g]);      // 0 on the first group (m_run = -inf)
      m_run[g] = m_new[g];
    }
    const float p0 = exp2f(s0 - (odd ? m_new[2] : m_new[0])), p1 = exp2f(s1 - (odd ? m_new[3] : m_new[1]));
#pragma unroll
    for (int g = 0; g < kAttnGT; ++g) l_run[g] *= corr[g];
    l_run[0] += odd ? 0.f : p0;                   // lane-local partial sums, reduced at the end
    l_run[1] += odd ? 0.f : p1;
    l_run[2] += odd ? p0 : 0.f;
    l_run[3] += odd ? p1 : 0.f;
    __syncwarp();
    s_p[warp][lkey][gsel] = p0;
    s_p[warp][lkey][gsel + 1] = p1;
    __syncwarp();
#pragma unroll
    for (int g = 0; g < kAttnGT; ++g)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[g][e] *= corr[g];
#pragma unroll
    for (int it2 = 0; it2 < 8; ++it2) {
      const float4 pw = *reinterpret_cast<const float4 *>(&s_p[warp][it2 * 2 + hw][0]);
      const float pg[4] = {pw.x, pw.y, pw.z, pw.w};
      const uint32_t u[4] = {(uint32_t)vl[it2].x, (uint32_t)vl[it2].y, (uint32_t)vl[it2].z, (uint32_t)vl[it2].w};
      float x[8];
#pragma unroll
      for (int e = 0; e < 4; ++e) { x[2 * e] = bf16_lo(u[e]); x[2 * e + 1] = bf16_hi(u[e]); }
#pragma unroll
      for (int g = 0; g < kAttnGT; ++g)
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[g][e] = fmaf(pg[g], x[e], acc[g][e]);
    }
  }

  SKV_STAMP(row_id == 0 ? split : 999, 2);
  // ---- merge inside the warp (two key parities), then across the warps of the CTA ---------------------
#pragma unroll
  for (int g = 0; g < kAttnGT; ++g) {
    l_run[g] = warp_sum(l_run[g]);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[g][e] += __shfl_xor_sync(0xffffffffu, acc[g][e], 16);
  }
  if (hw == 0) {
#pragma unroll
    for (int g = 0; g < kAttnGT; ++g)
#pragma unroll
      for (int e = 0; e < 8; ++e) s_o[warp][g][sub * 8 + e] = acc[g][e];
  }
  if (lane == 0) {
#pragma unroll
    for (int g = 0; g < kAttnGT; ++g) { s_ml[warp][g][0] = m_run[g]; s_ml[warp][g][1] = l_run[g]; }
  }
  __syncthreads();
  if (threadIdx.x >= 256) return;                 // warps beyond 8 only stream keys; none of them can be thread 0
  const int d = threadIdx.x & (kHeadDim - 1);     // 256 threads = 128 dims x 2 head pairs
  const int gh = threadIdx.x >> 7;
  float *po = p.part_o + (size_t(row_id) * kAttnMaxSplits + split) * kAttnGT * kHeadDim;
  float *pml = p.part_ml + (size_t(row_id) * kAttnMaxSplits + split) * kAttnGT * 2;
#pragma unroll
  for (int gg = 0; gg < kAttnGT / 2; ++gg) {
    const int g = gh * (kAttnGT / 2) + gg;
    float m = -INFINITY;
    for (int w = 0; w < nwarps; ++w) m = fmaxf(m, s_ml[w][g][0]);
    float l = 0.f, o = 0.f;
#pragma unroll 4
    for (int w = 0; w < nwarps; ++w) {
      const float mw = s_ml[w][g][0];
      const float f = (mw == -INFINITY) ? 0.f : exp2f(mw - m);
      l += s_ml[w][g][1] * f;
      o += s_o[w][g][d] * f;
    }
    if (p.num_splits == 1) {                       // single split: finish here, no workspace round trip
      if (g0 + g < p.G)
        p.out[(size_t(b) * p.H * p.G + size_t(h) * p.G + g0 + g) * kHeadDim + d] = f32_to_bf16_bits(o / l);
    } else {
      __stcg(po + g * kHeadDim + d, o);
      if (d == 0) { __stcg(pml + g * 2, m); __stcg(pml + g * 2 + 1, l); }
    }
  }
  if (p.num_splits == 1) return;
  SKV_STAMP(row_id == 0 ? split : 999, 3);
  // ---- last-arriving CTA of this (b, head group) merges the splits ---------------------------------
  __threadfence();
  bar256();
  if (threadIdx.x == 0) {
    const int prev = atomicAdd(&p.counters[row_id], 1);
    s_last = (prev == p.num_splits - 1);
    if (s_last) p.counters[row_id] = 0;           // ready for the next launch
  }
  bar256();
  SKV_STAMP(row_id == 0 ? split : 999, 4);
  if (!s_last) return;
  __threadfence();
  SKV_STAMP(row_id == 0 ? split : 999, 5);
  const float *ro = p.part_o + size_t(row_id) * kAttnMaxSplits * kAttnGT * kHeadDim;
  const float *rml = p.part_ml + size_t(row_id) * kAttnMaxSplits * kAttnGT * 2;
  // weights of every (split, head): one warp per head, lanes over the splits (<= 64)
  if (warp < kAttnGT) {
    const int g = warp;
    float ms[2], ls[2], m = -INFINITY;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int sidx = lane + 32 * k;
      const bool on = sidx < p.num_splits;
      ms[k] = on ? __ldcg(rml + (sidx * kAttnGT + g) * 2) : -INFINITY;
      ls[k] = on ? __ldcg(rml + (sidx * kAttnGT + g) * 2 + 1) : 0.f;
      m = fmaxf(m, ms[k]);
    }
    m = warp_max(m);
    float lsum = 0.f;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const float f = (ms[k] == -INFINITY) ? 0.f : exp2f(ms[k] - m);
      lsum += ls[k] * f;
      if (lane + 32 * k < kAttnMaxSplits) s_f[lane + 32 * k][g] = f;
    }
    lsum = warp_sum(lsum);
    if (lane == 0) s_L[g] = lsum;
  }
  bar256();
  SKV_STAMP(row_id == 0 ? split : 999, 6);
  {
    const int g = threadIdx.x >> 6, c2 = threadIdx.x & 63;         // 256 threads = 4 heads x 64 dim pairs
    if (g0 + g < p.G) {
      const float *col = ro + size_t(g) * kHeadDim + c2 * 2;
      const int ns = p.num_splits;
      float ox = 0.f, oy = 0.f;
      int sidx = 0;
      for (; sidx + 8 <= ns; sidx += 8) {                          // eight independent 8-byte loads in flight
        float2 v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = __ldcg(reinterpret_cast<const float2 *>(col + size_t(sidx + k) * kAttnGT * kHeadDim));
#pragma unroll
        for (int k = 0; k < 8; ++k) { ox = fmaf(v[k].x, s_f[sidx + k][g], ox); oy = fmaf(v[k].y, s_f[sidx + k][g], oy); }
      }
      float2 v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k)
        v[k] = sidx + k < ns ? __ldcg(reinterpret_cast<const float2 *>(col + size_t(sidx + k) * kAttnGT * kHeadDim))
                             : make_float2(0.f, 0.f);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (sidx + k < ns) { ox = fmaf(v[k].x, s_f[sidx + k][g], ox); oy = fmaf(v[k].y, s_f[sidx + k][g], oy); }
      const float inv = 1.0f / s_L[g];
      const __nv_bfloat162 o2 = __floats2bfloat162_rn(ox * inv, oy * inv);
      *reinterpret_cast<__nv_bfloat162 *>(p.out + (size_t(b) * p.H * p.G + size_t(h) * p.G + g0 + g) * kHeadDim + c2 * 2) = o2;
    }
  }
  SKV_STAMP(row_id == 0 ? split : 999, 7);
}

// (a2) append: one row of K and V per (b,h)
__global__ void append_kv_kernel(const int4 *__restrict__ k_new, const int4 *__restrict__ v_new, int4 *__restrict__ k_buf,
                                 int4 *__restrict__ v_buf, int BH, int64_t buf_rows, int pos,
                                 const int32_t *__restrict__ pos_dev) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;      // BH * 16 int4
  if (i >= BH * 16) return;
  if (pos_dev) pos += *pos_dev;
  if (pos >= buf_rows) return;                              // generation region full: drop (host checks too)
  const int bh = i >> 4, c = i & 15;
  const size_t dst = (size_t(bh) * buf_rows + pos) * 16 + c;
  k_buf[dst] = k_new[i];
  v_buf[dst] = v_new[i];
}

}  // namespace skv

using namespace skv;

SKV_TRACE_READER(attention)

extern "C" {

int64_t skv_attention_workspace_bytes(int batch, int kv_heads, int q_groups, int head_dim, int max_splits) {
  (void)head_dim; (void)max_splits;
  const int64_t rows = int64_t(batch) * kv_heads * ((q_groups + kAttnGT - 1) / kAttnGT);
  return rows * kAttnMaxSplits * kAttnGT * (kHeadDim + 2) * 4 + rows * 4 + 256;
}

int skv_attention_decode(const skv_bf16 *q, const skv_bf16 *k, const skv_bf16 *v, skv_bf16 *out, int batch,
                         int kv_heads, int q_groups, int head_dim, int kv_len, int64_t stride_b, int64_t stride_h,
                         int64_t stride_row, float scale, int num_splits, void *workspace,
                         const int32_t *kv_len_dev, skv_stream_t stream) {
  return skv_attention_decode_ex(q, k, v, out, batch, kv_heads, q_groups, head_dim, kv_len, stride_b, stride_h, stride_row,
                                 scale, num_splits, workspace, kv_len_dev, 0, stream);
}

int skv_attention_decode_ex(const skv_bf16 *q, const skv_bf16 *k, const skv_bf16 *v, skv_bf16 *out, int batch,
                            int kv_heads, int q_groups, int head_dim, int kv_len, int64_t stride_b, int64_t stride_h,
                            int64_t stride_row, float scale, int num_splits, void *workspace,
                            const int32_t *kv_len_dev, int flags, skv_stream_t stream) {
  if (!q || !k || !v || !out || !workspace) { set_error("attention_decode: null pointer"); return SKV_ERR_ARGUMENT; }
  if (head_dim != kHeadDim || q_groups < 1 || kv_len < 1 || (stride_row % 8) || (stride_h % 8) || (stride_b % 8)) {
    set_error("attention_decode: need head_dim=128 and 16-byte aligned strides"); return SKV_ERR_UNSUPPORTED;
  }
  static const bool simt = [] { const char *e = getenv("SKV_ATTN_IMPL"); return e && e[0] == '1'; }();
  if (!simt && q_groups <= 16)   // product path; SKV_ATTN_IMPL=1 selects the SIMT kernel below (cross-check)
    return launch_attention_mma(q, k, v, out, batch, kv_heads, q_groups, kv_len, stride_b, stride_h, stride_row, scale,
                                num_splits, workspace, kv_len_dev, (flags & SKV_ATTN_PROGRAMMATIC) != 0, as_stream(stream));
  AttnParams p{};
  p.q = q; p.k = k; p.v = v; p.out = out; p.kv_len_dev = kv_len_dev;
  p.B = batch; p.H = kv_heads; p.G = q_groups; p.ngz = (q_groups + kAttnGT - 1) / kAttnGT; p.kv_len = kv_len;
  p.sb = stride_b; p.sh = stride_h; p.sr = stride_row;
  p.scale_log2 = scale * 1.4426950408889634f;
  const int64_t rows = int64_t(batch) * kv_heads * p.ngz;
  // Split planning: one CTA per SM (register footprint), so the launch costs ceil(rows*ns/148) waves of
  // (bytes streamed per SM + per-iteration latency + a fixed prologue), plus the merge of the partials.  Pick the split count that minimises
  // it, then size the CTA (8..12 warps, 16 keys per warp per iteration) so that a split is ONE iteration
  // per warp when the cache is short (batch 1: 8 rows x 18 splits x 9 warps = 144 CTAs, a single wave).
  int ns = 1, kps = kv_len;
  if (num_splits > 0) {
    ns = num_splits > kAttnMaxSplits ? kAttnMaxSplits : num_splits;
  } else {
    double best = 1e30;
    for (int c = 1; c <= kAttnMaxSplits; ++c) {
      const int keys = ((kv_len + c - 1) / c + kAttnGroupKeys - 1) / kAttnGroupKeys * kAttnGroupKeys;
      const int64_t waves = (rows * c + kNumSMs - 1) / kNumSMs;
      const int iters = (keys + kAttnMaxWarps * kAttnGroupKeys - 1) / (kAttnMaxWarps * kAttnGroupKeys);
      const double cost = double(waves) * (keys * 0.08 + iters * 8.0 + 6.0) + (c > 1 ? 4.0 + 0.25 * c : 0.0);
      if (cost < best - 1e-9) { best = cost; ns = c; }
    }
  }
  kps = ((kv_len + ns - 1) / ns + kAttnGroupKeys - 1) / kAttnGroupKeys * kAttnGroupKeys;
  ns = (kv_len + kps - 1) / kps;
  int warps = (kps / kAttnGroupKeys);
  if (warps > kAttnMaxWarps) {                       // several iterations per warp: balance them
    const int iters = (warps + kAttnMaxWarps - 1) / kAttnMaxWarps;
    warps = (warps + iters - 1) / iters;
  }
  if (warps < kAttnMinWarps) warps = kAttnMinWarps;
  p.keys_per_split = kps; p.num_splits = ns;
  unsigned char *w = static_cast<unsigned char *>(workspace);
  p.part_o = reinterpret_cast<float *>(w);
  p.part_ml = p.part_o + rows * kAttnMaxSplits * kAttnGT * kHeadDim;
  p.counters = reinterpret_cast<int *>(p.part_ml + rows * kAttnMaxSplits * kAttnGT * 2);
  dim3 grid(ns, kv_heads * p.ngz, batch);
  attention_decode_kernel<<<grid, warps * 32, 0, as_stream(stream)>>>(p);
  return launch_status("attention_decode_kernel");
}

int skv_append_kv(const skv_geom *g, const skv_bf16 *k_new, const skv_bf16 *v_new, skv_bf16 *k_buf, skv_bf16 *v_buf,
                  int pos, const int32_t *pos_dev, skv_stream_t stream) {
  if (!g || !k_new || !v_new || !k_buf || !v_buf) { set_error("append_kv: null pointer"); return SKV_ERR_ARGUMENT; }
  if (g->head_dim != kHeadDim || pos < 0 || pos >= g->buf_len) { set_error("append_kv: pos %d outside buffer", pos); return SKV_ERR_ARGUMENT; }
  const int BH = g->batch * g->kv_heads;
  append_kv_kernel<<<(BH * 16 + 127) / 128, 128, 0, as_stream(stream)>>>(
      reinterpret_cast<const int4 *>(k_new), reinterpret_cast<const int4 *>(v_new), reinterpret_cast<int4 *>(k_buf),
      reinterpret_cast<int4 *>(v_buf), BH, g->buf_len, pos, pos_dev);
  return launch_status("append_kv_kernel");
}

}  // extern "C"
