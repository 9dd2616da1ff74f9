// Split-K decode attention, product path: K / V tiles of 64 keys stream through a TMA-fed, 128-byte-swizzled
// shared-memory ring; four consumer warps run QK^T and PV on the tensor cores (mma.sync m16n8k16, bf16 in, f32
// accumulate; the <= 16 query heads of one kv head are the M dimension, so every K / V byte is read ONCE for
// all heads of the group) with an online softmax in the log2 domain; partials of the splits are merged by the
// last-arriving CTA.  Replaces flash_attn_with_kvcache at reference models/base.py:341.
//
// HBM bound at batch (2*H*L'*D*2 bytes per sequence), a latency chain at batch 1 -- hence: one 16-key slice per
// warp per tile, everything else (loads, barrier waits) off the math warps' critical path.
#include "mma_sync.cuh"

namespace skv {

constexpr int kAmTileKeys = 64;
constexpr int kAmMaxStages = 4;                          // ring depth is chosen per launch (2..4)
constexpr int kAmSlabBytes = kAmTileKeys * 128;          // 64 keys x 64 dims (128 B), swizzled
constexpr int kAmTileBytes = 4 * kAmSlabBytes;           // K lo, K hi, V lo, V hi = 32 KiB
constexpr int kAmConsumers = 4;                          // warps; each owns 16 keys of a tile
constexpr int kAmThreads = 32 * (1 + kAmConsumers);
constexpr int kAmMaxSplits = 64;
constexpr int kAmHeads = 16;                             // MMA M: query heads of the group, zero padded
constexpr int kAmOStride = kHeadDim + 4;                 // padded row of the per-warp O staging (bank spread)

struct AttnMmaParams {
  const skv_bf16 *q;        // [B, H*G, 128]
  skv_bf16 *out;            // [B, H*G, 128]
  float *part_o;            // [B*H, 64 splits, Gp, 128]
  float *part_ml;           // [B*H, 64 splits, Gp, 2]
  int *counters;            // [B*H]
  const int32_t *kv_len_dev;
  int H, G, Gp, kv_len, keys_per_split, num_splits;
  int stages;               // ring depth of this launch
  int o_rows;               // head rows of the per-warp staging: 8 when G <= 8, else 16
  float scale_log2;
};

__global__ void __launch_bounds__(kAmThreads, 1)
    attention_mma_kernel(const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v,
                         AttnMmaParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char *ring = smem;                                                  // p.stages x 32 KiB
  float *s_o = reinterpret_cast<float *>(smem + p.stages * kAmTileBytes);       // [4 warps][o_rows heads][132]
  float *s_ml = s_o + kAmConsumers * p.o_rows * kAmOStride;                     // [4][16][2]
  float *s_f = s_ml + kAmConsumers * kAmHeads * 2;                              // [64 splits][16 heads]
  float *s_L = s_f + kAmMaxSplits * kAmHeads;                                   // [16]
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(s_L + kAmHeads);
  uint64_t *empty_bar = full_bar + kAmMaxStages;
  int *s_last = reinterpret_cast<int *>(empty_bar + kAmMaxStages);

  pdl_launch_dependents();
  const int split = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row_id = b * p.H + h;
  const int G = p.G;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbarrier_init(&full_bar[s], 1); mbarrier_init(&empty_bar[s], kAmConsumers); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_k) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v) : "memory");
  }
  __syncthreads();
  // under a programmatic launch everything above overlapped the predecessor's tail; q, the K / V buffers and the
  // length counter are (transitively) the predecessors' outputs
  pdl_wait();

  const int kv_len = p.kv_len_dev ? min(p.kv_len, *p.kv_len_dev) : p.kv_len;
  const int k_begin = split * p.keys_per_split;
  const int k_end = min(kv_len, k_begin + p.keys_per_split);
  const int n_tiles = k_end > k_begin ? (k_end - k_begin + kAmTileKeys - 1) / kAmTileKeys : 0;

  if (warp == kAmConsumers) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0, ph = 0;
      for (int t = 0; t < n_tiles; ++t) {
        mbarrier_wait(&empty_bar[stage], ph ^ 1);
        mbarrier_expect_tx(&full_bar[stage], kAmTileBytes);
        unsigned char *dst = ring + stage * kAmTileBytes;
        const int row0 = k_begin + t * kAmTileKeys;          // rows past the tensor's extent are zero filled
        tma_load_4d(dst, &map_k, 0, row0, h, b, &full_bar[stage]);
        tma_load_4d(dst + kAmSlabBytes, &map_k, 64, row0, h, b, &full_bar[stage]);
        tma_load_4d(dst + 2 * kAmSlabBytes, &map_v, 0, row0, h, b, &full_bar[stage]);
        tma_load_4d(dst + 3 * kAmSlabBytes, &map_v, 64, row0, h, b, &full_bar[stage]);
        if (++stage == p.stages) { stage = 0; ph ^= 1; }
      }
    }
  } else {
    // ===== consumers: warp w owns keys [16w, 16w+16) of every tile =====
    const int g = lane >> 2, tig = lane & 3;                 // MMA fragment coordinates
    // Q as the A operand (rows = heads g and g+8, zero beyond G), all 8 k-steps resident: 32 registers
    uint32_t qa[8][4];
    {
      const skv_bf16 *q0 = p.q + (size_t(b) * p.H * G + size_t(h) * G) * kHeadDim;
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const int d0 = kk * 16 + tig * 2;
        qa[kk][0] = g < G ? *reinterpret_cast<const uint32_t *>(q0 + size_t(g) * kHeadDim + d0) : 0u;
        qa[kk][1] = g + 8 < G ? *reinterpret_cast<const uint32_t *>(q0 + size_t(g + 8) * kHeadDim + d0) : 0u;
        qa[kk][2] = g < G ? *reinterpret_cast<const uint32_t *>(q0 + size_t(g) * kHeadDim + d0 + 8) : 0u;
        qa[kk][3] = g + 8 < G ? *reinterpret_cast<const uint32_t *>(q0 + size_t(g + 8) * kHeadDim + d0 + 8) : 0u;
      }
    }
    float o[16][4];                                          // 16 dim tiles x (row g: c0,c1 | row g+8: c2,c3)
#pragma unroll
    for (int j = 0; j < 16; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f; }
    float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;   // rows g and g+8 (l: this thread's keys only)

    // ldmatrix lane addressing inside a 64-key x 64-dim swizzled slab: row = key, 16-byte chunk c -> c ^ (key & 7)
    const int lm_m = lane >> 3, lm_r = lane & 7;
    const int k_key = warp * 16 + (lm_m >> 1) * 8 + lm_r;    // K (non-transposed): matrices (keys lo|hi) x (dims lo8|hi8)
    const int v_key = warp * 16 + (lm_m & 1) * 8 + lm_r;     // V (transposed):     matrices (dims lo8|hi8) x (keys lo|hi)

    int stage = 0, ph = 0;
    for (int t = 0; t < n_tiles; ++t) {
      const int key0 = k_begin + t * kAmTileKeys + warp * 16;            // first key of this warp's slice
      mbarrier_wait(&full_bar[stage], ph);
      if (key0 < k_end) {                                                // warp uniform
        const uint32_t kbase = smem_addr(ring + stage * kAmTileBytes);
        const uint32_t vbase = kbase + 2 * kAmSlabBytes;
        // ---- S = Q K^T for 16 keys: two n-tiles of 8 keys, 8 k-steps ------------------------------------------
        float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const int c = kk * 2 + (lm_m & 1);                             // 16-byte chunk of the 256-byte row
          uint32_t kb[4];
          ldmatrix_x4(kb, kbase + (c >> 3) * kAmSlabBytes + k_key * 128 + (((c & 7) ^ (k_key & 7)) << 4));
          mma_bf16_16816(s0, qa[kk], kb[0], kb[1]);
          mma_bf16_16816(s1, qa[kk], kb[2], kb[3]);
        }
        // ---- online softmax (log2 domain).  Thread holds keys {2tig, 2tig+1, 8+2tig, 9+2tig} of rows g, g+8 ----
        const int kcol = key0 + tig * 2;
        const bool v00 = kcol < k_end, v01 = kcol + 1 < k_end, v10 = kcol + 8 < k_end, v11 = kcol + 9 < k_end;
        float t_lo[4] = {v00 ? s0[0] * p.scale_log2 : -INFINITY, v01 ? s0[1] * p.scale_log2 : -INFINITY,
                         v10 ? s1[0] * p.scale_log2 : -INFINITY, v11 ? s1[1] * p.scale_log2 : -INFINITY};
        float t_hi[4] = {v00 ? s0[2] * p.scale_log2 : -INFINITY, v01 ? s0[3] * p.scale_log2 : -INFINITY,
                         v10 ? s1[2] * p.scale_log2 : -INFINITY, v11 ? s1[3] * p.scale_log2 : -INFINITY};
        float mx_lo = fmaxf(fmaxf(t_lo[0], t_lo[1]), fmaxf(t_lo[2], t_lo[3]));
        float mx_hi = fmaxf(fmaxf(t_hi[0], t_hi[1]), fmaxf(t_hi[2], t_hi[3]));
        mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1));
        mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1));
        mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
        mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
        const float mn_lo = fmaxf(m_lo, mx_lo), mn_hi = fmaxf(m_hi, mx_hi);   // finite: key0 < k_end
        const float c_lo = exp2f(m_lo - mn_lo), c_hi = exp2f(m_hi - mn_hi);   // 0 on the first tile
        m_lo = mn_lo; m_hi = mn_hi;
        float p_lo[4], p_hi[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) { p_lo[e] = exp2f(t_lo[e] - mn_lo); p_hi[e] = exp2f(t_hi[e] - mn_hi); }
        l_lo = l_lo * c_lo + (p_lo[0] + p_lo[1]) + (p_lo[2] + p_lo[3]);
        l_hi = l_hi * c_hi + (p_hi[0] + p_hi[1]) + (p_hi[2] + p_hi[3]);
#pragma unroll
        for (int j = 0; j < 16; ++j) { o[j][0] *= c_lo; o[j][1] *= c_lo; o[j][2] *= c_hi; o[j][3] *= c_hi; }
        // P (16 heads x 16 keys) as the A operand of P.V
        uint32_t pa[4] = {pack_bf16x2(p_lo[0], p_lo[1]), pack_bf16x2(p_hi[0], p_hi[1]),
                          pack_bf16x2(p_lo[2], p_lo[3]), pack_bf16x2(p_hi[2], p_hi[3])};
        // ---- O += P V: 16 n-tiles of 8 dims, one k-step of 16 keys ---------------------------------------------
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const int c = jj * 2 + (lm_m >> 1);
          uint32_t vb[4];
          ldmatrix_x4_trans(vb, vbase + (c >> 3) * kAmSlabBytes + v_key * 128 + (((c & 7) ^ (v_key & 7)) << 4));
          mma_bf16_16816(o[2 * jj], pa, vb[0], vb[1]);
          mma_bf16_16816(o[2 * jj + 1], pa, vb[2], vb[3]);
        }
      }
      __syncwarp();
      if (lane == 0) mbarrier_arrive(&empty_bar[stage]);
      if (++stage == p.stages) { stage = 0; ph ^= 1; }
    }
    // ---- per-warp result to shared memory ---------------------------------------------------------------
    l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1); l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1);
    l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2); l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
    float *wo = s_o + warp * p.o_rows * kAmOStride;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      *reinterpret_cast<float2 *>(wo + g * kAmOStride + j * 8 + tig * 2) = make_float2(o[j][0], o[j][1]);
      if (p.o_rows > 8) *reinterpret_cast<float2 *>(wo + (g + 8) * kAmOStride + j * 8 + tig * 2) = make_float2(o[j][2], o[j][3]);
    }
    if (tig == 0) {
      s_ml[(warp * kAmHeads + g) * 2] = m_lo; s_ml[(warp * kAmHeads + g) * 2 + 1] = l_lo;
      s_ml[(warp * kAmHeads + g + 8) * 2] = m_hi; s_ml[(warp * kAmHeads + g + 8) * 2 + 1] = l_hi;
    }
  }
  __syncthreads();
  if (warp == kAmConsumers) return;                          // the producer warp is done
  // ---- merge the four warps: thread = dim ------------------------------------------------------------------
  const int d = threadIdx.x;                                 // 0..127
  float *po = p.part_o + (size_t(row_id) * kAmMaxSplits + split) * p.Gp * kHeadDim;
  float *pml = p.part_ml + (size_t(row_id) * kAmMaxSplits + split) * p.Gp * 2;
  for (int hq = 0; hq < G; ++hq) {
    float m = -INFINITY;
#pragma unroll
    for (int w = 0; w < kAmConsumers; ++w) m = fmaxf(m, s_ml[(w * kAmHeads + hq) * 2]);
    float l = 0.f, acc = 0.f;
#pragma unroll
    for (int w = 0; w < kAmConsumers; ++w) {
      const float mw = s_ml[(w * kAmHeads + hq) * 2];
      const float f = (mw == -INFINITY) ? 0.f : exp2f(mw - m);
      l += s_ml[(w * kAmHeads + hq) * 2 + 1] * f;
      acc += s_o[(w * p.o_rows + hq) * kAmOStride + d] * f;
    }
    if (p.num_splits == 1) {
      p.out[(size_t(b) * p.H * G + size_t(h) * G + hq) * kHeadDim + d] = f32_to_bf16_bits(acc / l);
    } else {
      __stcg(po + hq * kHeadDim + d, acc);
      if (d == 0) { __stcg(pml + hq * 2, m); __stcg(pml + hq * 2 + 1, l); }
    }
  }
  if (p.num_splits == 1) return;
  // ---- last-arriving CTA of this (b, kv head) merges the splits -------------------------------------------
  __threadfence();
  asm volatile("bar.sync 1, 128;" ::: "memory");
  if (threadIdx.x == 0) {
    const int prev = atomicAdd(&p.counters[row_id], 1);
    *s_last = (prev == p.num_splits - 1);
    if (*s_last) p.counters[row_id] = 0;                     // ready for the next launch
  }
  asm volatile("bar.sync 1, 128;" ::: "memory");
  if (!*s_last) return;
  __threadfence();
  const float *ro = p.part_o + size_t(row_id) * kAmMaxSplits * p.Gp * kHeadDim;
  const float *rml = p.part_ml + size_t(row_id) * kAmMaxSplits * p.Gp * 2;
  // weights of every (split, head): warp w takes heads w, w+4, ...; lanes run over the splits (<= 64)
  for (int hq = warp; hq < G; hq += kAmConsumers) {
    float ms[2], ls[2], m = -INFINITY;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int sidx = lane + 32 * k;
      const bool on = sidx < p.num_splits;
      ms[k] = on ? __ldcg(rml + (size_t(sidx) * p.Gp + hq) * 2) : -INFINITY;
      ls[k] = on ? __ldcg(rml + (size_t(sidx) * p.Gp + hq) * 2 + 1) : 0.f;
      m = fmaxf(m, ms[k]);
    }
    m = warp_max(m);
    float lsum = 0.f;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const float f = (ms[k] == -INFINITY) ? 0.f : exp2f(ms[k] - m);
      lsum += ls[k] * f;
      s_f[(lane + 32 * k) * kAmHeads + hq] = f;
    }
    lsum = warp_sum(lsum);
    if (lane == 0) s_L[hq] = lsum;
  }
  asm volatile("bar.sync 1, 128;" ::: "memory");
  {
    // thread = (head slot of 4, dim quad): 128 threads cover 4 heads x 32 float4 columns per pass
    const int ns = p.num_splits;
    for (int h0 = 0; h0 < G; h0 += 4) {
      const int hq = h0 + (threadIdx.x >> 5), c4 = threadIdx.x & 31;
      if (hq < G) {
        const float *col = ro + size_t(hq) * kHeadDim + c4 * 4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int sidx = 0;
        for (; sidx + 4 <= ns; sidx += 4) {                  // four independent 16-byte loads in flight
          float4 v[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) v[k] = __ldcg(reinterpret_cast<const float4 *>(col + size_t(sidx + k) * p.Gp * kHeadDim));
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float f = s_f[(sidx + k) * kAmHeads + hq];
            acc.x = fmaf(v[k].x, f, acc.x); acc.y = fmaf(v[k].y, f, acc.y);
            acc.z = fmaf(v[k].z, f, acc.z); acc.w = fmaf(v[k].w, f, acc.w);
          }
        }
        for (; sidx < ns; ++sidx) {
          const float4 v = __ldcg(reinterpret_cast<const float4 *>(col + size_t(sidx) * p.Gp * kHeadDim));
          const float f = s_f[sidx * kAmHeads + hq];
          acc.x = fmaf(v.x, f, acc.x); acc.y = fmaf(v.y, f, acc.y); acc.z = fmaf(v.z, f, acc.z); acc.w = fmaf(v.w, f, acc.w);
        }
        const float inv = 1.0f / s_L[hq];
        uint2 packed;
        packed.x = pack_bf16x2(acc.x * inv, acc.y * inv);
        packed.y = pack_bf16x2(acc.z * inv, acc.w * inv);
        *reinterpret_cast<uint2 *>(p.out + (size_t(b) * p.H * G + size_t(h) * G + hq) * kHeadDim + c4 * 4) = packed;
      }
    }
  }
}

// 4-D bf16 tensor map over a K or V cache view: dims {128, rows, heads, batch}, strides in elements, box
// {64 dims, 64 rows, 1, 1}, 128-byte swizzle.  Cached per (pointer, shape, strides).
int make_tensor_map_kv(CUtensorMap *out, const void *base, uint64_t rows, uint64_t heads, uint64_t batch, int64_t stride_row,
                       int64_t stride_h, int64_t stride_b);

int launch_attention_mma(const skv_bf16 *q, const skv_bf16 *k, const skv_bf16 *v, skv_bf16 *out, int batch, int kv_heads,
                         int q_groups, int kv_len, int64_t stride_b, int64_t stride_h, int64_t stride_row, float scale,
                         int num_splits, void *workspace, const int32_t *kv_len_dev, bool pdl, cudaStream_t st) {
  CUtensorMap map_k, map_v;
  int rc = make_tensor_map_kv(&map_k, k, kv_len, kv_heads, batch, stride_row, stride_h, stride_b);
  if (rc) return rc;
  rc = make_tensor_map_kv(&map_v, v, kv_len, kv_heads, batch, stride_row, stride_h, stride_b);
  if (rc) return rc;
  AttnMmaParams p{};
  p.q = q; p.out = out; p.kv_len_dev = kv_len_dev;
  p.H = kv_heads; p.G = q_groups; p.Gp = (q_groups + 3) / 4 * 4; p.kv_len = kv_len;
  p.scale_log2 = scale * 1.4426950408889634f;
  const int64_t rows = int64_t(batch) * kv_heads;
  // Split planning: one CTA per SM (128 KiB ring).  Cost of a split count c: waves of CTAs x (tiles per CTA + a
  // fixed prologue / merge worth ~3 tiles) + the merge of c partials.  Splits start at tile boundaries when that
  // costs nothing; at batch 1 (8 rows) the winner is 18 splits of 144 keys = 144 CTAs, one wave.
  int ns = 1;
  if (num_splits > 0) {
    ns = num_splits > kAmMaxSplits ? kAmMaxSplits : num_splits;
  } else {
    double best = 1e30;
    for (int c = 1; c <= kAmMaxSplits; ++c) {
      const int keys = ((kv_len + c - 1) / c + 15) / 16 * 16;
      const int tiles = (keys + kAmTileKeys - 1) / kAmTileKeys;
      const int64_t waves = (rows * c + kNumSMs - 1) / kNumSMs;
      const double cost = double(waves) * (tiles + 3.0) + (c > 1 ? 1.0 + 0.05 * c : 0.0);
      if (cost < best - 1e-9) { best = cost; ns = c; }
    }
  }
  int kps = ((kv_len + ns - 1) / ns + 15) / 16 * 16;
  ns = (kv_len + kps - 1) / kps;
  p.keys_per_split = kps; p.num_splits = ns;
  unsigned char *w = static_cast<unsigned char *>(workspace);
  p.part_o = reinterpret_cast<float *>(w);
  p.part_ml = p.part_o + rows * kAmMaxSplits * p.Gp * kHeadDim;
  p.counters = reinterpret_cast<int *>(p.part_ml + rows * kAmMaxSplits * p.Gp * 2);
  // Ring depth: no deeper than the tiles a CTA streams (2..4).  With a short ring (batch 1: 3 tiles) the CTA leaves
  // room on its SM for one scoring CTA of the next layer, which then streams its landmarks behind this kernel
  // (programmatic dependent launch).
  const int tiles_per_cta = (kps + kAmTileKeys - 1) / kAmTileKeys;        // tiles are cut from the split's first key
  p.stages = tiles_per_cta < 2 ? 2 : (tiles_per_cta > kAmMaxStages ? kAmMaxStages : tiles_per_cta);
  p.o_rows = q_groups <= 8 ? 8 : 16;
  const int smem = p.stages * kAmTileBytes + (kAmConsumers * p.o_rows * kAmOStride + kAmConsumers * kAmHeads * 2 +
                                             kAmMaxSplits * kAmHeads + kAmHeads) * 4 + 2 * kAmMaxStages * 8 + 16 + 64;
  if (int rc2 = ensure_dynamic_smem(reinterpret_cast<const void *>(attention_mma_kernel), smem, "attention_mma_kernel")) return rc2;
  cudaError_t e = launch_kernel(attention_mma_kernel, dim3(ns, kv_heads, batch), dim3(kAmThreads), smem, st, pdl,
                                map_k, map_v, p);
  if (e != cudaSuccess) { set_error("attention_mma_kernel: %s", cudaGetErrorString(e)); return int(e); }
  return launch_status("attention_mma_kernel");
}

}  // namespace skv
