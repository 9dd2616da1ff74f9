// Dense remainder of the decode layer: one decode token per sequence through
//   rmsnorm -> fused QKV (+bias) -> RoPE(q,k)            reference models/llama.py:354-385, glm.py:312-357
//   o_proj + residual -> rmsnorm -> gate/up -> SiLU*mul -> down + residual      llama.py:387-427, glm.py:359-403
//   final rmsnorm -> lm_head -> argmax                                        base.py:150-159, tensor_op.py:291-293
// The reference runs these through library kernels (flashinfer rmsnorm, cuBLAS F.linear, vllm rotary_embedding /
// silu_and_mul, torch adds): ~11 launches per layer.  With batch <= 64 every linear is a weight stream
// (0.436 GB per Llama-8B layer against 8 KB of activations per sequence), i.e. HBM bound, so the design is one
// weight-streaming kernel with all element-wise work in its epilogue:
//
//  * W [out, in] is cut into tiles of 128 rows x 64-column slabs (16 KiB).  The (tile, slab) space is linearised
//    and split EVENLY over one persistent CTA per SM (stream-K): every SM streams the same number of bytes whatever
//    the matrix shape (4096 x 4096 has only 32 row tiles; 148 SMs).
//  * warp 0 = TMA producer: W slab (two 64-row boxes, 128-byte swizzle, L2 evict-first) + the matching 64 columns
//    of x (rows beyond the batch are zero filled by TMA) into a shared-memory ring.  The weights are static, so under
//    a programmatic dependent launch the ring is filled BEFORE griddepcontrol.wait, while the previous kernel drains.
//  * warp 1 = MMA issuer: tcgen05.mma M=128 (W rows) x N=16/32/64 (batch, padded) x K=16, fp32 accumulators in TMEM,
//    double buffered so the stream continues while a finished tile is in its epilogue.
//  * warps 2..5 = epilogue, thread = W row.  A tile whose slabs were split between CTAs is finished by the
//    last-arriving CTA, which sums the partials in CTA order (deterministic), then: bias, bf16 round, residual add /
//    SiLU*mul (gate and up rows interleaved into the same tile) / rotary embedding (one tile = one head).
#include "sm100_ptx.cuh"

#include <cstdlib>

namespace skv {

constexpr int kDnWBytes = 128 * 128;          // W part of a stage: 128 rows x 64 bf16
constexpr int kDnMaxStages = 12;
constexpr int kDnThreads = 64 + 128;
constexpr int kDnMaxBatch = 64;
constexpr int kDnMaxCtas = 2 * kNumSMs;
constexpr int kDnMaxTiles = 4096;             // arrival counters in the workspace
constexpr int64_t kDnCounterBytes = kDnMaxTiles * 4;
constexpr int64_t kDnPartBytes = int64_t(kDnMaxCtas) * 2 * kDnMaxBatch * 128 * 4;

struct DenseParams {
  int B, K, KS, Nout, T, mode, stages, total;
  float *part;            // [ctas][2][NPAD][128 rows] fp32 partial tiles
  int *tile_cnt;          // [T] arrivals per split tile; zero between launches
  const __nv_bfloat16 *resid, *bias, *cos_sin;
  __nv_bfloat16 *out, *k_out, *v_out;
  float *out_f32;
  int up_row0;            // SILU_MUL: first "up" row of W (= F); tile t = gate rows [64t, 64t+64) + up rows [F+64t, ..)
  int q_heads, kv_heads, rope_mode, cos_sin_rows, pos;
  int head0;              // QKV_ROPE: W holds the heads [head0, head0 + T) of the fused wqkv
  int kv_row;             // QKV_ROPE with kv_stride_h != 0: k / v go to row kv_row (+ *pos_dev) of a cache buffer
  int kv_rows;            // rows of that buffer: a row at or beyond it is dropped (as append_kv_kernel does)
  long long kv_stride_b, kv_stride_h;
  int static_w;           // W may be read before griddepcontrol.wait
  int trace_id;           // SKV_TRACE builds: launch counter
  int reorder;            // split tiles first (SegOrder)
  const unsigned char *pf_ptr;   // next linear's weights: pulled into L2 once this CTA has issued its last load
  long long pf_bytes;
  const int *pos_dev;
};

__device__ __forceinline__ int dn_begin(int c, int total, int n) { return int((long long)c * total / n); }
__device__ __forceinline__ int dn_cta_of(int w, int total, int n) {
  int c = int((long long)w * n / total);
  while (c + 1 < n && dn_begin(c + 1, total, n) <= w) ++c;
  while (c > 0 && dn_begin(c, total, n) > w) --c;
  return c;
}
// The slabs [w0, w1) of one CTA as tile segments, in PROCESSING order: the segment that finishes a tile begun by the
// previous CTA first, then the segment that begins a tile the next CTA finishes, then the tiles the CTA owns alone.
// Split tiles need a cross-CTA fix-up (partials through L2, last arriver sums); doing them first hides that latency
// behind the stream of the complete tiles instead of leaving it at the end of the kernel.
struct SegOrder {
  int w0, w1, KS, ta, n, last_partial;
  __device__ SegOrder(int w0_, int w1_, int KS_, int reorder) : w0(w0_), w1(w1_), KS(KS_) {
    ta = w0 / KS;
    const int tb = (w1 - 1) / KS;
    n = w1 > w0 ? tb - ta + 1 : 0;
    last_partial = (reorder && n >= 2 && (w1 % KS) != 0) ? 1 : 0;
  }
  __device__ void get(int i, int &t, int &s0, int &s1) const {
    if (i == 0) { t = ta; s0 = w0 - ta * KS; s1 = min(KS, w1 - ta * KS); return; }
    if (last_partial) {
      if (i == 1) { t = ta + n - 1; s0 = 0; s1 = w1 - t * KS; return; }
      t = ta + i - 1; s0 = 0; s1 = KS; return;
    }
    t = ta + i; s0 = 0; s1 = min(KS, w1 - t * KS);
  }
};

__device__ __forceinline__ float bfr(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

template <int NPAD>
__device__ __forceinline__ void dense_epilogue(const DenseParams &p, int t, int row, float (&v)[NPAD], float *s_x,
                                               int pos) {
  constexpr int LD = NPAD + 1;
  if (p.mode == SKV_DENSE_F32) {
    const int n = t * 128 + row;
    if (n < p.Nout) {
#pragma unroll
      for (int b = 0; b < NPAD; ++b)
        if (b < p.B) p.out_f32[size_t(b) * p.Nout + n] = bfr(v[b]);
    }
  } else if (p.mode == SKV_DENSE_RESID) {
    const int n = t * 128 + row;
    if (n < p.Nout) {
      const float bias = p.bias ? __bfloat162float(p.bias[n]) : 0.f;
      // all residual loads first: out may alias resid (same element, same thread), so the compiler cannot hoist them
      // over the stores by itself and the batch rows would pay one L2 round trip each
      float rr[NPAD];
#pragma unroll
      for (int b = 0; b < NPAD; ++b) rr[b] = (p.resid && b < p.B) ? __bfloat162float(p.resid[size_t(b) * p.Nout + n]) : 0.f;
#pragma unroll
      for (int b = 0; b < NPAD; ++b)
        if (b < p.B) {
          float y = bfr(v[b] + bias);
          if (p.resid) y = bfr(rr[b] + y);
          p.out[size_t(b) * p.Nout + n] = __float2bfloat16_rn(y);
        }
    }
  } else if (p.mode == SKV_DENSE_SILU_MUL) {
    // rows 0..63 = gate, 64..127 = up of the same 64 output features (vllm silu_and_mul: silu in fp32, rounded to
    // bf16, then a bf16 multiply)
    if (row >= 64) {
#pragma unroll
      for (int b = 0; b < NPAD; ++b)
        if (b < p.B) s_x[(row - 64) * LD + b] = bfr(v[b]);
    }
    epi_bar();
    if (row < 64) {
      const int n = t * 64 + row;
#pragma unroll
      for (int b = 0; b < NPAD; ++b)
        if (b < p.B) {
          const float g = bfr(v[b]);
          const float sl = bfr(g / (1.0f + expf(-g)));
          p.out[size_t(b) * p.up_row0 + n] = __float2bfloat16_rn(sl * s_x[row * LD + b]);
        }
    }
    epi_bar();
  } else {  // SKV_DENSE_QKV_ROPE: tile = head
    const int th = t + p.head0;                                    // head index in the fused wqkv numbering
    const int n = th * 128 + row;
    const float bias = p.bias ? __bfloat162float(p.bias[n]) : 0.f;
    const bool is_q = th < p.q_heads, is_k = !is_q && th < p.q_heads + p.kv_heads;
    const bool rope = p.rope_mode != 0 && (is_q || is_k);          // uniform over the tile
#pragma unroll
    for (int b = 0; b < NPAD; ++b) v[b] = bfr(v[b] + bias);
    if (rope) {
#pragma unroll
      for (int b = 0; b < NPAD; ++b)
        if (b < p.B) s_x[row * LD + b] = v[b];
      epi_bar();
      // vllm rotary_embedding: first = x*cos - y*sin, second = y*cos + x*sin, evaluated in fp32 and rounded once (what
      // the vllm build of this image does, bit for bit: tests/golden/dense_golden.npz)
      int partner = -1;
      bool first = true;
      float cs = 1.f, sn = 0.f;
      if (p.rope_mode == 1) {
        const int j = row & 63;
        first = row < 64;
        partner = first ? row + 64 : row - 64;
        cs = __bfloat162float(p.cos_sin[size_t(pos) * 128 + j]);
        sn = __bfloat162float(p.cos_sin[size_t(pos) * 128 + 64 + j]);
      } else if (row < 64) {
        const int j = row >> 1;
        first = (row & 1) == 0;
        partner = row ^ 1;
        cs = __bfloat162float(p.cos_sin[size_t(pos) * 64 + j]);
        sn = __bfloat162float(p.cos_sin[size_t(pos) * 64 + 32 + j]);
      }
      if (partner >= 0) {
#pragma unroll
        for (int b = 0; b < NPAD; ++b)
          if (b < p.B) {
            const float other = s_x[partner * LD + b];
            v[b] = first ? bfr(v[b] * cs - other * sn) : bfr(v[b] * cs + other * sn);
          }
      }
      epi_bar();
    }
    if (is_q || p.kv_stride_h == 0) {
      __nv_bfloat16 *dst;
      int heads, head;
      if (is_q) { dst = p.out; heads = p.q_heads; head = th; }
      else if (is_k) { dst = p.k_out; heads = p.kv_heads; head = th - p.q_heads; }
      else { dst = p.v_out; heads = p.kv_heads; head = th - p.q_heads - p.kv_heads; }
#pragma unroll
      for (int b = 0; b < NPAD; ++b)
        if (b < p.B) dst[(size_t(b) * heads + head) * 128 + row] = __float2bfloat16_rn(v[b]);
    } else {
      // append in place: row kv_row (+ the device-side generation counter) of the layer's K / V buffer [B,H,rows,128]
      __nv_bfloat16 *dst = is_k ? p.k_out : p.v_out;
      const int head = is_k ? th - p.q_heads : th - p.q_heads - p.kv_heads;
      const long long r = (long long)p.kv_row + (p.pos_dev ? *p.pos_dev : 0);
      const bool in_range = r >= 0 && r < p.kv_rows;       // generation region full: drop the row, never write past it
#pragma unroll
      for (int b = 0; b < NPAD; ++b)
        if (b < p.B && in_range) dst[b * p.kv_stride_b + head * p.kv_stride_h + r * 128 + row] = __float2bfloat16_rn(v[b]);
    }
  }
}

template <int NPAD>
__global__ void __launch_bounds__(kDnThreads, 1)
    dense_streamk_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x,
                         DenseParams p) {
  constexpr int kStageBytes = kDnWBytes + NPAD * 128;
  constexpr int kAccCols = NPAD;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char *ring = smem;
  float *s_x = reinterpret_cast<float *>(smem + p.stages * kStageBytes);
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(s_x + 128 * (NPAD + 1));
  uint64_t *empty_bar = full_bar + kDnMaxStages;
  uint64_t *acc_full = empty_bar + kDnMaxStages;
  uint64_t *acc_empty = acc_full + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);
  int *s_flag = reinterpret_cast<int *>(tmem_slot + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x, nc = gridDim.x;
  const int w0 = dn_begin(c, p.total, nc), w1 = dn_begin(c + 1, p.total, nc);
  const SegOrder order(w0, w1, p.KS, p.reorder);
  const int stages = p.stages;

  auto tile_rows = [&](int t, int &r0, int &r1) {
    if (p.mode == SKV_DENSE_SILU_MUL) { r0 = t * 64; r1 = p.up_row0 + t * 64; }
    else { r0 = t * 128; r1 = r0 + 64; }
  };

  // trace rows: 16 per launch (first and last 8 CTAs), 16 launches
  const int trow = (c < 8 ? c : (c >= nc - 8 ? 8 + c - (nc - 8) : -1000)) + (p.trace_id & 15) * 16;
  if (threadIdx.x == 0) SKV_STAMP_NS(trow, 0);
  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbarrier_init(&full_bar[s], 1); mbarrier_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbarrier_init(&acc_full[a], 1); mbarrier_init(&acc_empty[a], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // The weights do not depend on the previous kernel: fill the ring now (under a programmatic launch this runs
    // while the predecessor is still finishing).
    const uint64_t pol = l2_policy_evict_first();
    if (p.static_w) {
      int it = 0;
      for (int i = 0; i < order.n && it < stages; ++i) {
        int t, s0, s1, r0, r1;
        order.get(i, t, s0, s1);
        tile_rows(t, r0, r1);
        for (int s = s0; s < s1 && it < stages; ++s, ++it) {
          unsigned char *dst = ring + it * kStageBytes;
          mbarrier_expect_tx(&full_bar[it], kStageBytes);
          tma_load_2d_hint(dst, &map_w, s * 64, r0, &full_bar[it], pol);
          tma_load_2d_hint(dst + kDnWBytes / 2, &map_w, s * 64, r1, &full_bar[it], pol);
        }
      }
    }
  }
  if (threadIdx.x == 0) SKV_STAMP_NS(trow, 1);
  if (warp == 1) tcgen05_alloc(tmem_slot, 2 * kAccCols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) SKV_STAMP_NS(trow, 3);
  // x, the residual, the workspace and every output belong to the predecessor until here
  pdl_wait();
  if (threadIdx.x == 0) SKV_STAMP_NS(trow, 2);

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      const uint64_t pol = l2_policy_evict_first();
      int it = 0;
      for (int i = 0; i < order.n; ++i) {
        int t, s0, s1, r0, r1;
        order.get(i, t, s0, s1);
        tile_rows(t, r0, r1);
        for (int s = s0; s < s1; ++s, ++it) {
          const int stage = it % stages;
          unsigned char *dst = ring + stage * kStageBytes;
          if (it >= stages || !p.static_w) {
            if (it >= stages) mbarrier_wait(&empty_bar[stage], ((it / stages) & 1) ^ 1);
            mbarrier_expect_tx(&full_bar[stage], kStageBytes);
            tma_load_2d_hint(dst, &map_w, s * 64, r0, &full_bar[stage], pol);
            tma_load_2d_hint(dst + kDnWBytes / 2, &map_w, s * 64, r1, &full_bar[stage], pol);
          }
          tma_load_2d(dst + kDnWBytes, &map_x, s * 64, 0, &full_bar[stage]);
        }
      }
      // This CTA's stream is issued.  What follows in the kernel (split-tile fix-up, epilogues, hand-over to the next
      // kernel) leaves HBM idle for microseconds: use them to pull the head of the NEXT linear's weights into L2.
      if (p.pf_bytes > 0) {
        const long long slice = ((p.pf_bytes + nc - 1) / nc + 4095) & ~4095LL;
        const long long lo = (long long)c * slice, hi = min(p.pf_bytes, lo + slice);
        for (long long o = lo; o < hi; o += 4096) {
          const int n = int(min(4096LL, hi - o)) & ~15;
          if (n > 0) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.pf_ptr + o), "r"(n) : "memory");
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16_m128(kAccCols);
      int it = 0;
      for (int seg = 0; seg < order.n; ++seg) {
        int t, s0, s1;
        order.get(seg, t, s0, s1);
        const int acc = seg & 1;
        mbarrier_wait(&acc_empty[acc], ((seg >> 1) & 1) ^ 1);
        tcgen05_fence_after();
        const uint32_t d = tmem_base + acc * kAccCols;
        for (int s = s0; s < s1; ++s, ++it) {
          const int stage = it % stages;
          mbarrier_wait(&full_bar[stage], (it / stages) & 1);
          if (it == 0) SKV_STAMP_NS(trow, 4);
          tcgen05_fence_after();
          const uint64_t da = umma_desc_k_sw128(smem_addr(ring + stage * kStageBytes));
          const uint64_t db = umma_desc_k_sw128(smem_addr(ring + stage * kStageBytes + kDnWBytes));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            tcgen05_mma_bf16(d, da + uint64_t(kk * 2), db + uint64_t(kk * 2), idesc, (s > s0 || kk > 0) ? 1u : 0u);
          tcgen05_commit(&empty_bar[stage]);
        }
        tcgen05_commit(&acc_full[acc]);
      }
      SKV_STAMP_NS(trow, 5);
    }
  } else {
    // ===== epilogue: thread = W row of the tile =====
    const int lg = warp & 3;                       // TMEM lane group this warp may read
    const int row = lg * 32 + lane;
    const bool leader = threadIdx.x == 64;
    int pos = 0;
    if (p.mode == SKV_DENSE_QKV_ROPE && p.rope_mode) {
      pos = p.pos + (p.pos_dev ? *p.pos_dev : 0);
      pos = max(0, min(pos, p.cos_sin_rows - 1));
    }
    for (int seg = 0; seg < order.n; ++seg) {
      int t, s0, s1;
      order.get(seg, t, s0, s1);
      const bool final_seg = seg == order.n - 1;
      const int acc = seg & 1;
      mbarrier_wait(&acc_full[acc], (seg >> 1) & 1);
      if (leader && final_seg) SKV_STAMP_NS(trow, 6);
      tcgen05_fence_after();
      float v[NPAD];
      {
        uint32_t r[NPAD];
        const uint32_t taddr = tmem_base + (uint32_t(lg * 32) << 16) + acc * kAccCols;
        if constexpr (NPAD == 16) {
          tcgen05_ld16(taddr, r);
        } else if constexpr (NPAD == 32) {
          tcgen05_ld32(taddr, r);
        } else {                                       // 64 batch columns: two 32-column loads
          uint32_t lo[32], hi[32];
          tcgen05_ld32(taddr, lo);
          tcgen05_ld32(taddr + 32, hi);
          tcgen05_wait_ld();
#pragma unroll
          for (int b = 0; b < 32; ++b) { r[b] = lo[b]; r[32 + b] = hi[b]; }
        }
        tcgen05_wait_ld();
#pragma unroll
        for (int b = 0; b < NPAD; ++b) v[b] = __uint_as_float(r[b]);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbarrier_arrive(&acc_empty[acc]);

      if (p.mode == SKV_DENSE_PARTIAL) {
        // no reduction here: plane (this CTA - first CTA of the tile); the consumer adds the planes in CTA order
        const int j = c - dn_cta_of(t * p.KS, p.total, nc);
        const int n = t * 128 + row;
        if (n < p.Nout) {
#pragma unroll
          for (int b = 0; b < NPAD; ++b)
            if (b < p.B) p.out_f32[(size_t(j) * p.B + b) * p.Nout + n] = v[b];
        }
        continue;
      }
      if (!(s0 == 0 && s1 == p.KS)) {
        // split tile: park the partial sums; the CTA that arrives last adds them up in CTA order
        const int slot = (t * p.KS < w0) ? 0 : 1;
        constexpr int kPartFloats = 128 * NPAD;        // one partial tile: [batch column][row] fp32 -- a warp's 32 rows
                                                       // of one column are one 128-byte line (coalesced both ways)
        float *mine = p.part + (size_t(c) * 2 + slot) * kPartFloats + row;
#pragma unroll
        for (int b = 0; b < NPAD; ++b)
          if (b < p.B) __stcg(mine + b * 128, v[b]);
        __threadfence();
        if (leader && final_seg) SKV_STAMP_NS(trow, 7);
        epi_bar();
        if (leader) {
          const int cf = dn_cta_of(t * p.KS, p.total, nc), cl = dn_cta_of(t * p.KS + p.KS - 1, p.total, nc);
          const int old = atomicAdd(&p.tile_cnt[t], 1);
          const int last = old == cl - cf;
          if (last) p.tile_cnt[t] = 0;                 // every contributor has arrived: ready for the next launch
          s_flag[0] = last; s_flag[1] = cf; s_flag[2] = cl;
        }
        epi_bar();
        const int last = s_flag[0], cf = s_flag[1], cl = s_flag[2];
        epi_bar();                                     // s_flag may be rewritten by the next split tile
        if (leader && final_seg) SKV_STAMP_NS(trow, 8);
        if (!last) continue;
        __threadfence();
#pragma unroll
        for (int b = 0; b < NPAD; ++b) v[b] = 0.f;
        // sum in CTA order (deterministic); the loads of kGroup contributors are in flight together
        constexpr int kGroup = NPAD == 16 ? 8 : (NPAD == 32 ? 4 : 1);
        for (int c0 = cf; c0 <= cl; c0 += kGroup) {
          float tmp[kGroup][NPAD];
#pragma unroll
          for (int g = 0; g < kGroup; ++g) {
            const int cc = min(c0 + g, cl);
            const int sl = (t * p.KS < dn_begin(cc, p.total, nc)) ? 0 : 1;
            const float *src = p.part + (size_t(cc) * 2 + sl) * kPartFloats + row;
#pragma unroll
            for (int b = 0; b < NPAD; ++b)
              if (b < p.B) tmp[g][b] = __ldcg(src + b * 128);
          }
#pragma unroll
          for (int g = 0; g < kGroup; ++g) {
            if (c0 + g <= cl) {
#pragma unroll
              for (int b = 0; b < NPAD; ++b)
                if (b < p.B) v[b] += tmp[g][b];
            }
          }
        }
      }
      if (leader && final_seg) SKV_STAMP_NS(trow, 9);
      dense_epilogue<NPAD>(p, t, row, v, s_x, pos);
      if (leader && final_seg) SKV_STAMP_NS(trow, 10);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) SKV_STAMP_NS(trow, 11);
  if (warp == 1) tcgen05_dealloc(tmem_base, 2 * kAccCols);
}

// ---- rmsnorm (+ optional embedding lookup) ---------------------------------------------------------------
constexpr int kNormThreads = 512;
__global__ void __launch_bounds__(kNormThreads)
    rmsnorm_kernel(const __nv_bfloat16 *x, const int64_t *tok, const __nv_bfloat16 *emb, const __nv_bfloat16 *w, float eps,
                   __nv_bfloat16 *x_out, __nv_bfloat16 *h_out, int hidden) {
  __shared__ float s_part[kNormThreads / 32];
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  pdl_launch_dependents();
  pdl_wait();
  const __nv_bfloat16 *src = tok ? emb + size_t(tok[b]) * hidden : x + size_t(b) * hidden;
  const int nvec = hidden >> 3;
  int4 v[2];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int j = threadIdx.x + i * kNormThreads;
    v[i] = make_int4(0, 0, 0, 0);
    if (j < nvec) {
      v[i] = reinterpret_cast<const int4 *>(src)[j];
      const uint32_t *u = reinterpret_cast<const uint32_t *>(&v[i]);
#pragma unroll
      for (int k = 0; k < 4; ++k) { const float a = bf16_lo(u[k]), c = bf16_hi(u[k]); ss += a * a; ss += c * c; }
    }
  }
  ss = warp_sum(ss);
  if (lane == 0) s_part[warp] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < kNormThreads / 32; ++i) tot += s_part[i];
  const float r = rsqrtf(tot / float(hidden) + eps);
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int j = threadIdx.x + i * kNormThreads;
    if (j < nvec) {
      if (x_out && tok) reinterpret_cast<int4 *>(x_out + size_t(b) * hidden)[j] = v[i];
      const int4 wv = reinterpret_cast<const int4 *>(w)[j];
      const uint32_t *u = reinterpret_cast<const uint32_t *>(&v[i]);
      const uint32_t *uw = reinterpret_cast<const uint32_t *>(&wv);
      int4 o;
      uint32_t *uo = reinterpret_cast<uint32_t *>(&o);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        uo[k] = pack_bf16(__float2bfloat16_rn(bf16_lo(u[k]) * r * bf16_lo(uw[k])),
                          __float2bfloat16_rn(bf16_hi(u[k]) * r * bf16_hi(uw[k])));
      reinterpret_cast<int4 *>(h_out + size_t(b) * hidden)[j] = o;
    }
  }
}

// ---- consumer of SKV_DENSE_PARTIAL: planes -> (+bias) -> bf16 -> + residual -> rmsnorm ------------------------------
// One CTA per sequence; a thread owns 8 consecutive features (one tile, so one plane count).  The planes are added in
// CTA order, exactly like the in-kernel fix-up of the other modes.
__global__ void __launch_bounds__(kNormThreads)
    finish_rmsnorm_kernel(const float *planes, int B, int Nout, int KS, int total, int nc, const __nv_bfloat16 *bias,
                          __nv_bfloat16 *x, const __nv_bfloat16 *w, float eps, __nv_bfloat16 *h_out) {
  __shared__ float s_part[kNormThreads / 32];
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  pdl_launch_dependents();
  pdl_wait();
  const int nvec = Nout >> 3;
  float xv[2][8];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int j = threadIdx.x + i * kNormThreads;
    if (j < nvec) {
      const int n0 = j * 8, t = n0 >> 7;
      const int cf = dn_cta_of(t * KS, total, nc), cl = dn_cta_of(t * KS + KS - 1, total, nc);
      const int nseg = cl - cf + 1;
      float acc[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = 0.f;
      const int4 rv = reinterpret_cast<const int4 *>(x + size_t(b) * Nout)[j];
      for (int s0 = 0; s0 < nseg; s0 += 8) {
        float4 tmp[8][2];
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const int sj = min(s0 + g, nseg - 1);
          const float4 *src = reinterpret_cast<const float4 *>(planes + (size_t(sj) * B + b) * Nout + n0);
          tmp[g][0] = __ldcg(src); tmp[g][1] = __ldcg(src + 1);
        }
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          if (s0 + g < nseg) {
            acc[0] += tmp[g][0].x; acc[1] += tmp[g][0].y; acc[2] += tmp[g][0].z; acc[3] += tmp[g][0].w;
            acc[4] += tmp[g][1].x; acc[5] += tmp[g][1].y; acc[6] += tmp[g][1].z; acc[7] += tmp[g][1].w;
          }
        }
      }
      const uint32_t *ur = reinterpret_cast<const uint32_t *>(&rv);
      int4 bv = make_int4(0, 0, 0, 0);
      if (bias) bv = reinterpret_cast<const int4 *>(bias)[j];
      const uint32_t *ub = reinterpret_cast<const uint32_t *>(&bv);
      int4 o;
      uint32_t *uo = reinterpret_cast<uint32_t *>(&o);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float y0 = bfr(acc[2 * k] + bf16_lo(ub[k])), y1 = bfr(acc[2 * k + 1] + bf16_hi(ub[k]));
        const float x0 = bfr(bf16_lo(ur[k]) + y0), x1 = bfr(bf16_hi(ur[k]) + y1);
        xv[i][2 * k] = x0; xv[i][2 * k + 1] = x1;
        ss += x0 * x0; ss += x1 * x1;
        uo[k] = pack_bf16(__float2bfloat16_rn(x0), __float2bfloat16_rn(x1));
      }
      reinterpret_cast<int4 *>(x + size_t(b) * Nout)[j] = o;
    }
  }
  ss = warp_sum(ss);
  if (lane == 0) s_part[warp] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < kNormThreads / 32; ++i) tot += s_part[i];
  const float r = rsqrtf(tot / float(Nout) + eps);
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int j = threadIdx.x + i * kNormThreads;
    if (j < nvec) {
      const int4 wv = reinterpret_cast<const int4 *>(w)[j];
      const uint32_t *uw = reinterpret_cast<const uint32_t *>(&wv);
      int4 o;
      uint32_t *uo = reinterpret_cast<uint32_t *>(&o);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        uo[k] = pack_bf16(__float2bfloat16_rn(xv[i][2 * k] * r * bf16_lo(uw[k])),
                          __float2bfloat16_rn(xv[i][2 * k + 1] * r * bf16_hi(uw[k])));
      reinterpret_cast<int4 *>(h_out + size_t(b) * Nout)[j] = o;
    }
  }
}

// ---- greedy token ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) argmax_kernel(const float *logits, int64_t *tok, int vocab) {
  __shared__ float s_v[32];
  __shared__ int s_i[32];
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  pdl_launch_dependents();
  pdl_wait();
  const float *row = logits + size_t(b) * vocab;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < vocab; i += 1024) {
    const float x = row[i];
    if (x > best || (x == best && i < bi)) { best = x; bi = i; }
  }
  auto merge = [&](float ov, int oi) { if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; } };
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) merge(__shfl_xor_sync(0xffffffffu, best, o), __shfl_xor_sync(0xffffffffu, bi, o));
  if (lane == 0) { s_v[warp] = best; s_i[warp] = bi; }
  __syncthreads();
  if (warp == 0) {
    best = s_v[lane]; bi = s_i[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) merge(__shfl_xor_sync(0xffffffffu, best, o), __shfl_xor_sync(0xffffffffu, bi, o));
    if (lane == 0) tok[b] = bi == 0x7fffffff ? 0 : bi;
  }
}

// ---- L2 prefetch of a byte range (weights of an upcoming linear) --------------------------------------------
// One elected thread per warp issues bulk L2 prefetches of 8 KiB pieces; the kernel returns as soon as they are issued.
constexpr int kPfPiece = 8192;
__global__ void __launch_bounds__(128) l2_prefetch_kernel(const unsigned char *base, long long bytes) {
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  if (lane == 0) {
    for (long long off = warp * kPfPiece; off < bytes; off += nwarps * kPfPiece) {
      const long long n = min((long long)kPfPiece, bytes - off) & ~15LL;
      if (n > 0)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(base + off), "r"(int(n)) : "memory");
    }
  }
}

static int env_int(const char *name, int dflt) {
  const char *e = getenv(name);
  return (e && *e) ? atoi(e) : dflt;
}

// How a plain [out, in] linear splits over the SMs: shared by the launcher and by the consumer of SKV_DENSE_PARTIAL.
struct DenseSched { int T, KS, total, ctas, max_planes; };
static DenseSched dense_sched(int in_features, int out_features, int T) {
  DenseSched d{};
  d.T = T; d.KS = in_features / 64; d.total = d.T * d.KS;
  int ctas = kNumSMs * env_int("SKV_DENSE_CTAS_PER_SM", 1);
  if (ctas > kDnMaxCtas) ctas = kDnMaxCtas;
  if (ctas < 1) ctas = 1;
  if (ctas > d.total) ctas = d.total;
  d.ctas = ctas;
  const int share = d.total / ctas;                       // shortest slab range of a CTA
  d.max_planes = (d.KS + share - 1) / share + 1;          // CTAs that can touch one tile
  return d;
}

template <int NPAD>
static int launch_dense(const CUtensorMap &mw, const CUtensorMap &mx, DenseParams &p, int ctas, cudaStream_t st) {
  constexpr int kStageBytes = kDnWBytes + NPAD * 128;
  const int fixed = 128 * (NPAD + 1) * 4 + (2 * kDnMaxStages + 4) * 8 + 8 + 16 + 64;
  // default: 6 stages (96 KiB of weights in flight per SM, 14 MB over the GPU ~ the bandwidth-delay product).  Measured
  // on the lm_head stream: 5 stages 171 us, 6: 163 us (= the measured HBM copy peak), 8: 167 us, 11: 183 us -- a deeper
  // ring oversubscribes the memory system and the SMs finish up to 10 us apart.
  int stages = env_int("SKV_DENSE_STAGES", 6);
  if (stages < 2) stages = 2;
  if (stages > kDnMaxStages) stages = kDnMaxStages;
  while (stages > 2 && stages * kStageBytes + fixed > 227 * 1024) --stages;
  p.stages = stages;
  p.reorder = env_int("SKV_DENSE_ORDER", 1);
  static int launches = 0;
  p.trace_id = launches++;
  const int smem = stages * kStageBytes + fixed;
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void *>(dense_streamk_kernel<NPAD>), smem, "dense_streamk_kernel")) return rc;
  cudaError_t e = launch_kernel(dense_streamk_kernel<NPAD>, dim3(ctas), dim3(kDnThreads), smem, st, /*pdl=*/true, mw, mx, p);
  if (e != cudaSuccess) { set_error("dense_streamk_kernel: %s", cudaGetErrorString(e)); return int(e); }
  return launch_status("dense_streamk_kernel");
}

}  // namespace skv

using namespace skv;

extern "C" {

int64_t skv_dense_workspace_bytes(void) { return kDnCounterBytes + kDnPartBytes; }

int skv_dense_linear(const skv_bf16 *x, const skv_bf16 *W, int batch, int in_features, int out_features,
                     const skv_dense_epilogue *epi, void *workspace, skv_stream_t stream) {
  if (!x || !W || !epi || !workspace || !epi->out) { set_error("dense_linear: null pointer"); return SKV_ERR_ARGUMENT; }
  if (batch < 1 || batch > kDnMaxBatch || in_features < 64 || in_features % 64 || out_features < 1) {
    set_error("dense_linear: batch %d (1..%d), in_features %d (multiple of 64), out_features %d", batch, kDnMaxBatch,
              in_features, out_features);
    return SKV_ERR_UNSUPPORTED;
  }
  DenseParams p{};
  p.static_w = epi->static_weights != 0;
  if (epi->prefetch_next && epi->prefetch_bytes > 0 && (reinterpret_cast<uintptr_t>(epi->prefetch_next) & 15) == 0) {
    p.pf_ptr = reinterpret_cast<const unsigned char *>(epi->prefetch_next);
    p.pf_bytes = epi->prefetch_bytes;
  }
  p.B = batch; p.K = in_features; p.KS = in_features / 64; p.Nout = out_features; p.mode = epi->mode;
  p.tile_cnt = reinterpret_cast<int *>(workspace);
  p.part = reinterpret_cast<float *>(reinterpret_cast<char *>(workspace) + kDnCounterBytes);
  p.bias = reinterpret_cast<const __nv_bfloat16 *>(epi->bias);
  p.resid = reinterpret_cast<const __nv_bfloat16 *>(epi->resid);
  p.out = reinterpret_cast<__nv_bfloat16 *>(epi->out);
  p.out_f32 = reinterpret_cast<float *>(epi->out);
  switch (epi->mode) {
    case SKV_DENSE_F32:
    case SKV_DENSE_RESID:
    case SKV_DENSE_PARTIAL:
      p.T = (out_features + 127) / 128;
      break;
    case SKV_DENSE_SILU_MUL:
      if (out_features % 128) { set_error("dense_linear: silu_mul needs out_features = 2F, F %% 64 == 0"); return SKV_ERR_UNSUPPORTED; }
      p.up_row0 = out_features / 2;
      p.T = p.up_row0 / 64;
      break;
    case SKV_DENSE_QKV_ROPE:
      {
        const int all = epi->q_heads + 2 * epi->kv_heads;
        const int n_heads = epi->head_count > 0 ? epi->head_count : all - epi->head0;
        const bool has_kv = epi->head0 + n_heads > epi->q_heads;
        if (epi->q_heads < 1 || epi->kv_heads < 1 || epi->head0 < 0 || n_heads < 1 || epi->head0 + n_heads > all ||
            out_features != n_heads * 128 || (has_kv && (!epi->k_out || !epi->v_out)) || epi->rope_mode < 0 ||
            epi->rope_mode > 2 || (epi->rope_mode && (!epi->cos_sin || epi->cos_sin_rows < 1))) {
          set_error("dense_linear: qkv_rope needs out_features = 128 x (heads of the launch), k_out / v_out when the launch "
                    "covers k / v heads, and a cos_sin table");
          return SKV_ERR_ARGUMENT;
        }
        p.T = n_heads;
        p.head0 = epi->head0;
        p.kv_row = epi->kv_row; p.kv_stride_b = epi->kv_stride_b; p.kv_stride_h = epi->kv_stride_h;
        p.kv_rows = epi->kv_rows;
        if (has_kv && epi->kv_stride_h != 0 && (epi->kv_rows < 1 || epi->kv_row < 0 || epi->kv_row >= epi->kv_rows)) {
          set_error("dense_linear: in-place k / v append needs kv_rows (rows of the cache buffer) and 0 <= kv_row < kv_rows "
                    "(got row %d of %d)", epi->kv_row, epi->kv_rows);
          return SKV_ERR_ARGUMENT;
        }
      }
      p.q_heads = epi->q_heads; p.kv_heads = epi->kv_heads; p.rope_mode = epi->rope_mode;
      p.cos_sin = reinterpret_cast<const __nv_bfloat16 *>(epi->cos_sin);
      p.cos_sin_rows = epi->cos_sin_rows; p.pos = epi->pos; p.pos_dev = epi->pos_dev;
      p.k_out = reinterpret_cast<__nv_bfloat16 *>(epi->k_out);
      p.v_out = reinterpret_cast<__nv_bfloat16 *>(epi->v_out);
      break;
    default:
      set_error("dense_linear: unknown epilogue mode %d", epi->mode);
      return SKV_ERR_ARGUMENT;
  }
  if (p.T > kDnMaxTiles) { set_error("dense_linear: more than %d row tiles", kDnMaxTiles); return SKV_ERR_UNSUPPORTED; }
  const DenseSched sched = dense_sched(in_features, out_features, p.T);
  p.total = sched.total;
  const int ctas = sched.ctas;
  const int npad = batch <= 16 ? 16 : (batch <= 32 ? 32 : 64);
  CUtensorMap mw, mx;
  int rc = make_tensor_map_2d(&mw, W, uint64_t(out_features), uint64_t(in_features), 64, 64);
  if (rc) return rc;
  rc = make_tensor_map_2d(&mx, x, uint64_t(batch), uint64_t(in_features), uint32_t(npad), 64);
  if (rc) return rc;
  if (npad == 16) return launch_dense<16>(mw, mx, p, ctas, as_stream(stream));
  if (npad == 32) return launch_dense<32>(mw, mx, p, ctas, as_stream(stream));
  return launch_dense<64>(mw, mx, p, ctas, as_stream(stream));
}

int64_t skv_dense_partial_bytes(int batch, int in_features, int out_features) {
  if (batch < 1 || in_features < 64 || in_features % 64 || out_features < 1) return 0;
  const DenseSched d = dense_sched(in_features, out_features, (out_features + 127) / 128);
  return int64_t(d.max_planes) * batch * out_features * 4;
}

int skv_dense_finish_rmsnorm(const float *planes, int batch, int in_features, int out_features, const skv_bf16 *bias,
                             skv_bf16 *x_inout, const skv_bf16 *w, float eps, skv_bf16 *h_out, skv_stream_t stream) {
  if (!planes || !x_inout || !w || !h_out) { set_error("dense_finish_rmsnorm: null pointer"); return SKV_ERR_ARGUMENT; }
  if (batch < 1 || in_features < 64 || in_features % 64 || out_features < 8 || out_features % 8 ||
      out_features > 2 * kNormThreads * 8) {
    set_error("dense_finish_rmsnorm: in_features %d (multiple of 64), out_features %d (multiple of 8, <= %d)", in_features,
              out_features, 2 * kNormThreads * 8);
    return SKV_ERR_UNSUPPORTED;
  }
  const DenseSched d = dense_sched(in_features, out_features, (out_features + 127) / 128);
  cudaError_t e = launch_kernel(finish_rmsnorm_kernel, dim3(batch), dim3(kNormThreads), 0, as_stream(stream), true, planes,
                                batch, out_features, d.KS, d.total, d.ctas, reinterpret_cast<const __nv_bfloat16 *>(bias),
                                reinterpret_cast<__nv_bfloat16 *>(x_inout), reinterpret_cast<const __nv_bfloat16 *>(w), eps,
                                reinterpret_cast<__nv_bfloat16 *>(h_out));
  if (e != cudaSuccess) { set_error("finish_rmsnorm_kernel: %s", cudaGetErrorString(e)); return int(e); }
  return launch_status("finish_rmsnorm_kernel");
}

int skv_rmsnorm(const skv_bf16 *x, const int64_t *tok, const skv_bf16 *emb, const skv_bf16 *w, float eps, skv_bf16 *x_out,
                skv_bf16 *h_out, int batch, int hidden, skv_stream_t stream) {
  if ((!x && !tok) || (tok && !emb) || !w || !h_out) { set_error("rmsnorm: null pointer"); return SKV_ERR_ARGUMENT; }
  if (batch < 1 || hidden < 8 || hidden % 8 || hidden > 2 * kNormThreads * 8) {
    set_error("rmsnorm: hidden %d must be a multiple of 8, at most %d", hidden, 2 * kNormThreads * 8);
    return SKV_ERR_UNSUPPORTED;
  }
  cudaError_t e = launch_kernel(rmsnorm_kernel, dim3(batch), dim3(kNormThreads), 0, as_stream(stream), true,
                                reinterpret_cast<const __nv_bfloat16 *>(x), tok, reinterpret_cast<const __nv_bfloat16 *>(emb),
                                reinterpret_cast<const __nv_bfloat16 *>(w), eps, reinterpret_cast<__nv_bfloat16 *>(x_out),
                                reinterpret_cast<__nv_bfloat16 *>(h_out), hidden);
  if (e != cudaSuccess) { set_error("rmsnorm_kernel: %s", cudaGetErrorString(e)); return int(e); }
  return launch_status("rmsnorm_kernel");
}

int skv_l2_prefetch(const void *ptr, int64_t bytes, skv_stream_t stream) {
  if (!ptr || bytes < 0) { set_error("l2_prefetch: bad argument"); return SKV_ERR_ARGUMENT; }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) { set_error("l2_prefetch: pointer must be 16-byte aligned"); return SKV_ERR_ARGUMENT; }
  if (bytes < 16) return SKV_OK;
  long long pieces = (bytes + kPfPiece - 1) / kPfPiece;
  int grid = int(pieces < 4LL * kNumSMs ? (pieces + 3) / 4 : kNumSMs);
  cudaError_t e = launch_kernel(l2_prefetch_kernel, dim3(grid), dim3(128), 0, as_stream(stream), true,
                                reinterpret_cast<const unsigned char *>(ptr), (long long)bytes);
  if (e != cudaSuccess) { set_error("l2_prefetch_kernel: %s", cudaGetErrorString(e)); return int(e); }
  return launch_status("l2_prefetch_kernel");
}

int skv_argmax(const float *logits, int64_t *tok, int batch, int vocab, skv_stream_t stream) {
  if (!logits || !tok) { set_error("argmax: null pointer"); return SKV_ERR_ARGUMENT; }
  if (batch < 1 || vocab < 1) { set_error("argmax: empty input"); return SKV_ERR_UNSUPPORTED; }
  cudaError_t e = launch_kernel(argmax_kernel, dim3(batch), dim3(1024), 0, as_stream(stream), true, logits, tok, vocab);
  if (e != cudaSuccess) { set_error("argmax_kernel: %s", cudaGetErrorString(e)); return int(e); }
  return launch_status("argmax_kernel");
}

}  // extern "C"

SKV_TRACE_READER(dense)
