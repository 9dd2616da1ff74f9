// Landmark scoring, product path: a TMA-fed shared-memory ring streams the landmark matrix of one (b,h) once
// (HBM bound: N*D*2 bytes), the G <= 16 query heads are the N operand of a 128 x 16 x 128 tcgen05 MMA per
// 128-landmark tile (fp32 accumulators in TMEM, double buffered), and four epilogue warps turn each tile into
// bf16(alpha * q.k) scores plus the per-tile softmax partials that select_kernel consumes.
//
// Replaces the reference's batch_gemm_softmax GEMM + partial-softmax launch (kernels/batch_gemm_softmax.cu:229-293,
// CUTLASS 2.x Sm80 mma.sync, tile 32x256x32) on this path; same rounding points: fp32 accumulate, * alpha,
// round to bf16, max / sum-of-exp over the ROUNDED values (epilogue_visitor_with_softmax).
//
// Warp roles (64 + kEpi x 128 threads): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, then kEpi (4, or 2)
// epilogue warpgroups, each with its own TMEM accumulator, taking tiles g, g + kEpi, ... (warp w reads TMEM lanes
// 32*(w%4)..+31 = landmark rows of the tile; the per-tile epilogue is a latency chain of shuffles, so several tiles
// are kept in flight).  Two CTAs fit per SM (3 x 32 KiB ring each).
#include "sm100_ptx.cuh"

#include <cstdlib>

namespace skv {

constexpr int kSmTileRows = 128;                 // landmarks per tile = UMMA M
constexpr int kSmMaxStages = 3;                  // ring depth is a template parameter: 3 x 32 KiB, two CTAs per SM
                                                 // (default) or 2 x 32 KiB, three CTAs per SM (SKV_SCORE_STAGES=2)
constexpr int kSmSlabBytes = kSmTileRows * 128;  // one 64-column (128 B) swizzle slab of a tile
constexpr int kSmTileBytes = 2 * kSmSlabBytes;   // 32 KiB
constexpr int kSmQSlabBytes = 16 * 128;          // B operand: 16 rows (heads, zero padded) x 64 columns
constexpr int kSmMaxEpi = 4;                     // epilogue warpgroups (template parameter kEpi: 2 or 4), each with its own
                                                 // accumulator buffer; group g takes tiles g, g + kEpi, ...
constexpr int kSmAccCols = 16;                   // fp32 accumulator columns per tile (UMMA N)

struct ScoreMmaParams {
  const skv_bf16 *q;        // [BH, G, 128]
  const skv_bf16 *lm;       // [BH, N, 128] (the tensor map's base): L2 prefetch of the tiles beyond the ring
  long long lm_rows;        // BH * N
  int l2_prefetch;          // SKV_SCORE_L2PF (default 1)
  skv_bf16 *scores;         // [BH, G, ldn]
  float *part_max, *part_sum;   // [BH, G, T]
  int G, N, ldn, T, ctas_per_row;
  int dynamic_lm;           // landmarks may be the previous kernel's output: nothing is loaded ahead of griddepcontrol.wait
  float alpha;
  skv_append ap;
  int64_t buf_rows;
  const int32_t *wait_flags, *wait_epoch_dev;   // kv-head split: peers' flags of the previous layer (skv_geom.wait_flags)
  int wait_count;
  int wait_only;                                // SKV_GEOM_WAIT_FLAGS_ONLY: the flag words cover the predecessor's outputs
};

template <int kSmStages, int kCtasPerSm, int kEpi>
__global__ void __launch_bounds__(64 + kEpi * 128, kCtasPerSm)
    score_tcgen05_kernel(const __grid_constant__ CUtensorMap map_lm, ScoreMmaParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char *sA = smem;                                          // kSmStages x 32 KiB
  unsigned char *sQ = smem + kSmStages * kSmTileBytes;               // 2 slabs x 2 KiB
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(sQ + 2 * kSmQSlabBytes);
  uint64_t *empty_bar = full_bar + kSmStages;
  uint64_t *acc_full = empty_bar + kSmStages;                        // [kEpi]
  uint64_t *acc_empty = acc_full + kSmMaxEpi;                        // [kEpi]
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + kSmMaxEpi);
  float *s_red = reinterpret_cast<float *>(tmem_slot + 2);           // [kEpi acc][4 warps][16 heads][2]
  constexpr int kSmThreads = 64 + kEpi * 128;

  const int bh = blockIdx.y, cx = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_my = (p.T - cx + p.ctas_per_row - 1) / p.ctas_per_row;   // tiles cx, cx + cpr, ... < T
  const int tr = bh == 0 ? cx : 999;
  SKV_STAMP(tr, 0);
  // wall-clock rows of the trace build: row 0 = CTA (0,0), row 1 = the last CTA of the grid
  const int trn = (bh == 0 && cx == 0) ? 0 : ((bh == int(gridDim.y) - 1 && cx == int(gridDim.x) - 1) ? 1 : -1);
  if (threadIdx.x == 0) SKV_STAMP_NS(trn, 20);

  pdl_launch_dependents();                                   // the selection kernel may be scheduled behind us
  if (threadIdx.x == 0) {
    for (int s = 0; s < kSmStages; ++s) { mbarrier_init(&full_bar[s], 1); mbarrier_init(&empty_bar[s], 1); }
    for (int a = 0; a < kEpi; ++a) { mbarrier_init(&acc_full[a], 1); mbarrier_init(&acc_empty[a], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // the ring is empty: fill it right away.  The landmarks are static, so under a programmatic launch this
    // runs while the previous kernel of the stream (the previous layer's attention) is still finishing.
    const uint64_t pol = l2_policy_evict_first();            // landmarks are read once per step
    for (int i = 0; i < kSmStages && i < n_my && !p.dynamic_lm; ++i) {
      const int row0 = bh * p.N + (cx + i * p.ctas_per_row) * kSmTileRows;
      mbarrier_expect_tx(&full_bar[i], kSmTileBytes);
      tma_load_2d_hint(sA + i * kSmTileBytes, &map_lm, 0, row0, &full_bar[i], pol);
      tma_load_2d_hint(sA + i * kSmTileBytes + kSmSlabBytes, &map_lm, 64, row0, &full_bar[i], pol);
      if (i < 4) SKV_STAMP_ANY(tr, 2 + i);
    }
    // a CTA with one or two tiles more than the ring holds (batch 1: 4 tiles, 3 stages) pulls them into L2 now, so the
    // load issued behind the dependency wait is an L2 hit instead of an HBM round trip on the chain
    if (p.l2_prefetch && !p.dynamic_lm && n_my > kSmStages && n_my <= kSmStages + 2) {
      for (int i = kSmStages; i < n_my; ++i) {
        const long long row0 = (long long)bh * p.N + (long long)(cx + i * p.ctas_per_row) * kSmTileRows;
        const long long rows = min((long long)kSmTileRows, p.lm_rows - row0);
        if (rows > 0)
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.lm + row0 * kHeadDim), "r"(int(rows) * kHeadDim * 2) : "memory");
      }
    }
  }
  // TMEM does not depend on the predecessor either (attention and selection hold none; a rebuild CTA that is still
  // resident keeps 128 of the 512 columns)
  if (warp == 1) tcgen05_alloc(tmem_slot, kEpi * kSmAccCols);
  // everything below touches data of the predecessor (q, the K/V buffers, the score scratch).  With
  // SKV_GEOM_WAIT_FLAGS_ONLY the predecessor's merging CTAs raise one word per row behind the row's output, which is all
  // this launch consumes from it: the wait for the predecessor's COMPLETION moves to the end of the kernel (dependents
  // keep the transitive guarantee), and its completion + flush latency leaves the chain.
  const bool flags_only = p.wait_flags != nullptr && p.wait_only != 0;
  if (!flags_only) pdl_wait();
  if (threadIdx.x == 0) SKV_STAMP_NS(trn, 21);
  if (p.wait_flags) {
    // kv-head split: the query of this layer depends on the previous layer's GATHERED attention output; every peer raises
    // a flag word per row (system scope) behind the row's stores into this rank's buffer.
    const int want = *p.wait_epoch_dev;
    for (int i = threadIdx.x; i < p.wait_count; i += kSmThreads) {
      int seen;
      do {
        asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(seen) : "l"(p.wait_flags + i) : "memory");
      } while (seen < want);
    }
    __syncthreads();
  }

  // fused append (a2): the first CTA of every (b,h) writes the new token's K and V rows
  if (p.ap.k_new && cx == 0 && warp == 2) {
    const int64_t pos = p.ap.pos + (p.ap.pos_dev ? *p.ap.pos_dev : 0);
    if (pos < p.buf_rows) {
      const int c = lane & 15;
      const int4 *src = reinterpret_cast<const int4 *>(lane < 16 ? p.ap.k_new : p.ap.v_new) + size_t(bh) * 16 + c;
      int4 *dst = reinterpret_cast<int4 *>(lane < 16 ? p.ap.k_buf : p.ap.v_buf) + (size_t(bh) * p.buf_rows + pos) * 16 + c;
      *dst = *src;
    }
  }

  // B operand: the G query rows of this (b,h), zero padded to 16, written in the 128-byte-swizzled K-major
  // layout the MMA descriptor expects (chunk c of row r lands at chunk c ^ (r & 7) of its 128-byte line)
  for (int i = threadIdx.x; i < 2 * 16 * 8; i += kSmThreads) {
    const int slab = i >> 7, r = (i >> 3) & 15, c = i & 7;
    int4 v = make_int4(0, 0, 0, 0);
    if (r < p.G) v = *reinterpret_cast<const int4 *>(p.q + (size_t(bh) * p.G + r) * kHeadDim + slab * 64 + c * 8);
    *reinterpret_cast<int4 *>(sQ + slab * kSmQSlabBytes + (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)) = v;
  }
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  SKV_STAMP(tr, 1);
  if (threadIdx.x == 0) SKV_STAMP_NS(trn, 22);

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      const uint64_t pol = l2_policy_evict_first();
      for (int i = p.dynamic_lm ? 0 : kSmStages; i < n_my; ++i) {   // static landmarks: the first kSmStages tiles were issued during set-up
        const int stage = i % kSmStages, ph = (i / kSmStages) & 1;
        const int tile = cx + i * p.ctas_per_row;
        mbarrier_wait(&empty_bar[stage], ph ^ 1);
        mbarrier_expect_tx(&full_bar[stage], kSmTileBytes);
        const int row0 = bh * p.N + tile * kSmTileRows;      // rows past N belong to the next (b,h) or are zero
        tma_load_2d_hint(sA + stage * kSmTileBytes, &map_lm, 0, row0, &full_bar[stage], pol);
        tma_load_2d_hint(sA + stage * kSmTileBytes + kSmSlabBytes, &map_lm, 64, row0, &full_bar[stage], pol);
        if (i < 4) SKV_STAMP_ANY(tr, 2 + i);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16_m128(kSmAccCols);
      const uint64_t db0 = umma_desc_k_sw128(smem_addr(sQ)), db1 = umma_desc_k_sw128(smem_addr(sQ + kSmQSlabBytes));
      for (int i = 0; i < n_my; ++i) {
        const int stage = i % kSmStages, ph = (i / kSmStages) & 1;
        const int acc = i % kEpi, aph = (i / kEpi) & 1;
        mbarrier_wait(&acc_empty[acc], aph ^ 1);
        mbarrier_wait(&full_bar[stage], ph);
        if (i < 4) SKV_STAMP_ANY(tr, 6 + i);
        if (i < 4) SKV_STAMP_NS(trn, 24 + i);          // tile i of this CTA is in shared memory (wall clock)
        tcgen05_fence_after();
        const uint64_t da0 = umma_desc_k_sw128(smem_addr(sA + stage * kSmTileBytes));
        const uint64_t da1 = umma_desc_k_sw128(smem_addr(sA + stage * kSmTileBytes + kSmSlabBytes));
        const uint32_t d = tmem_base + acc * kSmAccCols;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) tcgen05_mma_bf16(d, da0 + uint64_t(kk * 2), db0 + uint64_t(kk * 2), idesc, kk > 0);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) tcgen05_mma_bf16(d, da1 + uint64_t(kk * 2), db1 + uint64_t(kk * 2), idesc, 1);
        tcgen05_commit(&empty_bar[stage]);                   // smem slot reusable once these MMAs retire
        tcgen05_commit(&acc_full[acc]);                      // accumulator ready for the epilogue
      }
    }
  } else {
    // ===== epilogue: thread = landmark row of the tile =====
    const int lg = warp & 3;                                 // TMEM lane group this warp may read
    const int grp = (warp - 2) >> 2;                         // epilogue warpgroup 0 / 1 <-> accumulator buffer
    const int ew = (warp - 2) & 3;                           // 0..3 inside the group
    for (int i = grp; i < n_my; i += kEpi) {
      const int acc = grp, aph = (i / kEpi) & 1;
      const int tile = cx + i * p.ctas_per_row;
      mbarrier_wait(&acc_full[acc], aph);
      if (i < 4 && (threadIdx.x == 64 || threadIdx.x == 192)) SKV_STAMP_ANY(tr, 10 + i);
      tcgen05_fence_after();
      uint32_t r[16];
      tcgen05_ld16(tmem_base + (uint32_t(lg * 32) << 16) + acc * kSmAccCols, r);
      tcgen05_wait_ld();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbarrier_arrive(&acc_empty[acc]);       // accumulator may be overwritten
      const int n = tile * kSmTileRows + lg * 32 + lane;
      const bool valid = n < p.N;
      float *red = s_red + (acc * 4 + ew) * 32;
      // four heads at a time so the shuffle chains of the reductions overlap
#pragma unroll
      for (int g0 = 0; g0 < 16; g0 += 4) {
        if (g0 < p.G) {                                      // warp uniform
          float x[4], m[4], e[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint16_t bits = f32_to_bf16_bits(__uint_as_float(r[g0 + k]) * p.alpha);
            const bool on = valid && g0 + k < p.G;
            if (on) p.scores[(size_t(bh) * p.G + g0 + k) * p.ldn + n] = bits;
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
      asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");   // the four warps of this group
      if (ew == 0 && lane < p.G) {
        const float *rb = s_red + acc * 4 * 32 + lane * 2;
        float m = fmaxf(fmaxf(rb[0], rb[32]), fmaxf(rb[64], rb[96]));
        float sum = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          const float mw = rb[w * 32];
          sum += (mw == -INFINITY) ? 0.f : rb[w * 32 + 1] * expf(mw - m);
        }
        p.part_max[(size_t(bh) * p.G + lane) * p.T + tile] = m;
        p.part_sum[(size_t(bh) * p.G + lane) * p.T + tile] = sum;
      }
      // s_red[acc] belongs to this group alone; a second barrier keeps the next tile of the group from
      // overwriting it while warp 0 of the group still combines
      asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
    }
  }
  if (threadIdx.x == 64) SKV_STAMP_ANY(tr, 14);
  tcgen05_fence_before();
  __syncthreads();
  SKV_STAMP(tr, 15);
  if (threadIdx.x == 0) SKV_STAMP_NS(trn, 23);
  if (warp == 1) tcgen05_dealloc(tmem_base, kEpi * kSmAccCols);
  if (flags_only) pdl_wait();                   // see above: completion of the predecessor, for the kernels behind this one
}

int launch_scores_tcgen05(const skv_bf16 *q, const skv_bf16 *lm, skv_bf16 *scores, float *pmax, float *psum, int BH, int G,
                          int N, int ldn, int T, float alpha, cudaStream_t st, const skv_append *append,
                          int64_t buf_rows, bool dynamic_landmarks, const int32_t *wait_flags,
                          const int32_t *wait_epoch_dev, int wait_count, bool wait_only) {
  CUtensorMap map;
  int rc = make_tensor_map_2d(&map, lm, uint64_t(BH) * N, kHeadDim, kSmTileRows, 64);
  if (rc) return rc;
  ScoreMmaParams p{};
  p.q = q; p.scores = scores; p.part_max = pmax; p.part_sum = psum;
  p.lm = lm; p.lm_rows = (long long)BH * N;
  { const char *le = getenv("SKV_SCORE_L2PF"); p.l2_prefetch = (le && le[0] == '0') ? 0 : 1; }
  p.G = G; p.N = N; p.ldn = ldn; p.T = T; p.alpha = alpha; p.buf_rows = buf_rows;
  p.dynamic_lm = dynamic_landmarks ? 1 : 0;
  if (append) p.ap = *append;
  if (wait_flags && wait_epoch_dev && wait_count > 0) {
    p.wait_flags = wait_flags; p.wait_epoch_dev = wait_epoch_dev; p.wait_count = wait_count;
    p.wait_only = wait_only ? 1 : 0;
  }
  // two (or three) CTAs per SM; every CTA of a row strides over that row's tiles
  static const int stages = [] { const char *e = getenv("SKV_SCORE_STAGES"); return (e && e[0] == '2') ? 2 : 3; }();
  static const int target = [] {
    const char *e = getenv("SKV_SCORE_CTAS");
    return (e && *e) ? atoi(e) : (stages == 2 ? 3 : 2) * kNumSMs;
  }();
  int cpr = (target + BH - 1) / BH;
  if (cpr > T) cpr = T;
  if (cpr < 1) cpr = 1;
  p.ctas_per_row = cpr;
  // epilogue warpgroups: four (576 threads, 48 registers), each with its own accumulator buffer.  Measured against two
  // (SKV_SCORE_EPI=2): neutral at batch 1, where the 8 dependent MMAs of a tile (~0.5 us) and the tile that is loaded
  // behind the dependency wait pace the CTA, not the epilogues; 47.5 -> 46.5 us per launch at batch 8 (122K)
  const char *ee = getenv("SKV_SCORE_EPI");
  const int epi = (ee && *ee && atoi(ee) == 2) ? 2 : 4;
  const int smem = stages * kSmTileBytes + 2 * kSmQSlabBytes + (2 * kSmMaxStages + 2 * kSmMaxEpi) * 8 + 8 + kSmMaxEpi * 4 * 32 * 4 + 64;
  const void *fn = stages == 2 ? reinterpret_cast<const void *>(score_tcgen05_kernel<2, 3, 2>)
                 : epi == 4    ? reinterpret_cast<const void *>(score_tcgen05_kernel<3, 2, 4>)
                               : reinterpret_cast<const void *>(score_tcgen05_kernel<3, 2, 2>);
  if (int rc2 = ensure_dynamic_smem(fn, smem, "score_tcgen05_kernel")) return rc2;
  cudaError_t e = stages == 2
      ? launch_kernel(score_tcgen05_kernel<2, 3, 2>, dim3(cpr, BH), dim3(64 + 2 * 128), smem, st, /*pdl=*/true, map, p)
      : epi == 4 ? launch_kernel(score_tcgen05_kernel<3, 2, 4>, dim3(cpr, BH), dim3(64 + 4 * 128), smem, st, /*pdl=*/true, map, p)
                 : launch_kernel(score_tcgen05_kernel<3, 2, 2>, dim3(cpr, BH), dim3(64 + 2 * 128), smem, st, /*pdl=*/true, map, p);
  if (e != cudaSuccess) { set_error("score_tcgen05_kernel: %s", cudaGetErrorString(e)); return int(e); }
  return launch_status("score_tcgen05_kernel");
}

}  // namespace skv

SKV_TRACE_READER(score_mma)
