// The back half of the decode step as ONE launch (north_star: "(5) is fused with (2) so rebuilt keys never round-trip
// HBM"): V miss fetch (a6) + low-rank K rebuild with RoPE (a7) + attention (a8) + split merge.
//
// The reference joins two streams in front of flash_attn_with_kvcache (models/base.py:326-341): the attention cannot
// start before the last missed V chunk has crossed PCIe, and it re-reads every rebuilt K row from HBM.  Here the grid
// has two kinds of CTAs that run CONCURRENTLY behind the selection kernel:
//
//   * tile CTAs (one per 16 entries of the miss list, as in fetch_rebuild_kernel): warps 0-3 rebuild the 128 K rows of
//     the tile (cp.async gather of U rows, SV by TMA, tcgen05.mma into TMEM, bf16 RoPE in the TMEM epilogue), write them
//     to the K buffer AND, from the same registers, take q.K for the G query heads of the kv head; warps 4-7 fetch the
//     tile's 16 V chunks (pinned host memory over PCIe, or the HBM tiers), write them to the V buffer AND stage them in
//     shared memory (the SV panel's space, free once the MMA has retired); the CTA then forms the softmax partial
//     (m, l, P.V) of its 128 keys on chip.  The rebuilt K rows are never read back from HBM.
//   * split CTAs (the split-K attention of attention_mma.cu: TMA ring + mma.sync): everything that is resident when the
//     selection finishes -- local window, outliers, the chunk-cache HITS (the in-place chunk table leaves them in their
//     slots), generated rows -- with the slots of the miss list masked out.  They run while the PCIe transfer is in
//     flight, so at batch 1 the attention leaves the critical path.
//
// The last CTA of a (b, kv head) row to arrive (a tile CTA whenever something crossed PCIe) merges all partials.
// In-place chunk table only (geom->table_mode == 1) and G <= 8; other cases keep skv_fetch_and_rebuild + attention.
#include "mma_sync.cuh"

namespace skv {

constexpr int kFaThreads = 256;
constexpr int kFaTileChunks = 16;                   // entries of the miss list per tile CTA
constexpr int kFaTileRows = kFaTileChunks * kChunkTokens;   // 128
constexpr int kFaMaxSlabs = 4;                      // rank <= 256
constexpr int kFaSlabBytes = 128 * 128;             // 16 KiB: 128 rows x 64 columns of bf16, 128-byte swizzled
constexpr int kFaStages = 3;                        // split CTAs: ring depth
constexpr int kFaTileKeys = 64;
constexpr int kFaKvSlab = kFaTileKeys * 128;        // 8 KiB: 64 keys x 64 dims
constexpr int kFaRingTile = 4 * kFaKvSlab;          // K lo, K hi, V lo, V hi = 32 KiB
constexpr int kFaMaxParts = 128;                    // <= 64 tiles (S = 1024) + <= 64 splits per row
constexpr int kFaMaxG = 8;
constexpr int kFaOStride = kHeadDim + 4;
constexpr int kFaVRingChunks = 4;                   // V chunks that are requested before the MMA has freed the operand area
constexpr int kFaTailBytes = 4096 + 1024 + kFaVRingChunks * kChunkBytes;   // q (8 x 128 f32) | barriers, indices,
                                                    // reduction scratch | staging of the first V chunks

struct FaParams {
  // K rebuild
  const int64_t *resident;     // [BH, S] chunk id per slot (after the selection kernel's update)
  const int32_t *cnts;         // [BH] chunk-cache hits
  const int32_t *miss_slots;   // [BH, S] slot of the i-th miss, i < S - cnt
  const __nv_bfloat16 *U;      // [B, L, R]
  const __nv_bfloat16 *cos_sin;
  __nv_bfloat16 *k_buf;        // [BH, Lbuf, 128]
  int64_t buf_stride_bh;       // elements
  int sparse_start;            // first row of the sparse region
  int H, S, L, R, n_slabs, k_steps, rope;
  // V fetch
  const skv_bf16 *v_src;       // chunk-major store (pinned host or HBM)
  int64_t v_src_stride_bh;     // elements
  skv_bf16 *v_buf;
  skv_vstore vs;
  // attention
  const skv_bf16 *q;           // [B, H*G, 128]
  skv_bf16 *out;               // [B, H*G, 128]
  float *part_o;               // [BH, kFaMaxParts, Gp, 128]
  float *part_ml;              // [BH, kFaMaxParts, Gp, 2]
  int *tickets, *done;         // [BH] each: arrival tickets / partials published (both left at zero)
  const int32_t *kv_len_dev;
  int kv_len, G, Gp, keys_per_split, n_splits, n_tiles, BH;
  int tail_off;                // byte offset of the fixed tail region in dynamic shared memory
  float scale_log2;
  skv_exchange xc;             // kv-head split: peers that receive this rank's attention output (world == 0: none)
  const int32_t *row_flags;    // [BH] or null: per-row hand-over from the selection kernel (skv_geom.row_flags)
  int early_splits;            // split CTAs request their first tiles ahead of the row hand-over (SKV_FA_EARLY, default 1)
  int order;                   // CTA order of the grid (SKV_FA_ORDER): 0 = a row's tiles together, 1 = tile index slowest,
                               // 2 = like 1 with the split CTAs in front of the second half of the tile indices
};

// Everything behind this point reads what the selection kernel wrote for row bh (chunk table, hit count, miss list) and,
// through it, its predecessors' outputs (the appended K / V row).  With row flags the CTA waits for ITS row's word, which
// the selection raises (release) as soon as the row's table is written, instead of for the completion and flush of the
// whole selection grid (~2 us later); the acquire makes the table -- and, by cumulativity, what the selection kernel's
// own dependency wait had made visible to it -- visible to this CTA.  The kernel still executes griddepcontrol.wait once,
// at its very end, so that ITS dependents keep the transitive guarantee of the chain.
__device__ __forceinline__ void fa_wait_row(const FaParams &p, const int bh) {
  if (!p.row_flags) { pdl_wait(); return; }
  if (threadIdx.x == 0) {
    int seen;
    do {
      asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(p.row_flags + bh) : "memory");
    } while (seen == 0);
  }
  __syncthreads();
}

__device__ __forceinline__ void fa_tile_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
// TMA bulk copy global (device or pinned host memory) -> shared memory, completion on an mbarrier
__device__ __forceinline__ void bulk_load(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_addr(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}
__device__ __forceinline__ void fa_cta_sync2() { asm volatile("bar.sync 2, 256;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// Who merges a (b, kv head) row: every CTA of the row draws a ticket (early: the atomic's latency hides behind the
// CTA's own last piece of work); the CTA holding the LAST ticket merges.  The others publish their partial and then
// bump `done` (barrier, one thread fences for the CTA -- fences are cumulative -- then the atomic); the merger waits for
// done == n_parts - 1 (by then the others have normally long finished: it is the CTA whose V chunks crossed PCIe last)
// and needs neither a fence nor an atomic round trip of its own between its last arithmetic and the merge.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int fa_take_ticket(const FaParams &p, const int bh) { return atomicAdd(&p.tickets[bh], 1); }

__device__ __forceinline__ void fa_publish(const FaParams &p, const int bh) {        // all threads, partial already stored
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(&p.done[bh], 1);
  }
}

__device__ __forceinline__ void fa_wait_others(const FaParams &p, const int bh, const int n_parts) {   // all threads
  if (threadIdx.x == 0) {
    int seen;
    do {
      asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(p.done + bh) : "memory");
    } while (seen < n_parts - 1);
    p.done[bh] = 0;                                          // both counters are ready for the next launch
    p.tickets[bh] = 0;
  }
  __syncthreads();
}

// The last CTA of a row merges the row's partials (its own included: stored just before, same CTA, a barrier between).
// All 256 threads.  The partial sums are spread over all threads (with G <= 4 two thread groups take alternate partials)
// and requested sixteen at a time together with the (m, l) pairs the weights come from: one L2 round trip for up to 32
// partials.  s_f: [kFaMaxParts][8] floats, s_L: [8] floats, s_acc: [8][32] float4 of free shared memory.
__device__ __forceinline__ void fa_merge(const FaParams &p, const int bh, const int n_parts, float *s_f, float *s_L,
                                         float4 *s_acc) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gpad = p.G <= 4 ? 4 : 8, groups = 8 / gpad;      // thread = (partial group, head, 4 dims)
  const int grp = warp / gpad, hq = warp - grp * gpad, c4 = lane;
  const bool act = hq < p.G;
  const float *rml = p.part_ml + size_t(bh) * kFaMaxParts * p.Gp * 2;
  const size_t stride = size_t(p.Gp) * kHeadDim;
  const float *col = p.part_o + size_t(bh) * kFaMaxParts * stride + size_t(act ? hq : 0) * kHeadDim + c4 * 4;
  constexpr int kBatch = 16;
  float4 v[kBatch];
#pragma unroll
  for (int k = 0; k < kBatch; ++k) {
    const int idx = grp + groups * k;
    v[k] = (act && idx < n_parts) ? __ldcg(reinterpret_cast<const float4 *>(col + size_t(idx) * stride))
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (warp < p.G) {                                          // weight of every (partial, head): warp = head, lanes = partials
    float ms[4], ls[4], m = -INFINITY;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int idx = lane + 32 * k;
      const bool on = idx < n_parts;
      ms[k] = on ? __ldcg(rml + (size_t(idx) * p.Gp + warp) * 2) : -INFINITY;
      ls[k] = on ? __ldcg(rml + (size_t(idx) * p.Gp + warp) * 2 + 1) : 0.f;
      m = fmaxf(m, ms[k]);
    }
    m = warp_max(m);
    float lsum = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float f = (ms[k] == -INFINITY) ? 0.f : exp2f(ms[k] - m);
      lsum += ls[k] * f;
      s_f[(lane + 32 * k) * kFaMaxG + warp] = f;
    }
    lsum = warp_sum(lsum);
    if (lane == 0) s_L[warp] = lsum;
  }
  __syncthreads();
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (act) {
#pragma unroll
    for (int k = 0; k < kBatch; ++k) {
      const int idx = grp + groups * k;
      if (idx < n_parts) {
        const float f = s_f[idx * kFaMaxG + hq];
        acc.x = fmaf(v[k].x, f, acc.x); acc.y = fmaf(v[k].y, f, acc.y); acc.z = fmaf(v[k].z, f, acc.z); acc.w = fmaf(v[k].w, f, acc.w);
      }
    }
    for (int base = kBatch; grp + groups * base < n_parts; base += kBatch) {
#pragma unroll
      for (int k = 0; k < kBatch; ++k) {
        const int idx = grp + groups * (base + k);
        v[k] = idx < n_parts ? __ldcg(reinterpret_cast<const float4 *>(col + size_t(idx) * stride)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int k = 0; k < kBatch; ++k) {
        const int idx = grp + groups * (base + k);
        if (idx < n_parts) {
          const float f = s_f[idx * kFaMaxG + hq];
          acc.x = fmaf(v[k].x, f, acc.x); acc.y = fmaf(v[k].y, f, acc.y); acc.z = fmaf(v[k].z, f, acc.z); acc.w = fmaf(v[k].w, f, acc.w);
        }
      }
    }
  }
  if (groups == 2) {
    if (grp == 1 && act) s_acc[hq * 32 + c4] = acc;
    __syncthreads();
    if (grp == 0 && act) {
      const float4 x = s_acc[hq * 32 + c4];
      acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
    }
  }
  if (grp == 0 && act) {
    const float inv = 1.0f / s_L[hq];
    uint2 packed;
    packed.x = pack_bf16x2(acc.x * inv, acc.y * inv);
    packed.y = pack_bf16x2(acc.z * inv, acc.w * inv);
    const size_t at = (size_t(bh) * p.G + hq) * kHeadDim + c4 * 4;
    *reinterpret_cast<uint2 *>(p.out + at) = packed;
    // kv-head split: the row's output goes straight into every peer's gathered buffer over NVLink (P2P stores); no
    // collective launch on the chain
    for (int r = 0; r < p.xc.world; ++r)
      if (p.xc.peer_out[r]) *reinterpret_cast<uint2 *>(static_cast<skv_bf16 *>(p.xc.peer_out[r]) + at) = packed;
  }
  if (p.xc.world == 1 && !p.xc.peer_out[0]) {
    // One-GPU form (SKV_GEOM_WAIT_FLAGS_ONLY in the next scoring launch): the row's output is in `out`; raise the row's
    // word behind it.  Device scope is enough, the consumer runs on this GPU.
    __syncthreads();
    if (threadIdx.x == 0) {
      const int epoch = *p.xc.epoch_dev;
      __threadfence();
      asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p.xc.peer_flags[0] + bh), "r"(epoch) : "memory");
    }
  } else if (p.xc.world > 0) {
    // The row's flag word on every peer: the barrier makes the CTA's remote stores precede thread 0's fence (fences are
    // cumulative), one system-scope fence orders them in front of the flag stores.
    __syncthreads();
    if (threadIdx.x == 0) {
      const int epoch = *p.xc.epoch_dev;
      __threadfence_system();
      for (int r = 0; r < p.xc.world; ++r)
        asm volatile("st.relaxed.sys.global.s32 [%0], %1;" ::"l"(p.xc.peer_flags[r] + bh), "r"(epoch) : "memory");
    }
  }
}

// kv-head split: wait until every row word of the layer has reached the current epoch (the gathered attention output of
// the layer is then complete in THIS rank's buffer).
__global__ void __launch_bounds__(128) wait_flags_kernel(const int32_t *flags, int n, const int32_t *epoch_dev) {
  pdl_launch_dependents();
  pdl_wait();
  const int want = *epoch_dev;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    int seen;
    do {
      asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(seen) : "l"(flags + i) : "memory");
    } while (seen < want);
  }
}

// ------------------------------------------------------------------------------------------------
// Tile CTA: 16 entries of the miss list of row bh.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void fa_tile_role(const CUtensorMap *map_sv_ptr, const CUtensorMap *map_u_ptr, const FaParams &p,
                                             const int bh, const int tile, unsigned char *smem) {
  unsigned char *sA = smem;                                    // n_slabs x 16 KiB (U rows)
  unsigned char *sB = smem + p.n_slabs * kFaSlabBytes;         // n_slabs x 16 KiB (SV panel)
  // the operand area [0, tail_off) (tail_off >= 96 KiB) is reused once the MMA has retired: scratch at 0, V staging behind it
  unsigned char *tail = smem + p.tail_off;
  unsigned char *sQ = tail;                                    // q as the B operand of the q.K MMA: 2 slabs x [16 head
                                                               // rows (zero padded) x 64 dims] bf16, 128-byte swizzled
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(tail + 4096);
  uint64_t *mma_bar = full_bar + kFaMaxSlabs;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(mma_bar + 1);
  int *s_ids = reinterpret_cast<int *>(tmem_slot + 2);         // [16] chunk id (-1: entry beyond the miss list)
  int *s_slots = s_ids + kFaTileChunks;                        // [16] sparse slot
  float *s_red = reinterpret_cast<float *>(s_slots + kFaTileChunks);   // [4 warps][8 heads][2]
  int *s_flag = reinterpret_cast<int *>(s_red + 4 * kFaMaxG * 2);      // the CTA's ticket
  uint64_t *full_v = reinterpret_cast<uint64_t *>(s_flag + 2);         // [2]
  uint64_t *full_a = full_v + 2;                                       // [kFaMaxSlabs] U rows of the tile, per slab
  // scratch (valid after mma_bar): probabilities, P.V staging, merge weights -- 28.1 KiB at offset 0
  float *s_p = reinterpret_cast<float *>(smem);                // [128 rows][8 heads]           4 KiB
  float *s_o = s_p + kFaTileRows * kFaMaxG;                    // [4 row groups][8 heads][128] 16 KiB
  float *s_f = s_o + 4 * kFaMaxG * kHeadDim;                   // [128 parts][8]                4 KiB
  float *s_L = s_f + kFaMaxParts * kFaMaxG;                    // [8]
  float4 *s_acc = reinterpret_cast<float4 *>(s_L + 8);         // [8][32] float4                4 KiB
  // V staging, 2 KiB per chunk: chunks 0..3 in their own 8 KiB (written while the MMA still owns the operand area),
  // chunks 4..15 behind the scratch inside the operand area (valid only after mma_bar)
  unsigned char *vring = tail + 4096 + 1024;
  unsigned char *sV2 = smem + 32768;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = p.S, G = p.G;
  const int b = bh / p.H;
  const int tr = bh == 0 ? tile : -1;                           // SKV_TRACE builds: timeline row of this CTA
  if (tid == 0) SKV_STAMP_NS(tr, 0);

  if (tid == 0) {
    for (int s = 0; s < p.n_slabs; ++s) { mbarrier_init(&full_bar[s], 1); mbarrier_init(&full_a[s], kFaThreads / 32); }
    mbarrier_init(mma_bar, 1);
    mbarrier_init(&full_v[0], 1);
    mbarrier_init(&full_v[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(map_sv_ptr) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(map_u_ptr) : "memory");
    // The SV panel of the row is static state (written at prefill): it is requested ahead of the wait for the selection,
    // so the MMA finds it in shared memory.  A tile that turns out to be beyond the miss list waits for it before it
    // leaves (shared memory with a bulk copy in flight must not be handed to the next CTA).
    for (int s = 0; s < p.n_slabs; ++s) {
      mbarrier_expect_tx(&full_bar[s], kFaSlabBytes);
      tma_load_2d(sB + s * kFaSlabBytes, map_sv_ptr, s * 64, bh * kHeadDim, &full_bar[s]);
    }
  }
  if (warp == 0) tcgen05_alloc(tmem_slot, 128);
  // ---- everything below reads what the selection kernel (and, through it, its predecessors) wrote -------------------
  fa_wait_row(p, bh);
  int raw = 0;
  if (tid < kFaTileChunks) raw = __ldcg(p.miss_slots + size_t(bh) * S + tile * kFaTileChunks + tid);   // read before the
  const int cnt = __ldcg(p.cnts + bh);                                                          // hit count is known
  const int n_miss = S - cnt;
  const int n_active = (n_miss + kFaTileChunks - 1) / kFaTileChunks;
  if (tid < kFaTileChunks) {
    const int slot = raw & (S - 1);                             // S is a power of two; a no-op for real entries
    const int id = static_cast<int>(__ldcg(p.resident + size_t(bh) * S + slot));
    const bool on = tile * kFaTileChunks + tid < n_miss;
    s_ids[tid] = on ? id : -1;
    s_slots[tid] = on ? slot : 0;
  }
  {                                                             // one 16-byte piece of the q operand per thread
    const int slab = tid >> 7, hr = (tid >> 3) & 15, c = tid & 7;
    int4 v = make_int4(0, 0, 0, 0);
    if (hr < G) v = *reinterpret_cast<const int4 *>(p.q + (size_t(bh) * G + hr) * kHeadDim + slab * 64 + c * 8);
    *reinterpret_cast<int4 *>(sQ + slab * 2048 + (hr >> 3) * 1024 + (hr & 7) * 128 + ((c ^ (hr & 7)) << 4)) = v;
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tile >= n_active) {                                       // nothing of the miss list left for this tile
    if (warp == 0) tcgen05_dealloc(tmem_base, 128);
    if (tid == 0)
      for (int s = 0; s < p.n_slabs; ++s) mbarrier_wait(&full_bar[s], 0);      // the SV panel's copies have landed
    return;
  }
  if (tid == 0) SKV_STAMP_NS(tr, 1);
  const int n_parts = n_active + p.n_splits;

  // ================= operands of the tile's 128 x 128 x R product ======================================================
  // U rows of a chunk are 8 consecutive rows of the [B*L, R] matrix = exactly one 8-row x 64-column box per slab in the
  // 128-byte-swizzled K-major layout (1 KiB).  A TMA issue is a per-warp (uniform datapath) instruction, so boxes issued
  // by several lanes of ONE warp go out one after the other (~30 ns each: 48 boxes from one warp were 1.6 us of the
  // chain; 16-byte cp.async from all threads, the round-1 form, kept the LSU busy for ~2 us): every warp of the CTA
  // issues the boxes of two chunks and arrives on the slab barriers with its own byte count.  Columns beyond R are zero
  // filled by the TMA.
  {
    const int j = warp * 2 + (lane & 1);                        // lanes 0 / 1: the warp's two chunks
    const int id = lane < 2 ? s_ids[j] : -1;
    const unsigned live = __ballot_sync(0xffffffffu, id >= 0);
    if (lane < p.n_slabs) mbarrier_expect_tx(&full_a[lane], __popc(live) * 1024);
    __syncwarp();
    if (id >= 0) {
      const int row0 = b * p.L + id * kChunkTokens;
      for (int s = 0; s < p.n_slabs; ++s)
        tma_load_2d(sA + s * kFaSlabBytes + j * 1024, map_u_ptr, s * 64, row0, &full_a[s]);
    }
  }
  if (tid < 128) {
    if (tid == 0) SKV_STAMP_NS(tr, 6);
    if (tid == 0) {
      SKV_STAMP_NS(tr, 7);
      int step = 0;
      constexpr uint32_t idesc = umma_idesc_bf16_m128(128);
      for (int s = 0; s < p.n_slabs; ++s) {
        mbarrier_wait(&full_bar[s], 0);
        mbarrier_wait(&full_a[s], 0);
        tcgen05_fence_after();
        const uint64_t da = umma_desc_k_sw128(smem_addr(sA + s * kFaSlabBytes));
        const uint64_t db = umma_desc_k_sw128(smem_addr(sB + s * kFaSlabBytes));
        for (int kk = 0; kk < 4 && step < p.k_steps; ++kk, ++step)
          tcgen05_mma_bf16(tmem_base, da + uint64_t(kk * 2), db + uint64_t(kk * 2), idesc, step > 0);
      }
      tcgen05_commit(mma_bar);
    }
    __syncwarp();
  } else if (warp == 4) {
    // ================= warp 4: V chunks of the tile's 16 entries, by TMA bulk copies ===================================
    // Bulk copies, not LSU loads: completions of LSU loads leave an SM's L1 in issue order, so a PCIe read that stays
    // outstanding for ~30 us would hold back every later load of the SM (the U gather and the RoPE table reads of this
    // CTA, and whatever a split CTA sharing the SM reads).  The TMA unit tracks its requests on its own.
    // Four chunks go to a dedicated 8 KiB staging at once; the other twelve follow into the operand area as soon as the
    // MMA has retired (a few us later; the link is busy with the first wave of every tile CTA until then).
  }
  // source of every V chunk (lane j of warp 4 <-> entry j)
  const unsigned char *v_src_ptr = nullptr;
  bool v_on = false;
  if (warp == 4) {
    const int j = lane;
    const int id = j < kFaTileChunks ? s_ids[j] : -1;
    v_on = id >= 0;
    v_src_ptr = reinterpret_cast<const unsigned char *>(p.v_src + size_t(bh) * p.v_src_stride_bh) + size_t(v_on ? id : 0) * kChunkBytes;
    if (p.vs.tags && v_on) {                                    // HBM chunk store: read-only here (skv_vstore_fill updates it)
      const size_t w = size_t(bh) * p.vs.ways + (id & (p.vs.ways - 1));
      if (p.vs.tags[w] == id) v_src_ptr = reinterpret_cast<const unsigned char *>(p.vs.data) + w * kChunkBytes;
      else if (p.vs.pcie_count) atomicAdd(p.vs.pcie_count + bh, 1);
    }
    const unsigned m1 = __ballot_sync(0xffffffffu, v_on && j < kFaVRingChunks);
    if (lane == 0) mbarrier_expect_tx(&full_v[0], __popc(m1) * kChunkBytes);
    __syncwarp();
    if (v_on && j < kFaVRingChunks) bulk_load(vring + j * kChunkBytes, v_src_ptr, kChunkBytes, &full_v[0]);
    if (lane == 0) SKV_STAMP_NS(tr, 2);
  }

  // ---- epilogue, all 8 warps: thread = (tile row, column half).  Warp w reads TMEM lanes 32 (w % 4) ..; warps 0-3 take
  // columns [0,32) + [64,96), warps 4-7 columns [32,64) + [96,128) (the two halves of every neox RoPE pair set).  The
  // RoPE table words are requested BEFORE the wait for the MMA, so their L2 latency is off the chain. ----------------
  const int half = tid >> 7, row = tid & 127;
  const int j = row >> 3, r = row & 7;
  const int id = s_ids[j];
  const int slot = s_slots[j];
  const bool write = id >= 0;
  const long long token = (long long)(write ? id : 0) * kChunkTokens + r;
  uint32_t cw[16], sw[16];                                      // GLM uses the first 8 of each
  if (p.rope == 1) {
    const __nv_bfloat16 *cs = p.cos_sin + token * kHeadDim;
    const uint4 *c4 = reinterpret_cast<const uint4 *>(cs + half * 32);
    const uint4 *s4 = reinterpret_cast<const uint4 *>(cs + 64 + half * 32);
#pragma unroll
    for (int qd = 0; qd < 4; ++qd) {
      const uint4 a = __ldg(c4 + qd), bq = __ldg(s4 + qd);
      cw[4 * qd] = a.x; cw[4 * qd + 1] = a.y; cw[4 * qd + 2] = a.z; cw[4 * qd + 3] = a.w;
      sw[4 * qd] = bq.x; sw[4 * qd + 1] = bq.y; sw[4 * qd + 2] = bq.z; sw[4 * qd + 3] = bq.w;
    }
  } else {
    const __nv_bfloat16 *cs = p.cos_sin + token * (kHeadDim / 2);
    const uint4 *c4 = reinterpret_cast<const uint4 *>(cs + half * 16);
    const uint4 *s4 = reinterpret_cast<const uint4 *>(cs + 32 + half * 16);
#pragma unroll
    for (int qd = 0; qd < 2; ++qd) {
      const uint4 a = __ldg(c4 + qd), bq = __ldg(s4 + qd);
      cw[4 * qd] = a.x; cw[4 * qd + 1] = a.y; cw[4 * qd + 2] = a.z; cw[4 * qd + 3] = a.w;
      sw[4 * qd] = bq.x; sw[4 * qd + 1] = bq.y; sw[4 * qd + 2] = bq.z; sw[4 * qd + 3] = bq.w;
    }
#pragma unroll
    for (int e = 8; e < 16; ++e) { cw[e] = 0u; sw[e] = 0u; }
  }
  mbarrier_wait(mma_bar, 0);
  tcgen05_fence_after();
  if (tid == 0) SKV_STAMP_NS(tr, 3);
  if (warp == 4) {                                              // the operand area is free: the other twelve V chunks
    const unsigned m2 = __ballot_sync(0xffffffffu, v_on && lane >= kFaVRingChunks);
    if (lane == 0) mbarrier_expect_tx(&full_v[1], __popc(m2) * kChunkBytes);
    __syncwarp();
    if (v_on && lane >= kFaVRingChunks) bulk_load(sV2 + (lane - kFaVRingChunks) * kChunkBytes, v_src_ptr, kChunkBytes, &full_v[1]);
  }
  // The post-RoPE rows go to their slots in the K buffer AND, as bf16 in the swizzled K-major layout, into the operand
  // area (free since the rebuild MMA retired) as the A operand of a second tcgen05.mma: S[key][head] = K_tile . q^T
  // (M = 128 keys, N = 16 head columns, K = 128 dims) into columns 0..15 of the SAME accumulator, which every warp has
  // read by then.  Thread = key row reads its G dot products back with one tcgen05.ld: no SIMT dot products (they were
  // 1 100 instructions per thread, ~2 us of the chain).
  unsigned char *sK = smem + 65536;                             // 2 slabs x 16 KiB; [64 KiB, 96 KiB) of the operand area
  {
    const uint32_t taddr = tmem_base + (uint32_t((warp & 3) * 32) << 16);
    __nv_bfloat16 *orow = p.k_buf + size_t(bh) * p.buf_stride_bh + (size_t(p.sparse_start) + size_t(slot) * kChunkTokens + r) * kHeadDim;
    uint32_t lo[32], hi[32];
    tcgen05_ld32(taddr + half * 32, lo);              // columns [32*half, 32*half+32)
    tcgen05_ld32(taddr + 64 + half * 32, hi);         // columns [64+32*half, ...)
    tcgen05_wait_ld();
    tcgen05_fence_before();
    uint32_t olo[16], ohi[16];
    if (p.rope == 1) {
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const __nv_bfloat162 x1 = __floats2bfloat162_rn(__uint_as_float(lo[2 * e]), __uint_as_float(lo[2 * e + 1]));
        const __nv_bfloat162 x2 = __floats2bfloat162_rn(__uint_as_float(hi[2 * e]), __uint_as_float(hi[2 * e + 1]));
        const __nv_bfloat162 c = *reinterpret_cast<const __nv_bfloat162 *>(&cw[e]);
        const __nv_bfloat162 sn = *reinterpret_cast<const __nv_bfloat162 *>(&sw[e]);
        const __nv_bfloat162 o1 = __hfma2(x1, c, __hmul2(__hneg2(x2), sn));
        const __nv_bfloat162 o2 = __hfma2(x2, c, __hmul2(x1, sn));
        olo[e] = *reinterpret_cast<const uint32_t *>(&o1);
        ohi[e] = *reinterpret_cast<const uint32_t *>(&o2);
      }
    } else {
      // GLM: interleaved pairs over dims [0,64) with [cos32|sin32]; dims [64,128) copied.
#pragma unroll
      for (int e = 0; e < 8; ++e) {         // two pairs per iteration: elements 4e..4e+3
        const __nv_bfloat162 xe = __floats2bfloat162_rn(__uint_as_float(lo[4 * e]), __uint_as_float(lo[4 * e + 2]));
        const __nv_bfloat162 xo = __floats2bfloat162_rn(__uint_as_float(lo[4 * e + 1]), __uint_as_float(lo[4 * e + 3]));
        const __nv_bfloat162 c = *reinterpret_cast<const __nv_bfloat162 *>(&cw[e]);
        const __nv_bfloat162 sn = *reinterpret_cast<const __nv_bfloat162 *>(&sw[e]);
        const __nv_bfloat162 oe = __hfma2(xe, c, __hmul2(__hneg2(xo), sn));
        const __nv_bfloat162 oo = __hfma2(xo, c, __hmul2(xe, sn));
        olo[2 * e] = pack_bf16(__low2bfloat16(oe), __low2bfloat16(oo));
        olo[2 * e + 1] = pack_bf16(__high2bfloat16(oe), __high2bfloat16(oo));
      }
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const __nv_bfloat162 x2 = __floats2bfloat162_rn(__uint_as_float(hi[2 * e]), __uint_as_float(hi[2 * e + 1]));
        ohi[e] = *reinterpret_cast<const uint32_t *>(&x2);
      }
    }
    if (write) {
      uint4 *d1 = reinterpret_cast<uint4 *>(orow + half * 32);
      uint4 *d2 = reinterpret_cast<uint4 *>(orow + 64 + half * 32);
#pragma unroll
      for (int qd = 0; qd < 4; ++qd) {
        d1[qd] = make_uint4(olo[4 * qd], olo[4 * qd + 1], olo[4 * qd + 2], olo[4 * qd + 3]);
        d2[qd] = make_uint4(ohi[4 * qd], ohi[4 * qd + 1], ohi[4 * qd + 2], ohi[4 * qd + 3]);
      }
    }
    // A operand: dims [32 half, +32) are pieces 4 half .. 4 half + 3 of slab 0, dims [64 + 32 half, +32) the same pieces
    // of slab 1.  Rows beyond the miss list hold whatever the accumulator held (never NaN-propagating across rows: an
    // MMA row only feeds its own output row, which is masked below).
    unsigned char *krow = sK + (row >> 3) * 1024 + (row & 7) * 128;
#pragma unroll
    for (int qd = 0; qd < 4; ++qd) {
      const int piece = ((half * 4 + qd) ^ (row & 7)) << 4;
      *reinterpret_cast<uint4 *>(krow + piece) = make_uint4(olo[4 * qd], olo[4 * qd + 1], olo[4 * qd + 2], olo[4 * qd + 3]);
      *reinterpret_cast<uint4 *>(krow + kFaSlabBytes + piece) = make_uint4(ohi[4 * qd], ohi[4 * qd + 1], ohi[4 * qd + 2], ohi[4 * qd + 3]);
    }
  }
  fence_proxy_async_smem();                                     // generic-proxy writes (K rows, q) -> the MMA's async proxy
  fa_cta_sync2();                                               // every warp has read the accumulator and staged its rows
  if (tid == 0) {
    SKV_STAMP_NS(tr, 12);
    tcgen05_fence_after();
    constexpr uint32_t idesc_s = umma_idesc_bf16_m128(16);
#pragma unroll
    for (int sl = 0; sl < 2; ++sl) {
      const uint64_t da = umma_desc_k_sw128(smem_addr(sK + sl * kFaSlabBytes));
      const uint64_t db = umma_desc_k_sw128(smem_addr(sQ + sl * 2048));
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        tcgen05_mma_bf16(tmem_base, da + uint64_t(kk * 2), db + uint64_t(kk * 2), idesc_s, (sl | kk) > 0);
    }
    tcgen05_commit(mma_bar);                                    // second phase of the same barrier
  }
  if (half == 0) {
    mbarrier_wait(mma_bar, 1);
    tcgen05_fence_after();
    uint32_t sd[16];
    tcgen05_ld16(tmem_base + (uint32_t((warp & 3) * 32) << 16), sd);
    tcgen05_wait_ld();
    tcgen05_fence_before();
    float dot[kFaMaxG];
#pragma unroll
    for (int g = 0; g < kFaMaxG; ++g) dot[g] = __uint_as_float(sd[g]) * p.scale_log2;
    float t_dot[kFaMaxG];                                       // scaled q.k of this thread's row (log2 domain)
#pragma unroll
    for (int g = 0; g < kFaMaxG; ++g) t_dot[g] = (write && g < G) ? dot[g] : -INFINITY;
    // softmax partial over the tile's 128 rows: per-head maximum ...
#pragma unroll
    for (int g = 0; g < kFaMaxG; ++g) {
      if (g < G) {
        const float m = warp_max(t_dot[g]);
        if (lane == 0) s_red[(warp * kFaMaxG + g) * 2] = m;
      }
    }
    fa_tile_sync();                                             // also: all four warps are past their TMEM loads
    if (warp == 0) tcgen05_dealloc(tmem_base, 128);
    float l_part[kFaMaxG];
#pragma unroll
    for (int g = 0; g < kFaMaxG; ++g) {
      float pr = 0.f;
      if (g < G) {
        const float m = fmaxf(fmaxf(s_red[(0 * kFaMaxG + g) * 2], s_red[(1 * kFaMaxG + g) * 2]),
                              fmaxf(s_red[(2 * kFaMaxG + g) * 2], s_red[(3 * kFaMaxG + g) * 2]));
        pr = (t_dot[g] == -INFINITY) ? 0.f : exp2f(t_dot[g] - m);     // a tile has at least one real row: m is finite
      }
      s_p[row * kFaMaxG + g] = pr;
      l_part[g] = pr;
    }
    // ... and the sum of the probabilities
#pragma unroll
    for (int g = 0; g < kFaMaxG; ++g) {
      if (g < G) {
        const float l = warp_sum(l_part[g]);
        if (lane == 0) s_red[(warp * kFaMaxG + g) * 2 + 1] = l;
      }
    }
    if (tid == 0) SKV_STAMP_NS(tr, 4);
  }
  // ---- the V rows: everybody waits for both waves of bulk copies; thread 0 draws the CTA's ticket on the way ---------
  mbarrier_wait(&full_v[0], 0);
  if (tid == 128) SKV_STAMP_NS(tr, 5);
  mbarrier_wait(&full_v[1], 0);
  if (tid == 128) SKV_STAMP_NS(tr, 8);
  int ticket = 0;                                               // thread 0: drawn now, needed only behind the P.V below,
  if (tid == 0) ticket = fa_take_ticket(p, bh);                 // so the atomic's round trip is off the chain
  // the fetched chunks go to their sparse slots (fire and forget; they stay resident for the next tokens)
  {
    int4 *dbase = reinterpret_cast<int4 *>(p.v_buf + size_t(bh) * p.buf_stride_bh + size_t(p.sparse_start) * kHeadDim);
#pragma unroll
    for (int i = 0; i < kFaTileChunks * (kChunkBytes / 16) / kFaThreads; ++i) {
      const int piece = tid + kFaThreads * i;
      const int ch = piece >> 7, off = piece & 127;
      const int4 *sv = reinterpret_cast<const int4 *>(ch < kFaVRingChunks ? vring + ch * kChunkBytes
                                                                          : sV2 + (ch - kFaVRingChunks) * kChunkBytes);
      if (s_ids[ch] >= 0) dbase[size_t(s_slots[ch]) * (kChunkBytes / 16) + off] = sv[off];
    }
  }
  __syncthreads();                                              // probabilities (s_p, s_red) are complete
  if (tid == 0) SKV_STAMP_NS(tr, 9);

  // ================= all 8 warps: O[g][:] = sum_rows p[row][g] * V[row][:] ==========================================
  {
    const int c = tid & 63, rg = tid >> 6;                      // dims (2c, 2c+1); rows [32 rg, 32 rg + 32)
    float acc[kFaMaxG][2];
#pragma unroll
    for (int g = 0; g < kFaMaxG; ++g) { acc[g][0] = 0.f; acc[g][1] = 0.f; }
#pragma unroll 1
    for (int ch = 0; ch < 4; ++ch) {
      const int jj = rg * 4 + ch;
      if (s_ids[jj] < 0) continue;                              // entry beyond the miss list: its staging is stale
      const uint32_t *sv32 = reinterpret_cast<const uint32_t *>(jj < kFaVRingChunks ? vring + jj * kChunkBytes
                                                                                    : sV2 + (jj - kFaVRingChunks) * kChunkBytes);
#pragma unroll
      for (int rr = 0; rr < 8; ++rr) {
        const int prow = jj * 8 + rr;
        const uint32_t w = sv32[rr * 64 + c];
        const float v0 = bf16_lo(w), v1 = bf16_hi(w);
        const float4 pa = *reinterpret_cast<const float4 *>(s_p + prow * kFaMaxG);
        acc[0][0] = fmaf(pa.x, v0, acc[0][0]); acc[0][1] = fmaf(pa.x, v1, acc[0][1]);
        acc[1][0] = fmaf(pa.y, v0, acc[1][0]); acc[1][1] = fmaf(pa.y, v1, acc[1][1]);
        acc[2][0] = fmaf(pa.z, v0, acc[2][0]); acc[2][1] = fmaf(pa.z, v1, acc[2][1]);
        acc[3][0] = fmaf(pa.w, v0, acc[3][0]); acc[3][1] = fmaf(pa.w, v1, acc[3][1]);
        if (G > 4) {
          const float4 pb = *reinterpret_cast<const float4 *>(s_p + prow * kFaMaxG + 4);
          acc[4][0] = fmaf(pb.x, v0, acc[4][0]); acc[4][1] = fmaf(pb.x, v1, acc[4][1]);
          acc[5][0] = fmaf(pb.y, v0, acc[5][0]); acc[5][1] = fmaf(pb.y, v1, acc[5][1]);
          acc[6][0] = fmaf(pb.z, v0, acc[6][0]); acc[6][1] = fmaf(pb.z, v1, acc[6][1]);
          acc[7][0] = fmaf(pb.w, v0, acc[7][0]); acc[7][1] = fmaf(pb.w, v1, acc[7][1]);
        }
      }
    }
#pragma unroll
    for (int g = 0; g < kFaMaxG; ++g)
      if (g < G) *reinterpret_cast<float2 *>(s_o + (rg * kFaMaxG + g) * kHeadDim + 2 * c) = make_float2(acc[g][0], acc[g][1]);
  }
  if (tid == 0) *s_flag = ticket;
  __syncthreads();
  {
    const int hq = tid >> 5, c4 = tid & 31;
    float *po = p.part_o + (size_t(bh) * kFaMaxParts + tile) * p.Gp * kHeadDim;
    float *pml = p.part_ml + (size_t(bh) * kFaMaxParts + tile) * p.Gp * 2;
    if (hq < G) {
      float4 a = *reinterpret_cast<const float4 *>(s_o + (0 * kFaMaxG + hq) * kHeadDim + 4 * c4);
#pragma unroll
      for (int rg = 1; rg < 4; ++rg) {
        const float4 x = *reinterpret_cast<const float4 *>(s_o + (rg * kFaMaxG + hq) * kHeadDim + 4 * c4);
        a.x += x.x; a.y += x.y; a.z += x.z; a.w += x.w;
      }
      __stcg(reinterpret_cast<float4 *>(po + hq * kHeadDim + 4 * c4), a);
      if (c4 == 0) {
        const float m = fmaxf(fmaxf(s_red[(0 * kFaMaxG + hq) * 2], s_red[(1 * kFaMaxG + hq) * 2]),
                              fmaxf(s_red[(2 * kFaMaxG + hq) * 2], s_red[(3 * kFaMaxG + hq) * 2]));
        const float l = s_red[(0 * kFaMaxG + hq) * 2 + 1] + s_red[(1 * kFaMaxG + hq) * 2 + 1] +
                        s_red[(2 * kFaMaxG + hq) * 2 + 1] + s_red[(3 * kFaMaxG + hq) * 2 + 1];
        __stcg(pml + hq * 2, m);
        __stcg(pml + hq * 2 + 1, l);
      }
    }
  }
  if (tid == 0) SKV_STAMP_NS(tr, 10);
  if (*s_flag != n_parts - 1) {                                 // CTA uniform (written before the barriers above)
    fa_publish(p, bh);
    if (tid == 0) SKV_STAMP_NS(tr, 13);
    return;
  }
  fa_wait_others(p, bh, n_parts);
  if (tid == 0) SKV_STAMP_NS(tr, 11);
  fa_merge(p, bh, n_parts, s_f, s_L, s_acc);
  if (tid == 0) { SKV_STAMP_NS(tr, 13); SKV_STAMP_NS(tr, 14); }
}

// mma.sync m16n8k16 with the query heads in rows 0..7 of the A operand only (G <= 8): rows 8..15 are zero, so a1 = a3 = 0
// and the accumulator rows 8..15 stay zero; they go to a small set of scratch registers instead of 32 live ones.
__device__ __forceinline__ void mma_bf16_16816_lo(float &d0, float &d1, float &z0, float &z1, uint32_t a0, uint32_t a2,
                                                  uint32_t b0, uint32_t b1) {
  const uint32_t zero = 0u;
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d0), "+f"(d1), "+f"(z0), "+f"(z1)
               : "r"(a0), "r"(zero), "r"(a2), "r"(zero), "r"(b0), "r"(b1));
}

// one 64-key tile of K and V (two 64-dim slabs each) into a ring stage; rows past the tensor's extent are zero filled
__device__ __forceinline__ void fa_request_tile(unsigned char *dst, uint64_t *bar, const CUtensorMap *map_k,
                                                const CUtensorMap *map_v, int row0, int h, int b) {
  mbarrier_expect_tx(bar, kFaRingTile);
  tma_load_4d(dst, map_k, 0, row0, h, b, bar);
  tma_load_4d(dst + kFaKvSlab, map_k, 64, row0, h, b, bar);
  tma_load_4d(dst + 2 * kFaKvSlab, map_v, 0, row0, h, b, bar);
  tma_load_4d(dst + 3 * kFaKvSlab, map_v, 64, row0, h, b, bar);
}

// ------------------------------------------------------------------------------------------------
// Split CTA: keys [split * keys_per_split, ...) of row bh that are resident when the selection finishes.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void fa_split_role(const CUtensorMap *map_k, const CUtensorMap *map_v, const FaParams &p,
                                              const int bh, const int split, unsigned char *smem) {
  unsigned char *ring = smem;                                                   // 3 x 32 KiB
  unsigned char *tail = smem + p.tail_off;
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(tail);
  uint64_t *empty_bar = full_bar + kFaStages;
  uint32_t *s_mask = reinterpret_cast<uint32_t *>(empty_bar + kFaStages);       // [32] one bit per sparse slot
  int *s_flag = reinterpret_cast<int *>(s_mask + 32);
  // scratch over the ring once every tile has been consumed
  float *s_o = reinterpret_cast<float *>(ring);                                  // [4 warps][8 heads][132]
  float *s_ml = s_o + 4 * kFaMaxG * kFaOStride;                                  // [4 warps][8 heads][2]
  float *s_f = s_ml + 4 * kFaMaxG * 2;                                           // [128][8]
  float *s_L = s_f + kFaMaxParts * kFaMaxG;                                      // [8]
  float4 *s_acc = reinterpret_cast<float4 *>(s_L + 8);                           // [8][32] float4

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = bh / p.H, h = bh - b * p.H;
  const int G = p.G, S = p.S;
  const int tr = bh == 0 ? 64 + split : -1;                     // SKV_TRACE builds: timeline row of this CTA
  if (tid == 0) SKV_STAMP_NS(tr, 0);

  if (tid == 0) {
    for (int s = 0; s < kFaStages; ++s) { mbarrier_init(&full_bar[s], 1); mbarrier_init(&empty_bar[s], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(map_k) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(map_v) : "memory");
  }
  if (tid < 32) s_mask[tid] = 0u;
  __syncthreads();

  const int kv_len = p.kv_len_dev ? min(p.kv_len, *p.kv_len_dev) : p.kv_len;   // generation counter: constant during a token
  const int k_begin = split * p.keys_per_split;
  const int k_end = min(kv_len, k_begin + p.keys_per_split);
  const int n_tiles = k_end > k_begin ? (k_end - k_begin + kFaTileKeys - 1) / kFaTileKeys : 0;
  const int ss = p.sparse_start, se = p.sparse_start + S * kChunkTokens;
  int p_stage = 0, p_ph = 0, p_next = 0;
#define FA_REQUEST(t)                                           /* producer thread: tile t of the split into the ring */ \
  {                                                                                                                      \
    fa_request_tile(ring + p_stage * kFaRingTile, &full_bar[p_stage], map_k, map_v, k_begin + (t) * kFaTileKeys, h, b);   \
    if (++p_stage == kFaStages) { p_stage = 0; p_ph ^= 1; }                                                              \
  }
  // Row hand-over: the first tiles are requested AHEAD of the wait for the selection, while it is still running and HBM
  // is idle -- as long as they lie below the generated rows.  Local window, outliers and sparse slots were written by
  // launches of earlier tokens; the only rows of them that change during this launch are the slots of the miss list
  // (rewritten by the tile CTAs), and those are masked below whatever was read (old or new rows: finite either way).
  // The generated region holds the row the scoring launch of THIS step appended: tiles reaching it wait for the word.
  if (p.row_flags && p.early_splits && warp == 4 && lane == 0) {
    for (; p_next < n_tiles && p_next < kFaStages && k_begin + (p_next + 1) * kFaTileKeys <= se; ++p_next) FA_REQUEST(p_next)
  }
  fa_wait_row(p, bh);

  const int cnt = __ldcg(p.cnts + bh);
  const int n_miss = S - cnt;
  const int n_active = (n_miss + kFaTileChunks - 1) / kFaTileChunks;

  // the producer gets the (other) first tiles moving, then everybody marks the slots of the miss list (their K / V rows
  // are being overwritten by the tile CTAs right now and belong to THEIR partials)
  if (warp == 4 && lane == 0) {
    // the K / V rows were written through the generic proxy (previous launches, the fused append); with the row
    // hand-over no kernel boundary lies in between, so order them in front of this thread's async-proxy reads
    if (p.row_flags) asm volatile("fence.proxy.async;" ::: "memory");
    for (; p_next < n_tiles && p_next < kFaStages; ++p_next) FA_REQUEST(p_next)
  }
  for (int i = tid; i < n_miss; i += kFaThreads) {
    const int slot = __ldcg(p.miss_slots + size_t(bh) * S + i) & (S - 1);
    atomicOr(&s_mask[slot >> 5], 1u << (slot & 31));
  }
  __syncthreads();

  if (tid == 0) SKV_STAMP_NS(tr, 1);
  auto live = [&](int key) -> bool {                            // key < k_end and not a slot of the miss list
    if (key >= k_end) return false;
    if (key < ss || key >= se) return true;
    const int slot = (key - ss) >> 3;
    return ((s_mask[slot >> 5] >> (slot & 31)) & 1u) == 0u;
  };

  if (warp == 4) {
    // ===== TMA producer (the first kFaStages tiles are already in flight) =====
    if (lane == 0) {
      for (int t = p_next; t < n_tiles; ++t) {
        mbarrier_wait(&empty_bar[p_stage], p_ph ^ 1);
        FA_REQUEST(t)
      }
    }
  } else if (warp < 4) {
    // ===== consumers: warp w owns keys [16w, 16w+16) of every tile; heads are rows 0..7 of the MMA =====
    const int g = lane >> 2, tig = lane & 3;
    uint32_t qa[8][2];                                          // A operand rows g (k lo / k hi); rows g+8 are zero
    {
      const skv_bf16 *q0 = p.q + (size_t(bh) * G) * kHeadDim;
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const int d0 = kk * 16 + tig * 2;
        qa[kk][0] = g < G ? *reinterpret_cast<const uint32_t *>(q0 + size_t(g) * kHeadDim + d0) : 0u;
        qa[kk][1] = g < G ? *reinterpret_cast<const uint32_t *>(q0 + size_t(g) * kHeadDim + d0 + 8) : 0u;
      }
    }
    float o[16][2];
#pragma unroll
    for (int j = 0; j < 16; ++j) { o[j][0] = 0.f; o[j][1] = 0.f; }
    float z[4][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};   // rows 8..15 of the accumulators: always zero
    float m_run = -INFINITY, l_run = 0.f;

    const int lm_m = lane >> 3, lm_r = lane & 7;
    const int k_key = warp * 16 + (lm_m >> 1) * 8 + lm_r;
    const int v_key = warp * 16 + (lm_m & 1) * 8 + lm_r;

    int stage = 0, ph = 0;
    for (int t = 0; t < n_tiles; ++t) {
      const int key0 = k_begin + t * kFaTileKeys + warp * 16;
      mbarrier_wait(&full_bar[stage], ph);
      const int kcol = key0 + tig * 2;
      const bool v00 = live(kcol), v01 = live(kcol + 1), v10 = live(kcol + 8), v11 = live(kcol + 9);
      const bool any = __any_sync(0xffffffffu, v00 || v01 || v10 || v11);
      if (any) {                                                // warp uniform
        const uint32_t kbase = smem_addr(ring + stage * kFaRingTile);
        const uint32_t vbase = kbase + 2 * kFaKvSlab;
        float s0[2] = {0.f, 0.f}, s1[2] = {0.f, 0.f};
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const int c = kk * 2 + (lm_m & 1);
          uint32_t kb[4];
          ldmatrix_x4(kb, kbase + (c >> 3) * kFaKvSlab + k_key * 128 + (((c & 7) ^ (k_key & 7)) << 4));
          mma_bf16_16816_lo(s0[0], s0[1], z[0][0], z[0][1], qa[kk][0], qa[kk][1], kb[0], kb[1]);
          mma_bf16_16816_lo(s1[0], s1[1], z[1][0], z[1][1], qa[kk][0], qa[kk][1], kb[2], kb[3]);
        }
        float tt[4] = {v00 ? s0[0] * p.scale_log2 : -INFINITY, v01 ? s0[1] * p.scale_log2 : -INFINITY,
                       v10 ? s1[0] * p.scale_log2 : -INFINITY, v11 ? s1[1] * p.scale_log2 : -INFINITY};
        float mx = fmaxf(fmaxf(tt[0], tt[1]), fmaxf(tt[2], tt[3]));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        const float mn = fmaxf(m_run, mx);                      // finite: the slice has a live key
        const float corr = exp2f(m_run - mn);                   // 0 on the first live slice
        m_run = mn;
        float pp[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) pp[e] = exp2f(tt[e] - mn);  // exp2(-inf) = 0 for masked keys
        l_run = l_run * corr + (pp[0] + pp[1]) + (pp[2] + pp[3]);
#pragma unroll
        for (int j = 0; j < 16; ++j) { o[j][0] *= corr; o[j][1] *= corr; }
        const uint32_t pa0 = pack_bf16x2(pp[0], pp[1]), pa2 = pack_bf16x2(pp[2], pp[3]);
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const int c = jj * 2 + (lm_m >> 1);
          uint32_t vb[4];
          ldmatrix_x4_trans(vb, vbase + (c >> 3) * kFaKvSlab + v_key * 128 + (((c & 7) ^ (v_key & 7)) << 4));
          mma_bf16_16816_lo(o[2 * jj][0], o[2 * jj][1], z[(2 * jj) & 3][0], z[(2 * jj) & 3][1], pa0, pa2, vb[0], vb[1]);
          mma_bf16_16816_lo(o[2 * jj + 1][0], o[2 * jj + 1][1], z[(2 * jj + 1) & 3][0], z[(2 * jj + 1) & 3][1], pa0, pa2,
                            vb[2], vb[3]);
        }
      }
      __syncwarp();
      if (lane == 0) mbarrier_arrive(&empty_bar[stage]);
      if (++stage == kFaStages) { stage = 0; ph ^= 1; }
    }
    l_run += __shfl_xor_sync(0xffffffffu, l_run, 1);
    l_run += __shfl_xor_sync(0xffffffffu, l_run, 2);
    if (tid == 0) { SKV_STAMP_NS(tr, 2); *s_flag = fa_take_ticket(p, bh); }
    // every warp of the CTA is past its last tile before the ring is reused as staging
    asm volatile("bar.sync 1, 128;" ::: "memory");
    float *wo = s_o + warp * kFaMaxG * kFaOStride;
#pragma unroll
    for (int j = 0; j < 16; ++j) *reinterpret_cast<float2 *>(wo + g * kFaOStride + j * 8 + tig * 2) = make_float2(o[j][0], o[j][1]);
    if (tig == 0) { s_ml[(warp * kFaMaxG + g) * 2] = m_run; s_ml[(warp * kFaMaxG + g) * 2 + 1] = l_run; }
  }
  __syncthreads();
  // ---- merge the four consumer warps: thread = (head, 4 dims) ------------------------------------------------------
  {
    const int hq = tid >> 5, c4 = tid & 31;
    const int part = n_active + split;
    float *po = p.part_o + (size_t(bh) * kFaMaxParts + part) * p.Gp * kHeadDim;
    float *pml = p.part_ml + (size_t(bh) * kFaMaxParts + part) * p.Gp * 2;
    if (hq < G) {
      float m = -INFINITY;
#pragma unroll
      for (int w = 0; w < 4; ++w) m = fmaxf(m, s_ml[(w * kFaMaxG + hq) * 2]);
      float l = 0.f;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const float mw = s_ml[(w * kFaMaxG + hq) * 2];
        const float f = (mw == -INFINITY) ? 0.f : exp2f(mw - m);
        l += s_ml[(w * kFaMaxG + hq) * 2 + 1] * f;
        const float4 x = *reinterpret_cast<const float4 *>(s_o + (w * kFaMaxG + hq) * kFaOStride + 4 * c4);
        acc.x = fmaf(x.x, f, acc.x); acc.y = fmaf(x.y, f, acc.y); acc.z = fmaf(x.z, f, acc.z); acc.w = fmaf(x.w, f, acc.w);
      }
      __stcg(reinterpret_cast<float4 *>(po + hq * kHeadDim + 4 * c4), acc);
      if (c4 == 0) { __stcg(pml + hq * 2, m); __stcg(pml + hq * 2 + 1, l); }
    }
  }
  if (tid == 0) SKV_STAMP_NS(tr, 3);
  const int n_parts = n_active + p.n_splits;
  if (*s_flag != n_parts - 1) {                                 // CTA uniform: the ticket was drawn before the barrier above
    fa_publish(p, bh);
    if (tid == 0) SKV_STAMP_NS(tr, 5);
    return;
  }
  fa_wait_others(p, bh, n_parts);
  if (tid == 0) SKV_STAMP_NS(tr, 4);
  fa_merge(p, bh, n_parts, s_f, s_L, s_acc);
  if (tid == 0) { SKV_STAMP_NS(tr, 5); SKV_STAMP_NS(tr, 14); }
}

#undef FA_REQUEST

__global__ void __launch_bounds__(kFaThreads, 2)
    fetch_rebuild_attend_kernel(const __grid_constant__ CUtensorMap map_sv, const __grid_constant__ CUtensorMap map_u,
                                const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v,
                                FaParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  pdl_launch_dependents();
  const int n_tile_ctas = p.BH * p.n_tiles;
  int x = int(blockIdx.x);
  if (p.order == 2) {                              // [tiles 0 .. n/2) of all rows | splits | tiles n/2 .. n of all rows]
    const int first = p.BH * (p.n_tiles / 2), n_split_ctas = p.BH * p.n_splits;
    if (x >= first) x = x < first + n_split_ctas ? n_tile_ctas + (x - first) : x - n_split_ctas;
  }
  if (x < n_tile_ctas) {
    // tile CTAs come first in the grid: their PCIe reads are the long pole and should be in flight as early as possible
    if (p.order == 0) fa_tile_role(&map_sv, &map_u, p, x / p.n_tiles, x % p.n_tiles, smem);
    else fa_tile_role(&map_sv, &map_u, p, x % p.BH, x / p.BH, smem);     // the tiles past a row's miss list come last
  } else {
    const int id = x - n_tile_ctas;
    fa_split_role(&map_k, &map_v, p, id / p.n_splits, id % p.n_splits, smem);
  }
  if (p.row_flags) pdl_wait();   // see fa_wait_row: by now the selection grid has long completed
}

int make_tensor_map_2d(CUtensorMap *out, const void *base, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_cols);

static int64_t fa_align(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

}  // namespace skv

using namespace skv;

SKV_TRACE_READER(fused_attend)

extern "C" {

int64_t skv_attend_workspace_bytes(int batch, int kv_heads, int q_groups) {
  const int64_t rows = int64_t(batch) * kv_heads, Gp = (q_groups + 3) / 4 * 4;
  return fa_align(rows * kFaMaxParts * Gp * kHeadDim * 4, 256) + fa_align(rows * kFaMaxParts * Gp * 2 * 4, 256) +
         fa_align(rows * 2 * 4, 256);
}

int skv_fetch_rebuild_attend(const skv_geom *g, const skv_bf16 *v_store, int64_t store_stride, skv_bf16 *v_buf,
                             const skv_bf16 *U, const skv_bf16 *SV, const skv_bf16 *cos_sin, skv_bf16 *k_buf,
                             const int64_t *resident, const int32_t *cnts, const int32_t *offsets, const skv_vstore *vstore,
                             const skv_bf16 *q, skv_bf16 *out, int kv_len, const int32_t *kv_len_dev, float scale,
                             void *workspace, const skv_exchange *exchange, skv_stream_t stream) {
  if (!g || !v_store || !v_buf || !U || !SV || !cos_sin || !k_buf || !resident || !cnts || !offsets || !q || !out || !workspace) {
    set_error("fetch_rebuild_attend: null pointer"); return SKV_ERR_ARGUMENT;
  }
  if (vstore && (!vstore->tags || !vstore->data || vstore->ways < 1 || (vstore->ways & (vstore->ways - 1)))) {
    set_error("fetch_rebuild_attend: chunk store needs tags, data and a power-of-two way count"); return SKV_ERR_ARGUMENT;
  }
  if (g->head_dim != kHeadDim || g->chunk != kChunkTokens || !valid_map_size(g->select) || g->table_mode != 1 ||
      g->q_groups < 1 || g->q_groups > kFaMaxG || g->rank % 8 || g->rank < 8 || g->rank > kFaMaxSlabs * 64 ||
      kv_len < 1 || kv_len > g->buf_len || g->sparse_start + g->select * kChunkTokens > g->buf_len) {
    set_error("fetch_rebuild_attend: needs the in-place chunk table, head_dim 128, chunk 8, G <= 8, rank <= 256 (multiple of 8) "
              "and 1 <= kv_len <= buf_len (got table_mode %d G %d rank %d kv_len %d)", g->table_mode, g->q_groups, g->rank, kv_len);
    return SKV_ERR_UNSUPPORTED;
  }
  const int B = g->batch, H = g->kv_heads, BH = B * H, S = g->select;
  FaParams p{};
  p.resident = resident; p.cnts = cnts; p.miss_slots = offsets;
  p.U = reinterpret_cast<const __nv_bfloat16 *>(U);
  p.cos_sin = reinterpret_cast<const __nv_bfloat16 *>(cos_sin);
  p.k_buf = reinterpret_cast<__nv_bfloat16 *>(k_buf);
  p.buf_stride_bh = int64_t(g->buf_len) * kHeadDim;
  p.sparse_start = g->sparse_start;
  p.H = H; p.S = S; p.L = g->seq_len; p.R = g->rank;
  p.n_slabs = (g->rank + 63) / 64;
  p.k_steps = (g->rank + 15) / 16;
  p.rope = g->rope_glm ? 2 : 1;
  p.v_src = v_store; p.v_src_stride_bh = store_stride; p.v_buf = v_buf;
  if (vstore) p.vs = *vstore;
  p.q = q; p.out = out; p.kv_len_dev = kv_len_dev; p.kv_len = kv_len;
  p.row_flags = g->row_flags;
  { const char *ee = getenv("SKV_FA_EARLY"); p.early_splits = (ee && *ee) ? atoi(ee) : 1; }   // A/B switch, read per call
  p.order = -1;
  { const char *oe = getenv("SKV_FA_ORDER"); if (oe && *oe) p.order = atoi(oe); }   // A/B switch, read per call
  if (exchange && exchange->world > 0) {
    if (exchange->world > SKV_MAX_PEERS || !exchange->epoch_dev) {
      set_error("fetch_rebuild_attend: exchange needs 1..%d peers and epoch_dev", SKV_MAX_PEERS); return SKV_ERR_ARGUMENT;
    }
    const bool local = exchange->world == 1 && !exchange->peer_out[0];   // one-GPU hand-over: the row's flag word only
    for (int r = 0; r < exchange->world; ++r)
      if ((!exchange->peer_out[r] && !local) || !exchange->peer_flags[r]) { set_error("fetch_rebuild_attend: null peer pointer"); return SKV_ERR_ARGUMENT; }
    p.xc = *exchange;
  }
  p.G = g->q_groups; p.Gp = (g->q_groups + 3) / 4 * 4;
  p.BH = BH; p.n_tiles = S / kFaTileChunks;
  p.scale_log2 = scale * 1.4426950408889634f;
  // split planning: two CTAs per SM; about half of the tile CTAs stay (a 50-60 % hit rate), the rest of the slots goes
  // to the splits.  Cost of a split count c: waves x (tiles per split + ~3 tiles of fixed prologue / merge) + merge of c
  const int64_t slots = 2 * kNumSMs, busy = int64_t(BH) * p.n_tiles / 2;
  const int64_t avail = slots - busy > kNumSMs ? slots - busy : kNumSMs;
  int ns = 1;
  {
    double best = 1e30;
    for (int c = 1; c <= 64; ++c) {
      const int keys = ((kv_len + c - 1) / c + 15) / 16 * 16;
      const int tiles = (keys + kFaTileKeys - 1) / kFaTileKeys;
      const int64_t waves = (int64_t(BH) * c + avail - 1) / avail;
      const double cost = double(waves) * (tiles + 3.0) + (c > 1 ? 1.0 + 0.05 * c : 0.0);
      if (cost < best - 1e-9) { best = cost; ns = c; }
    }
  }
  int kps = ((kv_len + ns - 1) / ns + 15) / 16 * 16;
  ns = (kv_len + kps - 1) / kps;
  p.keys_per_split = kps; p.n_splits = ns;
  const int64_t rows = BH;
  unsigned char *w = static_cast<unsigned char *>(workspace);
  p.part_o = reinterpret_cast<float *>(w);
  w += fa_align(rows * kFaMaxParts * p.Gp * kHeadDim * 4, 256);
  p.part_ml = reinterpret_cast<float *>(w);
  w += fa_align(rows * kFaMaxParts * p.Gp * 2 * 4, 256);
  p.tickets = reinterpret_cast<int *>(w);
  p.done = p.tickets + rows;
  const int tile_bytes = 2 * p.n_slabs * kFaSlabBytes, ring_bytes = kFaStages * kFaRingTile;
  p.tail_off = tile_bytes > ring_bytes ? tile_bytes : ring_bytes;
  const int smem = p.tail_off + kFaTailBytes;

  CUtensorMap map_sv, map_u, map_k, map_v;
  int rc = make_tensor_map_2d(&map_sv, SV, uint64_t(BH) * kHeadDim, g->rank, kHeadDim, 64);
  if (rc) return rc;
  rc = make_tensor_map_2d(&map_u, U, uint64_t(B) * g->seq_len, g->rank, kChunkTokens, 64);
  if (rc) return rc;
  const int64_t stride_row = kHeadDim, stride_h = int64_t(g->buf_len) * kHeadDim, stride_b = stride_h * H;
  rc = make_tensor_map_kv(&map_k, k_buf, g->buf_len, H, B, stride_row, stride_h, stride_b);
  if (rc) return rc;
  rc = make_tensor_map_kv(&map_v, v_buf, g->buf_len, H, B, stride_row, stride_h, stride_b);
  if (rc) return rc;
  if (int rc2 = ensure_dynamic_smem(reinterpret_cast<const void *>(fetch_rebuild_attend_kernel), smem, "fetch_rebuild_attend_kernel"))
    return rc2;
  const int grid = BH * (p.n_tiles + p.n_splits);
  // CTA order: with one wave (batch 1) a row's tile CTAs sit together; with several waves the tile indices that are
  // almost never past a row's miss list go first, the splits next and the (mostly idle) high tile indices last
  // (batch 8, V in HBM: 121.5 -> 119.4 us per layer-step; batch 1: 32.5 against 33.4 us the other way round)
  if (p.order < 0) p.order = grid > 2 * kNumSMs ? 2 : 0;
  cudaError_t e = launch_kernel(fetch_rebuild_attend_kernel, dim3(grid), dim3(kFaThreads), smem, as_stream(stream),
                                /*pdl=*/true, map_sv, map_u, map_k, map_v, p);
  if (e != cudaSuccess) { set_error("fetch_rebuild_attend_kernel: %s", cudaGetErrorString(e)); return int(e); }
  return launch_status("fetch_rebuild_attend_kernel");
}

int skv_wait_flags(const int32_t *flags, int n, const int32_t *epoch_dev, skv_stream_t stream) {
  if (!flags || !epoch_dev || n < 1 || n > 4096) { set_error("wait_flags: bad argument"); return SKV_ERR_ARGUMENT; }
  cudaError_t e = launch_kernel(wait_flags_kernel, dim3(1), dim3(128), 0, as_stream(stream), /*pdl=*/true, flags, n, epoch_dev);
  if (e != cudaSuccess) { set_error("wait_flags_kernel: %s", cudaGetErrorString(e)); return int(e); }
  return launch_status("wait_flags_kernel");
}

}  // extern "C"
