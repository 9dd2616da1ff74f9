// Chunk-cache bookkeeping shared by the stand-alone reorder entry point and the fused select kernel.
//
// Semantics follow the reference's reorder_keys_and_mixed_offsets (kernels/map.cuh:754-796 with
// block_sort2 :508-539): hits first ordered by old slot (offset = old slot), misses after ordered by
// chunk id (offset = chunk id), cnt = number of hits.  The implementation is not the reference's
// (no hash table, one sort instead of two): sort the selected ids, binary-search every resident id
// in them, then two ballot scans place hits and misses.  The result is identical whenever the
// selected ids are distinct (top-k guarantees it); a resident id present twice resolves to its
// lowest slot (the reference leaves that case to a race).
#pragma once
#include "common.cuh"

namespace skv {

struct ChunkTableSmem {
  int *sorted_new;   // [S]
  int *slot_of;      // [S]
  int *warp_tot;     // [33]
};

// All threads of the block must call this (it synchronises); threads with tid >= S carry no element.
// old_id / new_id: this thread's resident id and selected id (int32-truncated), valid for tid < S.
// Outputs through pointers to global memory for this (b,h) row.
template <int S, bool kInPlace = false>
__device__ __forceinline__ void chunk_table_update(int tid, int old_id, int new_id, ChunkTableSmem sm,
                                                   int64_t *__restrict__ g_resident,
                                                   int32_t *__restrict__ g_offsets,
                                                   int32_t *__restrict__ g_cnt) {
  const bool active = tid < S;
  if (active) {
    sm.sorted_new[tid] = new_id;
    sm.slot_of[tid] = INT_MAX;
  }
  __syncthreads();
  // The ids usually arrive ascending already (the landmark index of the reference is an ascending list,
  // kv_cache.py:909-915, and the selection is emitted in landmark order): one vote skips the sort.
  const int in_order = __syncthreads_and(!active || tid == 0 || sm.sorted_new[tid - 1] <= new_id);
  // bitonic sort, ascending
#pragma unroll 1
  for (int k = 2; k <= (in_order ? 0 : S); k <<= 1) {
#pragma unroll 1
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (active) {
        int ixj = tid ^ j;
        if (ixj > tid) {
          int a = sm.sorted_new[tid], b = sm.sorted_new[ixj];
          bool asc = (tid & k) == 0;
          if ((a > b) == asc) {
            sm.sorted_new[tid] = b;
            sm.sorted_new[ixj] = a;
          }
        }
      }
      __syncthreads();
    }
  }
  // resident side: where (if anywhere) does my id sit among the selected ids?
  int found = -1;
  if (active && old_id != -1) {
    int lo = 0, hi = S;
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (sm.sorted_new[mid] < old_id) lo = mid + 1; else hi = mid;
    }
    if (lo < S && sm.sorted_new[lo] == old_id) {
      found = lo;
      atomicMin(&sm.slot_of[lo], tid);
    }
  }
  __syncthreads();
  const int keep = (found >= 0 && sm.slot_of[found] == tid) ? 1 : 0;   // hit, kept at its lowest slot
  const int my_sorted = active ? sm.sorted_new[tid] : 0;
  const int miss = (active && sm.slot_of[tid] == INT_MAX) ? 1 : 0;
  // one scan for both flags (S <= 1024 fits 16 bits): low half = hits before me, high half = misses before me
  int packed_total;
  const int packed = block_excl_scan(keep | (miss << 16), sm.warp_tot, packed_total);
  const int hpos = packed & 0xFFFF, mpos = packed >> 16;
  const int cnt = packed_total & 0xFFFF;
  if (!kInPlace) {
    if (keep) {
      g_resident[hpos] = static_cast<int64_t>(old_id);
      g_offsets[hpos] = tid;
    }
    if (miss) {
      g_resident[cnt + mpos] = static_cast<int64_t>(my_sorted);
      g_offsets[cnt + mpos] = my_sorted;
    }
  } else {
    // In-place table (B200 path): hits stay in their slots, so neither K nor V moves for them; the i-th
    // missed chunk (ascending id) takes the i-th freed slot (ascending slot).  offsets[i] = that slot.
    const int freed = (active && !keep) ? 1 : 0;
    const int fpos = tid - hpos;           // freed slots before me = slots before me that are not kept hits
    __syncthreads();                       // slot_of is dead from here on: reuse it as the free-slot list
    if (freed) sm.slot_of[fpos] = tid;
    __syncthreads();
    if (miss) {
      const int slot = sm.slot_of[mpos];
      g_resident[slot] = static_cast<int64_t>(my_sorted);
      g_offsets[mpos] = slot;
    }
  }
  if (tid == 0) *g_cnt = cnt;
}

}  // namespace skv
