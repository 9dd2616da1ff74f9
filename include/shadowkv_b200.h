/*
 * shadowkv_b200.h -- C-ABI of the B200-native ShadowKV sparse-KV decode step.
 *
 * Every entry point takes plain device/host pointers, sizes and a cudaStream_t
 * (passed as void*; NULL = legacy default stream).  No torch types cross this
 * boundary.  All functions return 0 on success, a positive cudaError_t value if
 * the CUDA runtime reported an error at launch, or a negative SKV_ERR_* code
 * for an argument the kernels do not support.  Launches are asynchronous.
 *
 * Part 1 mirrors, one to one, the 12 functions the reference binds with
 * pybind11 (reference kernels/main.cu:45-80, signatures kernels/functions.h);
 * each comment cites the reference interface it replaces.  Tensors that the
 * reference receives as torch::Tensor arrive here as their data_ptr(); element
 * strides stay in elements, exactly as the reference's Python wrappers compute
 * them (models/tensor_op.py:154-238, models/kv_cache.py:1006-1150).
 *
 * Part 2 adds the fused entry points of the B200 path (scoring + selection +
 * chunk-table update in two launches, tcgen05 K rebuild with RoPE epilogue,
 * cluster-staged hit compaction, miss fetch from pinned host memory, split-K
 * decode attention).  ShadowKVCache_CPU in shadowkv_b200/kv_cache.py routes
 * its reference-named methods to these.
 */
#ifndef SHADOWKV_B200_H
#define SHADOWKV_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SKV_OK 0
#define SKV_ERR_UNSUPPORTED (-1) /* shape/size outside what the kernels handle */
#define SKV_ERR_ARGUMENT (-2)    /* null pointer / inconsistent sizes           */
#define SKV_ERR_DRIVER (-3)      /* driver entry point (tensor map) unavailable */

typedef void *skv_stream_t; /* cudaStream_t */
typedef uint16_t skv_bf16;  /* raw bfloat16 bits */

/* Library / build identification; returns a static string. */
const char *skv_version(void);
/* Last error message of the calling thread ("" if none). */
const char *skv_last_error(void);
/* Implementations active in this process (the product kernels, or a cross-check / A-B variant selected through an
 * SKV_* environment variable); a string owned by the library. */
const char *skv_active_config(void);

/* ======================================================================
 * Part 1 -- the reference's 12 extension entry points
 * ====================================================================== */

/* functions.h:71-79  gather_copy (kernels/gather_copy.cu:63-146, copy.cuh:481-508)
 * Uncached gather: for every (b,h) block and i < map_size,
 *   dst[block*gpu_v_length + i*1024 ..] <- values[block*cpu_v_length + position_ids[block*map_size+i]*1024 ..]
 * (1024 bf16 = one 2 KiB chunk).  values may be pinned host memory.
 * map_size must be 128, 256, 512 or 1024. */
int skv_gather_copy(const skv_bf16 *values, skv_bf16 *v_cache_buffer, const int64_t *position_ids,
                    int batch_size, int heads, int cpu_v_length, int gpu_v_length, int map_size,
                    skv_stream_t stream);

/* functions.h:97-106  gather_copy_d2d_with_offsets (gather_copy.cu:165-240, copy.cuh:649-687)
 * In-place forward compaction of cache-hit chunks of the K buffer:
 *   for i < cnts[block]:  slot i <- slot offsets[block*map_size + i]
 * slot s of block lives at keys + block*gpu_k_stride + gpu_k_offset + s*1024 (elements). */
int skv_gather_copy_d2d_with_offsets(skv_bf16 *keys, const int32_t *offsets, const int32_t *cnts,
                                     int batch_size, int heads, int gpu_k_length, int gpu_k_offset,
                                     int gpu_k_stride, int map_size, skv_stream_t stream);

/* functions.h:123-128  reorder_keys_and_compute_offsets (gather_copy.cu:259-307, map.cuh:754-796)
 * Per block: classify cur_pos_ids against the resident table cached_pos_ids;
 * hits first (ordered by old slot, offset = old slot), misses after (ordered by
 * chunk id, offset = chunk id); cached_pos_ids is rewritten IN PLACE,
 * cnts[block] = number of hits.  Ids are truncated to int32; -1 is reserved. */
int skv_reorder_keys_and_compute_offsets(int64_t *cached_pos_ids, const int64_t *cur_pos_ids,
                                         int32_t *offsets, int32_t *cnts, int batch_size, int heads,
                                         int map_size, skv_stream_t stream);

/* functions.h:151-158  gather_copy_with_offsets (gather_copy.cu:332-419, copy.cuh:785-846)
 * Cached V fetch: hits compacted in place inside v_cache_buffer, misses read
 * from the pinned host store `values`.  temp and signals are accepted for
 * signature compatibility and are not used (no inter-block handshake here). */
int skv_gather_copy_with_offsets(const skv_bf16 *values, skv_bf16 *v_cache_buffer, skv_bf16 *temp,
                                 const int32_t *offsets, const int32_t *cnts, int32_t *signals,
                                 int batch_size, int heads, int cpu_v_length, int gpu_v_length,
                                 int gpu_v_offset, int gpu_v_stride, int map_size, skv_stream_t stream);

/* functions.h:187-193  apply_rotary_pos_emb (kernels/rope.cu:61-107): separate cos / sin tables,
 * int64 token positions, neox half-split, bf16 arithmetic. */
int skv_apply_rotary_pos_emb(const skv_bf16 *x, const skv_bf16 *cos, const skv_bf16 *sin,
                             const int64_t *position_ids, skv_bf16 *output, int batch_size, int heads,
                             int seq_len, int embed_dim, int stride_xb, int stride_xh, int stride_xs,
                             int stride_xe, int stride_cos, int stride_sin, int stride_pid_b,
                             int stride_pid_h, int stride_pid_s, int half_dim, skv_stream_t stream);

/* functions.h:240-246  apply_rotary_pos_emb_new (rope_new.cu:55-126): combined [cos|sin] table. */
int skv_apply_rotary_pos_emb_new(const skv_bf16 *x, const skv_bf16 *cos_sin, const int64_t *position_ids,
                                 skv_bf16 *output, int batch_size, int heads, int seq_len, int embed_dim,
                                 int stride_xb, int stride_xh, int stride_xs, int stride_xe,
                                 int stride_cos_sin, int stride_pid_b, int stride_pid_h, int stride_pid_s,
                                 int half_dim, skv_stream_t stream);

/* functions.h:281-287  apply_rotary_pos_emb_new_v2 (rope_new.cu:135-205): int32 CHUNK ids,
 * token position = id*chunk_size + s%chunk_size. */
int skv_apply_rotary_pos_emb_new_v2(const skv_bf16 *x, const skv_bf16 *cos_sin, const int32_t *position_ids,
                                    skv_bf16 *output, int batch_size, int heads, int seq_len, int embed_dim,
                                    int stride_xb, int stride_xh, int stride_xs, int stride_xe,
                                    int stride_cos_sin, int stride_pid_b, int stride_pid_h,
                                    int stride_pid_s, int half_dim, int chunk_size, skv_stream_t stream);

/* functions.h:328-337 / 362-371 / 396-405  apply_rotary_pos_emb_push_cache{,_opt,_opt_glm}
 * (rope_new.cu:219-300, 321-411, 429-534): RoPE rows of chunks >= cnts[b,h] and write them into
 * the K buffer at sequence offset offset_output_s_start + s (rows at or beyond offset_output_s_end
 * are skipped).  _opt and the plain variant compute the same thing; _glm rotates interleaved pairs
 * of the first 64 dims with a [cos(32)|sin(32)] table and copies dims 64..127. */
int skv_apply_rotary_pos_emb_push_cache(const skv_bf16 *x, const skv_bf16 *cos_sin, const int32_t *position_ids,
                                        skv_bf16 *output_cache, const int32_t *cnts, int batch_size, int heads,
                                        int seq_len, int embed_dim, int stride_xb, int stride_xh, int stride_xs,
                                        int stride_xe, int stride_cos_sin, int stride_pid_b, int stride_pid_h,
                                        int stride_pid_s, int stride_output_b, int stride_output_h,
                                        int stride_output_s, int offset_output_s_start, int offset_output_s_end,
                                        int half_dim, int chunk_size, skv_stream_t stream);
int skv_apply_rotary_pos_emb_push_cache_opt(const skv_bf16 *x, const skv_bf16 *cos_sin, const int32_t *position_ids,
                                            skv_bf16 *output_cache, const int32_t *cnts, int batch_size, int heads,
                                            int seq_len, int embed_dim, int stride_xb, int stride_xh, int stride_xs,
                                            int stride_xe, int stride_cos_sin, int stride_pid_b, int stride_pid_h,
                                            int stride_pid_s, int stride_output_b, int stride_output_h,
                                            int stride_output_s, int offset_output_s_start,
                                            int offset_output_s_end, int half_dim, int chunk_size,
                                            skv_stream_t stream);
int skv_apply_rotary_pos_emb_push_cache_opt_glm(const skv_bf16 *x, const skv_bf16 *cos_sin,
                                                const int32_t *position_ids, skv_bf16 *output_cache,
                                                const int32_t *cnts, int batch_size, int heads, int seq_len,
                                                int embed_dim, int stride_xb, int stride_xh, int stride_xs,
                                                int stride_xe, int stride_cos_sin, int stride_pid_b,
                                                int stride_pid_h, int stride_pid_s, int stride_output_b,
                                                int stride_output_h, int stride_output_s,
                                                int offset_output_s_start, int offset_output_s_end, int half_dim,
                                                int chunk_size, skv_stream_t stream);

/* functions.h:431-437  batch_gather_gemm (kernels/batch_gather_gemm.cu:193-287)
 * output[bh, r, :] = bf16( sum_k a[b, ids[bh, r/chunk]*chunk + r%chunk, k] * b[bh, :, k] )  (pre-RoPE K)
 * for the 128-row tiles not entirely below offset_array[bh]*chunk_size (the reference skips whole
 * tiles of cache hits, gemm_universal_batch_gather_indices.h:727-738).  a = U [batch, seq_len, rank],
 * b = SV [batch*heads, embed_dim, rank], ids int32 [batch*heads, sparse_budget/chunk_size].
 * cos / sin / max_seq_len are accepted and unused, as in the reference (its RoPE epilogue is disabled). */
int skv_batch_gather_gemm(const skv_bf16 *a, const skv_bf16 *b, const skv_bf16 *cos, const skv_bf16 *sin,
                          const int32_t *position_ids, skv_bf16 *output, int batch_size, int heads,
                          int seq_len, int embed_dim, int rank, int sparse_budget, int max_seq_len,
                          int chunk_size, const int32_t *offset_array, skv_stream_t stream);

/* functions.h:460-463  batch_gemm_softmax (kernels/batch_gemm_softmax.cu:229-293)
 * D[bc, i, j] = bf16(alpha * sum_k A[bc, i, k] * B[bc, j, k]);  Softmax = bf16(exp(D - rowmax) / rowsum).
 * A [batch_count, m, k], B [batch_count, n, k], D / Softmax [batch_count, m, n] bf16.
 * Norm / Sum [batch_count, m, ceil(n/256)] f32 scratch: on return Norm[bc, i, 0] holds the row maximum
 * and Sum[bc, i, 0] the reciprocal of the row sum (what the reference's final reduction leaves there).
 * beta must be 0.  m <= 16, k == 128. */
int skv_batch_gemm_softmax(const skv_bf16 *A, const skv_bf16 *B, skv_bf16 *D, float *Norm, float *Sum,
                           skv_bf16 *Softmax, int batch_count, int m, int n, int k, float alpha, float beta,
                           skv_stream_t stream);

/* ======================================================================
 * Part 2 -- fused B200 entry points
 * ====================================================================== */

/* Geometry shared by the fused calls.  All strides in elements. */
typedef struct skv_geom {
  int32_t batch;        /* B                                   */
  int32_t kv_heads;     /* H                                   */
  int32_t q_groups;     /* G = q heads per kv head (<= 16)     */
  int32_t head_dim;     /* D, must be 128                      */
  int32_t chunk;        /* C, must be 8                        */
  int32_t select;       /* S chunks in the sparse budget: 128/256/512/1024 */
  int32_t n_landmarks;  /* N, multiple of 8                    */
  int32_t rank;         /* R, multiple of 8, <= 256            */
  int32_t seq_len;      /* L rows of U per sequence            */
  int32_t buf_len;      /* Lbuf rows of the K/V buffers        */
  int32_t sparse_start; /* first row of the sparse region      */
  int32_t rope_glm;     /* 0: neox [cos64|sin64]; 1: GLM [cos32|sin32] */
  int32_t cos_sin_rows; /* rows of the cos_sin table           */
  int32_t table_mode;   /* 0: reference slot order (hits compacted to the front, needs skv_compact_hits);
                         * 1: in-place table (hits keep their slots; offsets[i] = slot of the i-th miss) */
  int32_t flags;        /* SKV_GEOM_* bits                     */
  int32_t *row_flags;   /* optional [B*H] words (device, zero-initialised, one array per layer): per-row hand-over from
                         * the selection to skv_fetch_rebuild_attend.  When set, the selection kernel clears its row's
                         * word after its own dependency wait and before it releases its dependents, and raises it
                         * (release) once the row's chunk table is written; the CTAs of the fetch launch wait for their
                         * row's word instead of for the completion of the whole selection grid.  Pass the SAME pointer
                         * to skv_select_chunks / skv_select_from_scores and to the skv_fetch_rebuild_attend call that
                         * follows it on the stream; NULL (all other entry points ignore it): plain dependent launches */
  const int32_t *wait_flags;      /* optional, kv-head split (skv_exchange): the scoring launch (skv_score_landmarks /  */
  const int32_t *wait_epoch_dev;  /* skv_select_chunks) first waits until wait_flags[0 .. wait_count) >= *wait_epoch_dev */
  int32_t wait_count;             /* -- the previous layer's gathered attention output is then complete in this rank's  */
  int32_t reserved;               /* buffer, which is what the layer's query depends on -- instead of a skv_wait_flags  */
                                  /* launch of its own between the layers.  NULL / 0: no wait                           */
  int32_t *select_hint; /* optional [B*H] words (device, zero-initialised, one array per layer), in/out of the selection:
                         * the bf16 key the row's top-S threshold ended on.  The next selection of the same row starts its
                         * radix search in a window of single-key bins around it and, when the new threshold lies inside
                         * (a factor of two either way), needs one histogram exchange instead of two.  Purely a speed
                         * hint: the selection is identical with any content, 0 or NULL = no hint */
} skv_geom;

/* The scoring launch is a programmatic dependent launch that streams the landmarks into shared memory BEFORE it
 * waits for the preceding kernel of the stream.  That is only valid when `landmarks` was completely written before
 * that preceding kernel STARTED (decode: the landmarks are static state built at prefill time).  A caller whose
 * landmarks are produced by the preceding kernel of the same stream (prefill: skv_gather_rows -> skv_select_chunks)
 * sets this flag; the kernel then loads nothing ahead of griddepcontrol.wait. */
#define SKV_GEOM_DYNAMIC_LANDMARKS 1

/* Output hand-over into the scoring launch (round 2).  With wait_flags set, the flag words normally come ON TOP of the
 * launch's wait for its stream predecessor.  When the predecessor is the skv_fetch_rebuild_attend launch whose merging
 * CTAs raise exactly those words (skv_exchange), the words already say everything the scoring launch consumes from it:
 * all rows of the previous layer's attention output are written.  With this bit the launch therefore waits for the
 * words only, does its work, and executes its wait for the predecessor's completion (and flush) at its very END, so the
 * ~2 us between "last row merged" and "grid complete, memory flushed, dependents released" leave the per-layer chain
 * while every kernel behind the scoring launch keeps the transitive guarantee that the predecessor has completed.
 * Only valid when every output of the predecessor that this launch (q, the appended rows' buffers, the score scratch)
 * touches is covered by the words; ignored without wait_flags. */
#define SKV_GEOM_WAIT_FLAGS_ONLY 2

/* Optional fused append (a2) for the scoring launch: the new token's K and V rows [B,H,D] are written at
 * buffer row pos (+ *pos_dev when pos_dev != NULL) of every (b,h) by the first CTA of each row, which saves
 * the separate skv_append_kv launch on the critical path. */
typedef struct skv_append {
  const skv_bf16 *k_new, *v_new;
  skv_bf16 *k_buf, *v_buf;
  const int32_t *pos_dev;
  int32_t pos;
  int32_t reserved;
} skv_append;

/* Bytes of device scratch skv_select_chunks needs for `geom` (scores + partials). */
int64_t skv_select_workspace_bytes(const skv_geom *geom);

/* (a3)+(a4)+(a5) in two launches: landmark scoring with fused scale/bf16 round and per-tile
 * softmax partials; then per (b,h): final softmax, max over the G query heads, top-S selection
 * (ties: lowest landmark index), chunk-id gather and the chunk-table update of
 * skv_reorder_keys_and_compute_offsets.  q [B, H*G, D] bf16 (post-RoPE); landmarks [B,H,N,D];
 * landmark_idx [B,H,N] int64; resident [B,H,S] int64 in/out; offsets [B*H*S], cnts [B*H] out.
 * selected_out (optional, may be NULL) receives the S selected chunk ids per (b,h) in ascending
 * landmark order; score_out (optional) the bf16 max-over-groups softmax [B,H,N]; append (optional) fuses (a2).
 * Precondition on `landmarks`: see SKV_GEOM_DYNAMIC_LANDMARKS. */
int skv_select_chunks(const skv_geom *geom, const skv_bf16 *q, const skv_bf16 *landmarks,
                      const int64_t *landmark_idx, int64_t *resident, int32_t *offsets, int32_t *cnts,
                      int64_t *selected_out, skv_bf16 *score_out, void *workspace, const skv_append *append,
                      skv_stream_t stream);

/* The two launches of skv_select_chunks, individually (same arguments, same workspace): the scoring pass
 * alone (HBM-bound stream over the landmarks; this is the kernel bench.py reports the roofline for) and the
 * per-(b,h) selection + chunk-table update that consumes its scores. */
int skv_score_landmarks(const skv_geom *geom, const skv_bf16 *q, const skv_bf16 *landmarks, void *workspace,
                        const skv_append *append, skv_stream_t stream);
int skv_select_from_scores(const skv_geom *geom, const int64_t *landmark_idx, int64_t *resident, int32_t *offsets,
                           int32_t *cnts, int64_t *selected_out, skv_bf16 *score_out, void *workspace,
                           skv_stream_t stream);

/* (a6)+(a7) hit compaction for K and V in one cluster-staged launch (no temp buffer):
 * for i < cnts[bh]: slot i <- slot offsets[bh*S+i] in both buffers.  Either buffer may be NULL. */
int skv_compact_hits(const skv_geom *geom, skv_bf16 *k_buf, skv_bf16 *v_buf, const int32_t *offsets,
                     const int32_t *cnts, skv_stream_t stream);

/* Optional HBM-side chunk store between the sparse slots and the pinned host store (a B200 extension: 180 GB of
 * HBM hold far more V chunks than the sparse budget).  Direct mapped per (b,h): chunk c lives in way c % ways.
 * skv_fetch_misses only READS it (a missed chunk whose tag matches is copied HBM -> slot instead of crossing PCIe);
 * skv_vstore_fill, launched after the fetch, installs the chunks that did cross PCIe (one deterministic winner
 * per way).  Results are byte-identical with and without the store. */
typedef struct skv_vstore {
  int32_t *tags;        /* [B*H, ways] chunk id held by every way, -1 = empty                      */
  skv_bf16 *data;       /* [B*H, ways, 1024]                                                        */
  int32_t *pcie_count;  /* optional [B*H]: += number of chunks that had to come from the host store */
  int32_t ways;         /* power of two                                                             */
  int32_t reserved;
} skv_vstore;

/* (a6) miss fetch: for i >= cnts[bh]: V slot i <- v_host[bh, offsets[bh*S+i]] (pinned host memory,
 * read over PCIe by the SMs).  host_stride = elements between consecutive (b,h) blocks of v_host.
 * geom->table_mode == 1: for i < S - cnts[bh]: slot = offsets[bh*S+i]; V slot <- v_host[bh, resident[bh*S+slot]]
 * (resident is required in that mode, ignored otherwise). */
int skv_fetch_misses(const skv_geom *geom, const skv_bf16 *v_host, int64_t host_stride, skv_bf16 *v_buf,
                     const int32_t *offsets, const int32_t *cnts, const int64_t *resident, const skv_vstore *vstore,
                     skv_stream_t stream);

/* Installs the chunks the preceding skv_fetch_misses (same arguments) brought over PCIe into the HBM chunk store:
 * copies them from their sparse slots into their ways and updates the tags.  When several missed chunks of one
 * (b,h) map to the same way the last one in miss order wins.  Must run after the fetch and before the next one. */
int skv_vstore_fill(const skv_geom *geom, const skv_bf16 *v_buf, const int32_t *offsets, const int32_t *cnts,
                    const int64_t *resident, const skv_vstore *vstore, skv_stream_t stream);

/* (a7) K rebuild for the missed chunks with RoPE in the epilogue (TMA-fed tcgen05, fp32 accumulate in
 * TMEM, bf16 round, bf16 RoPE at token position id*C+r), written straight into the sparse region of
 * k_buf.  U [B,L,R], SV [B,H,D,R], cos_sin [rows, D] or [rows, D/2] (GLM), resident [B,H,S] int64.
 * offsets: only read when geom->table_mode == 1 (slot of the i-th miss); may be NULL otherwise.
 * impl: 0 = tcgen05 (product path), 1 = SIMT cross-check used by the tests. */
int skv_rebuild_keys(const skv_geom *geom, const skv_bf16 *U, const skv_bf16 *SV, const skv_bf16 *cos_sin,
                     const int64_t *resident, const int32_t *cnts, const int32_t *offsets, skv_bf16 *k_buf,
                     int impl, skv_stream_t stream);

/* (a2) append the new token's K and V rows at buffer row `pos` of every (b,h). k_new/v_new [B,H,D].
 * pos_dev (optional, device int32): when non-NULL the row is pos + *pos_dev, read on the device, so the
 * same launch can be replayed from a CUDA graph while the generation offset advances. */
int skv_append_kv(const skv_geom *geom, const skv_bf16 *k_new, const skv_bf16 *v_new, skv_bf16 *k_buf,
                  skv_bf16 *v_buf, int pos, const int32_t *pos_dev, skv_stream_t stream);

/* Bytes of device scratch skv_attention_decode needs (split partials + arrival counters). */
int64_t skv_attention_workspace_bytes(int batch, int kv_heads, int q_groups, int head_dim, int max_splits);

/* (a8) replacement for flash_attn_with_kvcache(q, k_cache, v_cache, causal=True) with one query token
 * (reference models/base.py:341): out[b, h*G+g, :] = softmax(q K^T / sqrt(D)) V over rows [0, kv_len).
 * k/v element strides: (batch, head, row); rows are D contiguous elements.  workspace must be zeroed
 * once at allocation (the kernel leaves its counters at zero).  num_splits <= 0 picks a split count.
 * kv_len_dev (optional, device int32): when non-NULL the attended length is min(kv_len, *kv_len_dev) read on
 * the device and kv_len only sizes the grid (CUDA-graph replay with a growing cache). */
int skv_attention_decode(const skv_bf16 *q, const skv_bf16 *k, const skv_bf16 *v, skv_bf16 *out, int batch,
                         int kv_heads, int q_groups, int head_dim, int kv_len, int64_t stride_b,
                         int64_t stride_h, int64_t stride_row, float scale, int num_splits, void *workspace,
                         const int32_t *kv_len_dev, skv_stream_t stream);

/* skv_attention_decode with launch flags.  SKV_ATTN_PROGRAMMATIC: launch with programmatic stream serialization,
 * i.e. the kernel may be scheduled while the preceding kernel of the stream still runs and waits on the device
 * (griddepcontrol.wait) before it reads q, K, V or kv_len_dev.  Only for a caller whose preceding operation in
 * `stream` is the single kernel that produced the last of those inputs (skv_fetch_and_rebuild in the decode step). */
#define SKV_ATTN_PROGRAMMATIC 1
int skv_attention_decode_ex(const skv_bf16 *q, const skv_bf16 *k, const skv_bf16 *v, skv_bf16 *out, int batch,
                            int kv_heads, int q_groups, int head_dim, int kv_len, int64_t stride_b,
                            int64_t stride_h, int64_t stride_row, float scale, int num_splits, void *workspace,
                            const int32_t *kv_len_dev, int flags, skv_stream_t stream);

/* (a6) + (a7) in ONE launch (reference models/base.py:326-338 runs them on two streams and joins before the
 * attention): the missed chunks' V rows come from the pinned host store (or the HBM chunk store) exactly as in
 * skv_fetch_misses, their K rows are rebuilt exactly as in skv_rebuild_keys (tcgen05 path), by different CTAs of the
 * same grid.  With a single predecessor the attention that follows can be launched programmatically
 * (SKV_ATTN_PROGRAMMATIC).  Same arguments and table modes as the two calls it replaces; in the reference table
 * mode the caller compacts the hits of both buffers first (skv_compact_hits). */
int skv_fetch_and_rebuild(const skv_geom *geom, const skv_bf16 *v_host, int64_t host_stride, skv_bf16 *v_buf,
                          const skv_bf16 *U, const skv_bf16 *SV, const skv_bf16 *cos_sin, skv_bf16 *k_buf,
                          const int64_t *resident, const int32_t *cnts, const int32_t *offsets,
                          const skv_vstore *vstore, skv_stream_t stream);

/* kv-head split of a batch-1 long context over the GPUs of one node: every rank attends over its
 * H/N kv heads; the only exchange is the layer's attention output [B, Hq/N, 128].  Instead of a collective launch per
 * layer, the CTA that merges a row inside skv_fetch_rebuild_attend stores the row's output straight into EVERY rank's
 * gathered buffer over NVLink (peer-mapped pointers, e.g. torch symmetric memory) and then raises THAT ROW's flag word
 * on every peer (one system-scope fence, then the stores) with the value of *epoch_dev (a device counter the caller
 * advances once per token and never rewinds): no counting of rows, no second fence on the chain.  The consumer waits for
 * all N * B * H/N row words of the layer: skv_geom.wait_flags (inside the next scoring launch) or skv_wait_flags.
 * world == 1 with peer_out[0] == NULL is the ONE-GPU form of the same hand-over: no copy of the row, only the row's flag
 * word (device-scope fence); with SKV_GEOM_WAIT_FLAGS_ONLY the next layer's scoring launch starts on the rows' words
 * instead of on the completion of this grid. */
#define SKV_MAX_PEERS 8
typedef struct skv_exchange {
  int32_t world;                      /* number of ranks (0: no exchange)                                              */
  int32_t rank;                       /* this rank                                                                      */
  void *peer_out[SKV_MAX_PEERS];      /* rank r's buffer slot for THIS rank's output of the layer: [B, Hq/N, 128] bf16  */
  int32_t *peer_flags[SKV_MAX_PEERS]; /* rank r's flag words for (layer, this rank): [B * H/N], one per local row       */
  int32_t *rows_done;                 /* unused (kept for layout compatibility); may be NULL                            */
  const int32_t *epoch_dev;           /* device counter: value the flags take for the current token                     */
} skv_exchange;

/* Bytes of device scratch skv_fetch_rebuild_attend needs (softmax partials of the tile and split CTAs + arrival
 * counters).  Zero-filled ONCE by the caller; the kernel leaves its counters at zero. */
int64_t skv_attend_workspace_bytes(int batch, int kv_heads, int q_groups);

/* (a6) + (a7) + (a8) in ONE launch: everything of the decode step behind the chunk selection.  The reference joins its
 * two streams and only then calls flash_attn_with_kvcache (models/base.py:326-341), which re-reads every rebuilt K row
 * from HBM.  Here the CTA that rebuilds 16 missed chunks' K rows (tcgen05, RoPE in the TMEM epilogue) and fetches their V
 * chunks also takes their softmax partial on chip, while other CTAs of the same grid attend over what is already
 * resident (local window, outliers, chunk-cache hits, generated rows) during the PCIe transfer; the last CTA of a
 * (b, kv head) row merges.  K / V buffers, chunk table: exactly as skv_fetch_and_rebuild leaves them; out [B, H*G, 128]:
 * softmax(q K^T * scale) V over rows [0, kv_len) (min(kv_len, *kv_len_dev) when kv_len_dev != NULL).
 * Requires geom->table_mode == 1 (in-place chunk table: hits keep their slots) and q_groups <= 8; `v_store` is the
 * chunk-major V store wherever it lives (pinned host memory or HBM), store_stride = elements between (b,h) blocks.
 * Launched programmatically behind the selection kernel (its only predecessor in `stream`). */
int skv_fetch_rebuild_attend(const skv_geom *geom, const skv_bf16 *v_store, int64_t store_stride, skv_bf16 *v_buf,
                             const skv_bf16 *U, const skv_bf16 *SV, const skv_bf16 *cos_sin, skv_bf16 *k_buf,
                             const int64_t *resident, const int32_t *cnts, const int32_t *offsets,
                             const skv_vstore *vstore, const skv_bf16 *q, skv_bf16 *out, int kv_len,
                             const int32_t *kv_len_dev, float scale, void *workspace,
                             const skv_exchange *exchange, skv_stream_t stream);

/* Blocks `stream` (one tiny kernel) until flags[0..n) >= *epoch_dev (n <= 4096 row words: the outputs of all ranks for
 * that layer have landed in this rank's gathered buffer). */
int skv_wait_flags(const int32_t *flags, int n, const int32_t *epoch_dev, skv_stream_t stream);

/* ======================================================================
 * Part 3 -- prefill-side state build (reference models/kv_cache.py:788-980, ShadowKVCache_CPU.prefill_kv_cache)
 * ====================================================================== */

/* Per chunk of 8 keys of the post-RoPE prompt keys k (element strides batch / head / row, head_dim 128 contiguous):
 * means[bh, c, :] = bf16 mean of the chunk (landmark candidate, kv_cache.py:856) and mincos[bh, c] = min over the 8
 * keys of cosine_similarity(mean, key) (:859-866), with the bf16 rounding points of the torch ops the reference
 * applies to bf16 tensors. */
int skv_chunk_stats(const skv_bf16 *k, int batch, int kv_heads, int chunks, int64_t stride_b, int64_t stride_h,
                    int64_t stride_row, skv_bf16 *means, skv_bf16 *mincos, skv_stream_t stream);

/* Per row (b,h): the n_outliers chunks with the smallest mincos (kv_cache.py:868; ties: lowest chunk index) in
 * ascending chunk order -> outliers [rows, n_outliers]; all other chunks ascending -> rest_idx
 * [rows, chunks - n_outliers] (the landmark index, :909-915). */
int skv_select_outliers(const skv_bf16 *mincos, int rows, int chunks, int n_outliers, int64_t *outliers,
                        int64_t *rest_idx, skv_stream_t stream);

/* dst[row, i, :] = src[row, idx[row, i], :] for 128-element bf16 rows (landmarks = means[rest_idx], :917-921). */
int skv_gather_rows(const skv_bf16 *src, const int64_t *idx, skv_bf16 *dst, int rows, int n_src, int n,
                    skv_stream_t stream);

/* Chunk gather without the map-size restriction of skv_gather_copy: for every row and i < n_ids,
 *   dst[row*dst_stride + dst_offset + i*1024 ..] <- src[row*src_stride + ids[row*n_ids + i]*1024 ..]
 * (outlier K/V chunks :870-907, first fill of the sparse region :957-970).  src may be device or pinned host. */
int skv_gather_chunks(const skv_bf16 *src, int64_t src_stride, skv_bf16 *dst, int64_t dst_stride, int64_t dst_offset,
                      const int64_t *ids, int n_ids, int rows, skv_stream_t stream);

/* ======================================================================
 * Part 4 -- dense remainder of the decode layer: what reference models/llama.py:354-427
 * (pre_attention_compute / post_attention_compute), models/glm.py:312-403 and the lm_head of models/base.py:157
 * do for ONE decode token per sequence through library calls (flashinfer rmsnorm, F.linear = cuBLAS,
 * vllm rotary_embedding and silu_and_mul, torch residual adds).  Here: one weight-streaming kernel
 * (TMA ring + tcgen05, stream-K over 148 SMs, HBM bound) with the element-wise work in its epilogue.
 * ====================================================================== */

#define SKV_DENSE_F32 0      /* out f32 [B, out_features] = float(bf16(x W^T))            (lm_head, base.py:157) */
#define SKV_DENSE_RESID 1    /* out bf16 = bf16(resid + bf16(x W^T (+ bias)))  (o_proj / down_proj + residual,
                              * llama.py:410-412, 424-426); resid NULL: plain F.linear                          */
#define SKV_DENSE_SILU_MUL 2 /* W = [gate; up] (2*F rows): out bf16 [B, F] = bf16(silu(gate)) * up, rounding points
                              * of vllm silu_and_mul (llama.py:416-421)                                          */
#define SKV_DENSE_QKV_ROPE 3 /* W = fused wqkv ((q_heads + 2 kv_heads) * 128 rows): q [B,q_heads,128], k, v
                              * [B,kv_heads,128] bf16 with the rotary embedding of vllm rotary_embedding applied to
                              * q and k at the token's position (llama.py:296, 377-385; glm.py:344, 421)         */

#define SKV_DENSE_PARTIAL 4  /* out f32 [planes][B][out_features]: every CTA stores the partial sums of its slab range of a
                              * tile as plane (CTA - first CTA of the tile); nothing is reduced inside the kernel.  The
                              * consumer, skv_dense_finish_rmsnorm, adds the planes in the same order the other modes
                              * use, so the result is bit-identical to SKV_DENSE_RESID + skv_rmsnorm, without the
                              * cross-CTA fix-up (fence, arrival counter, re-read) at the end of the linear.           */

typedef struct skv_dense_epilogue {
  int32_t mode;          /* SKV_DENSE_*                                                         */
  int32_t q_heads;       /* QKV_ROPE                                                            */
  int32_t kv_heads;      /* QKV_ROPE                                                            */
  int32_t rope_mode;     /* QKV_ROPE: 0 none, 1 neox over 128 dims (cos_sin row = [cos64|sin64]),
                          * 2 GLM: interleaved pairs over the first 64 dims (row = [cos32|sin32]) */
  int32_t cos_sin_rows;  /* rows of the cos_sin table                                           */
  int32_t pos;           /* token position (+ *pos_dev when pos_dev != NULL)                    */
  int32_t static_weights; /* nonzero: W was written before the PREVIOUS kernel of the stream started (model
                          * weights).  The kernel then streams W into shared memory before it waits for that
                          * kernel (programmatic dependent launch), hiding its ramp-up.  0: W may have been
                          * produced by the previous kernel; nothing is read early.               */
  int32_t reserved;
  const int32_t *pos_dev;
  const skv_bf16 *bias;  /* [out_features] or NULL (GLM's bqkv)                                 */
  const skv_bf16 *resid; /* RESID: [B, out_features] or NULL; may alias out                     */
  const skv_bf16 *cos_sin;
  void *out;             /* F32: float*; RESID / SILU_MUL: bf16*; QKV_ROPE: q                   */
  skv_bf16 *k_out;       /* QKV_ROPE                                                            */
  skv_bf16 *v_out;       /* QKV_ROPE                                                            */
  /* QKV_ROPE, optional (all zero: the whole fused wqkv, k / v written as [B,kv_heads,128]): */
  int32_t head0;         /* W holds only the heads [head0, head0 + head_count) of the fused wqkv numbering
                          * (q heads, then k heads, then v heads); out_features = 128 * head_count.  Lets the q
                          * projection run on the critical path and the k / v projection beside it.           */
  int32_t head_count;    /* 0: all heads from head0 on                                                       */
  int32_t kv_row;        /* with kv_stride_h != 0: k / v are APPENDED in place (update_kv_cache,              */
  int32_t kv_rows;       /* kv_cache.py:1227-1271): element [b, head, kv_row + *pos_dev, :] of the cache      */
  int64_t kv_stride_b;   /* buffers k_out / v_out with these element strides; kv_rows = rows of those buffers: */
                         /* a row at or beyond kv_rows is dropped on the device, never written                 */
  int64_t kv_stride_h;
  const void *prefetch_next; /* optional hint: the weights the NEXT linear of the stream will read (16-byte aligned).   */
  int64_t prefetch_bytes;    /* Every CTA pulls its share of the first prefetch_bytes into L2 once it has issued its own
                              * last load, so HBM keeps streaming through the kernel's tail and the hand-over.         */
} skv_dense_epilogue;

/* Device scratch of skv_dense_linear (split-tile partial sums and per-tile arrival counters).  The caller
 * zero-fills it ONCE; the kernel leaves the counters at zero.  One workspace serves every launch of a stream. */
int64_t skv_dense_workspace_bytes(void);

/* y = x W^T for a skinny x [batch <= 64, in_features] bf16 and a row-major W [out_features, in_features] bf16
 * (torch Linear layout), fp32 accumulation, epilogue per `epi`.  in_features % 64 == 0. */
int skv_dense_linear(const skv_bf16 *x, const skv_bf16 *W, int batch, int in_features, int out_features,
                     const skv_dense_epilogue *epi, void *workspace, skv_stream_t stream);

/* Bytes of the plane buffer SKV_DENSE_PARTIAL needs for this problem (depends on how the slabs split over the SMs). */
int64_t skv_dense_partial_bytes(int batch, int in_features, int out_features);

/* Consumer of SKV_DENSE_PARTIAL for the two places a linear feeds a residual add and a norm (o_proj, down_proj;
 * llama.py:408-413, 424-426 + the next layer_norm):  y = bf16(sum of planes (+ bias));  x = bf16(x + y) in place;
 * h = rmsnorm(x, w, eps).  `in_features` / `out_features` are those of the linear that wrote the planes. */
int skv_dense_finish_rmsnorm(const float *planes, int batch, int in_features, int out_features, const skv_bf16 *bias,
                             skv_bf16 *x_inout, const skv_bf16 *w, float eps, skv_bf16 *h_out, skv_stream_t stream);

/* h[b,:] = bf16(x[b,:] * rsqrt(mean(x^2) + eps) * w) (flashinfer rmsnorm, tensor_op.py:34-39).  With tok != NULL the
 * row is first looked up, x[b,:] = emb[tok[b],:] (F.embedding, base.py:141), and written to x_out (the residual
 * stream) when x_out != NULL. */
int skv_rmsnorm(const skv_bf16 *x, const int64_t *tok, const skv_bf16 *emb, const skv_bf16 *w, float eps,
                skv_bf16 *x_out, skv_bf16 *h_out, int batch, int hidden, skv_stream_t stream);

/* Hint: pull `bytes` bytes at `ptr` (16-byte aligned; weights of an upcoming skv_dense_linear) into L2.  Returns as
 * soon as the prefetches are issued; reads nothing the caller can observe.  Used by the whole-token runner to keep
 * HBM streaming while the sparse-KV step waits on PCIe. */
int skv_l2_prefetch(const void *ptr, int64_t bytes, skv_stream_t stream);

/* tok[b] = index of the largest logit of row b (lowest index on ties): greedy decoding,
 * sample_token with temperature 0 (tensor_op.py:291-293, called at base.py:612). */
int skv_argmax(const float *logits, int64_t *tok, int batch, int vocab, skv_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SHADOWKV_B200_H */
