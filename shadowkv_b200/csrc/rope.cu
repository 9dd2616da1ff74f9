// The six stand-alone RoPE entry points of the reference extension (kernels/rope.cu, rope_new.cu).
// bf16 arithmetic in the reference's operation order (see rope_pair_bf16 in common.cuh); two pairs per
// thread with packed bf16x2 math.  In the fused decode path RoPE runs in the K-rebuild epilogue instead
// (rebuild.cu); these exist for the drop-in API and as its cross-check.
#include "common.cuh"

namespace skv {

struct RopeParams {
  const __nv_bfloat16 *x;
  const __nv_bfloat16 *cos;   // cos table (or combined table)
  const __nv_bfloat16 *sin;   // sin table; for a combined table = cos + half_dim (neox) / + quarter (GLM)
  const void *pos;            // int64 token positions or int32 chunk ids
  __nv_bfloat16 *out;
  const int32_t *cnts;        // push-cache only
  int batch, heads, seq_len, half_dim;
  int sxb, sxh, sxs;
  int s_cos, s_sin;
  int spb, sph, sps;
  int sob, soh, sos;          // output strides (== x strides when not pushing)
  int out_s_start, out_s_end; // push-cache window
  int chunk;                  // chunk size when pos holds chunk ids
  int sin_second_off;         // rope.cu variant: second half uses cos/sin[tid+half]; else 0
};

__device__ __forceinline__ __nv_bfloat162 ld2(const __nv_bfloat16 *p) { return *reinterpret_cast<const __nv_bfloat162 *>(p); }
__device__ __forceinline__ void st2(__nv_bfloat16 *p, __nv_bfloat162 v) { *reinterpret_cast<__nv_bfloat162 *>(p) = v; }

// kGlm: interleaved pairs over the first half_dim dims, copy of the rest (rope_new.cu:429-493).
// kChunkIds: pos holds int32 chunk ids, token = id*chunk + s%chunk (rope_new.cu:135-205).
// kPush: write into the K buffer window, skipping chunks below cnts[b,h] (rope_new.cu:219-370).
template <bool kGlm, bool kChunkIds, bool kPush>
__global__ void __launch_bounds__(256) rope_kernel(RopeParams p) {
  const int per_row = p.half_dim / 2;                       // threads per row (2 pairs each)
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long rows = (long long)p.batch * p.heads * p.seq_len;
  const long long row = gid / per_row;
  if (row >= rows) return;
  const int t = int(gid % per_row);
  const int s = int(row % p.seq_len);
  const int h = int((row / p.seq_len) % p.heads);
  const int b = int(row / ((long long)p.seq_len * p.heads));

  long long token;
  if (kChunkIds) {
    const int32_t *ids = static_cast<const int32_t *>(p.pos);
    if (kPush) {
      if (s / p.chunk < p.cnts[b * p.heads + h]) return;    // cache hit: row already in place
      if (p.out_s_start + s >= p.out_s_end) return;
    }
    const int id = ids[(long long)b * p.spb + (long long)h * p.sph + (long long)(s / p.chunk) * p.sps];
    token = (long long)id * p.chunk + s % p.chunk;
  } else {
    const int64_t *tk = static_cast<const int64_t *>(p.pos);
    token = (int)tk[(long long)b * p.spb + (long long)h * p.sph + (long long)s * p.sps];   // int truncation as reference
  }
  const __nv_bfloat16 *xr = p.x + (long long)b * p.sxb + (long long)h * p.sxh + (long long)s * p.sxs;
  __nv_bfloat16 *orow = kPush ? p.out + (long long)b * p.sob + (long long)h * p.soh + (long long)(p.out_s_start + s) * p.sos
                              : p.out + (long long)b * p.sxb + (long long)h * p.sxh + (long long)s * p.sxs;
  const __nv_bfloat16 *cr = p.cos + token * p.s_cos;
  const __nv_bfloat16 *sr = p.sin + token * p.s_sin;

  if (!kGlm) {
    const int j = 2 * t;
    const __nv_bfloat162 x1 = ld2(xr + j), x2 = ld2(xr + j + p.half_dim);
    const __nv_bfloat162 c1 = ld2(cr + j), s1 = ld2(sr + j);
    const __nv_bfloat162 c2 = p.sin_second_off ? ld2(cr + j + p.sin_second_off) : c1;
    const __nv_bfloat162 s2 = p.sin_second_off ? ld2(sr + j + p.sin_second_off) : s1;
    st2(orow + j, __hfma2(x1, c1, __hmul2(__hneg2(x2), s1)));
    st2(orow + j + p.half_dim, __hfma2(x2, c2, __hmul2(x1, s2)));
  } else {
    // threads [0, per_row/2) rotate 2 interleaved pairs each; the rest copy the upper half
    const int q = p.half_dim / 2;                           // 32 rotated pairs
    if (t < q / 2) {
      const int j = 2 * t;                                  // pair index
      const __nv_bfloat162 a = ld2(xr + 2 * j), bq = ld2(xr + 2 * j + 2);   // (e0,o0) (e1,o1)
      const __nv_bfloat162 c = ld2(cr + j), sn = ld2(sr + j);
      const __nv_bfloat162 xe = __halves2bfloat162(__low2bfloat16(a), __low2bfloat16(bq));
      const __nv_bfloat162 xo = __halves2bfloat162(__high2bfloat16(a), __high2bfloat16(bq));
      const __nv_bfloat162 oe = __hfma2(xe, c, __hmul2(__hneg2(xo), sn));
      const __nv_bfloat162 oo = __hfma2(xo, c, __hmul2(xe, sn));
      st2(orow + 2 * j, __halves2bfloat162(__low2bfloat16(oe), __low2bfloat16(oo)));
      st2(orow + 2 * j + 2, __halves2bfloat162(__high2bfloat16(oe), __high2bfloat16(oo)));
    } else {
      const int u = t - q / 2;                              // 0..15 -> 4 elements each of dims [64,128)
      const int j = p.half_dim + 4 * u;
      st2(orow + j, ld2(xr + j));
      st2(orow + j + 2, ld2(xr + j + 2));
    }
  }
}

template <bool kGlm, bool kChunkIds, bool kPush>
static int launch_rope(const RopeParams &p, cudaStream_t st, const char *name) {
  if (p.half_dim <= 0 || (p.half_dim & 3) || (p.sxs & 1) || (p.sxh & 1) || (p.sxb & 1) || (p.s_cos & 1) ||
      (kPush && ((p.sos & 1) || (p.soh & 1) || (p.sob & 1))) || (kGlm && p.half_dim != 64)) {
    set_error("%s: strides/half_dim not supported (need even strides, half_dim %% 4 == 0)", name);
    return SKV_ERR_UNSUPPORTED;
  }
  const long long threads = (long long)p.batch * p.heads * p.seq_len * (p.half_dim / 2);
  if (threads <= 0) return SKV_OK;
  const long long blocks = (threads + 255) / 256;
  rope_kernel<kGlm, kChunkIds, kPush><<<(unsigned)blocks, 256, 0, st>>>(p);
  return launch_status(name);
}

static inline const __nv_bfloat16 *bf(const skv_bf16 *p) { return reinterpret_cast<const __nv_bfloat16 *>(p); }
static inline __nv_bfloat16 *bf(skv_bf16 *p) { return reinterpret_cast<__nv_bfloat16 *>(p); }

}  // namespace skv

using namespace skv;

extern "C" {

int skv_apply_rotary_pos_emb(const skv_bf16 *x, const skv_bf16 *cos, const skv_bf16 *sin, const int64_t *position_ids,
                             skv_bf16 *output, int batch_size, int heads, int seq_len, int embed_dim, int stride_xb,
                             int stride_xh, int stride_xs, int stride_xe, int stride_cos, int stride_sin,
                             int stride_pid_b, int stride_pid_h, int stride_pid_s, int half_dim, skv_stream_t stream) {
  (void)embed_dim; (void)stride_xe;
  if (!x || !cos || !sin || !position_ids || !output) { set_error("apply_rotary_pos_emb: null pointer"); return SKV_ERR_ARGUMENT; }
  if (stride_sin & 1) { set_error("apply_rotary_pos_emb: odd sin stride"); return SKV_ERR_UNSUPPORTED; }
  RopeParams p{};
  p.x = bf(x); p.cos = bf(cos); p.sin = bf(sin); p.pos = position_ids; p.out = bf(output);
  p.batch = batch_size; p.heads = heads; p.seq_len = seq_len; p.half_dim = half_dim;
  p.sxb = stride_xb; p.sxh = stride_xh; p.sxs = stride_xs; p.s_cos = stride_cos; p.s_sin = stride_sin;
  p.spb = stride_pid_b; p.sph = stride_pid_h; p.sps = stride_pid_s; p.sin_second_off = half_dim;
  return launch_rope<false, false, false>(p, as_stream(stream), "apply_rotary_pos_emb");
}

int skv_apply_rotary_pos_emb_new(const skv_bf16 *x, const skv_bf16 *cos_sin, const int64_t *position_ids,
                                 skv_bf16 *output, int batch_size, int heads, int seq_len, int embed_dim,
                                 int stride_xb, int stride_xh, int stride_xs, int stride_xe, int stride_cos_sin,
                                 int stride_pid_b, int stride_pid_h, int stride_pid_s, int half_dim,
                                 skv_stream_t stream) {
  (void)embed_dim; (void)stride_xe;
  if (!x || !cos_sin || !position_ids || !output) { set_error("apply_rotary_pos_emb_new: null pointer"); return SKV_ERR_ARGUMENT; }
  RopeParams p{};
  p.x = bf(x); p.cos = bf(cos_sin); p.sin = bf(cos_sin) + half_dim; p.pos = position_ids; p.out = bf(output);
  p.batch = batch_size; p.heads = heads; p.seq_len = seq_len; p.half_dim = half_dim;
  p.sxb = stride_xb; p.sxh = stride_xh; p.sxs = stride_xs; p.s_cos = stride_cos_sin; p.s_sin = stride_cos_sin;
  p.spb = stride_pid_b; p.sph = stride_pid_h; p.sps = stride_pid_s;
  return launch_rope<false, false, false>(p, as_stream(stream), "apply_rotary_pos_emb_new");
}

int skv_apply_rotary_pos_emb_new_v2(const skv_bf16 *x, const skv_bf16 *cos_sin, const int32_t *position_ids,
                                    skv_bf16 *output, int batch_size, int heads, int seq_len, int embed_dim,
                                    int stride_xb, int stride_xh, int stride_xs, int stride_xe, int stride_cos_sin,
                                    int stride_pid_b, int stride_pid_h, int stride_pid_s, int half_dim,
                                    int chunk_size, skv_stream_t stream) {
  (void)embed_dim; (void)stride_xe;
  if (!x || !cos_sin || !position_ids || !output || chunk_size <= 0) { set_error("apply_rotary_pos_emb_new_v2: bad argument"); return SKV_ERR_ARGUMENT; }
  RopeParams p{};
  p.x = bf(x); p.cos = bf(cos_sin); p.sin = bf(cos_sin) + half_dim; p.pos = position_ids; p.out = bf(output);
  p.batch = batch_size; p.heads = heads; p.seq_len = seq_len; p.half_dim = half_dim;
  p.sxb = stride_xb; p.sxh = stride_xh; p.sxs = stride_xs; p.s_cos = stride_cos_sin; p.s_sin = stride_cos_sin;
  p.spb = stride_pid_b; p.sph = stride_pid_h; p.sps = stride_pid_s; p.chunk = chunk_size;
  return launch_rope<false, true, false>(p, as_stream(stream), "apply_rotary_pos_emb_new_v2");
}

static int push_cache_common(bool glm, const char *name, const skv_bf16 *x, const skv_bf16 *cos_sin,
                             const int32_t *position_ids, skv_bf16 *output_cache, const int32_t *cnts,
                             int batch_size, int heads, int seq_len, int stride_xb, int stride_xh, int stride_xs,
                             int stride_cos_sin, int stride_pid_b, int stride_pid_h, int stride_pid_s,
                             int stride_output_b, int stride_output_h, int stride_output_s,
                             int offset_output_s_start, int offset_output_s_end, int half_dim, int chunk_size,
                             skv_stream_t stream) {
  if (!x || !cos_sin || !position_ids || !output_cache || !cnts || chunk_size <= 0) { set_error("%s: bad argument", name); return SKV_ERR_ARGUMENT; }
  RopeParams p{};
  p.x = bf(x); p.cos = bf(cos_sin); p.sin = bf(cos_sin) + (glm ? half_dim / 2 : half_dim);
  p.pos = position_ids; p.out = bf(output_cache); p.cnts = cnts;
  p.batch = batch_size; p.heads = heads; p.seq_len = seq_len; p.half_dim = half_dim;
  p.sxb = stride_xb; p.sxh = stride_xh; p.sxs = stride_xs; p.s_cos = stride_cos_sin; p.s_sin = stride_cos_sin;
  p.spb = stride_pid_b; p.sph = stride_pid_h; p.sps = stride_pid_s;
  p.sob = stride_output_b; p.soh = stride_output_h; p.sos = stride_output_s;
  p.out_s_start = offset_output_s_start; p.out_s_end = offset_output_s_end; p.chunk = chunk_size;
  return glm ? launch_rope<true, true, true>(p, as_stream(stream), name)
             : launch_rope<false, true, true>(p, as_stream(stream), name);
}

#define SKV_PUSH_ARGS                                                                                             \
  const skv_bf16 *x, const skv_bf16 *cos_sin, const int32_t *position_ids, skv_bf16 *output_cache,                \
      const int32_t *cnts, int batch_size, int heads, int seq_len, int embed_dim, int stride_xb, int stride_xh,   \
      int stride_xs, int stride_xe, int stride_cos_sin, int stride_pid_b, int stride_pid_h, int stride_pid_s,     \
      int stride_output_b, int stride_output_h, int stride_output_s, int offset_output_s_start,                   \
      int offset_output_s_end, int half_dim, int chunk_size, skv_stream_t stream
#define SKV_PUSH_FWD                                                                                              \
  x, cos_sin, position_ids, output_cache, cnts, batch_size, heads, seq_len, stride_xb, stride_xh, stride_xs,     \
      stride_cos_sin, stride_pid_b, stride_pid_h, stride_pid_s, stride_output_b, stride_output_h,                \
      stride_output_s, offset_output_s_start, offset_output_s_end, half_dim, chunk_size, stream

int skv_apply_rotary_pos_emb_push_cache(SKV_PUSH_ARGS) {
  (void)embed_dim; (void)stride_xe;
  return push_cache_common(false, "apply_rotary_pos_emb_push_cache", SKV_PUSH_FWD);
}
int skv_apply_rotary_pos_emb_push_cache_opt(SKV_PUSH_ARGS) {
  (void)embed_dim; (void)stride_xe;
  return push_cache_common(false, "apply_rotary_pos_emb_push_cache_opt", SKV_PUSH_FWD);
}
int skv_apply_rotary_pos_emb_push_cache_opt_glm(SKV_PUSH_ARGS) {
  (void)embed_dim; (void)stride_xe;
  return push_cache_common(true, "apply_rotary_pos_emb_push_cache_opt_glm", SKV_PUSH_FWD);
}

}  // extern "C"
