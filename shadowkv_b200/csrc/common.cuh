// Shared device/host helpers for the shadowkv_b200 kernels (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/shadowkv_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "shadowkv_b200 kernels are written for sm_100a (B200) only"
#endif

namespace skv {

constexpr int kChunkTokens = 8;     // C (reference copy.cuh:321-343: chunk = 128 thr x 16 B = 2 KiB)
constexpr int kHeadDim = 128;       // D
constexpr int kChunkElems = kChunkTokens * kHeadDim;   // 1024 bf16
constexpr int kChunkBytes = kChunkElems * 2;           // 2048 B
constexpr int kNumSMs = 148;

void set_error(const char *fmt, ...);

// Optional phase timeline for kernel tuning (build with `make TRACE=1`): thread 0 of a CTA stamps clock64()
// into a per-translation-unit table, read back through skv_debug_read_trace_<unit>().  Compiled out of the
// product library.
#ifdef SKV_TRACE
static __device__ long long g_skv_trace[8192];
#define SKV_STAMP(cta, slot)                                                             \
  do {                                                                                   \
    if (threadIdx.x == 0 && (cta) < 256) g_skv_trace[(cta) * 32 + (slot)] = clock64();   \
  } while (0)
#define SKV_STAMP_ANY(cta, slot)                                                         \
  do {                                                                                   \
    if ((cta) < 256) g_skv_trace[(cta) * 32 + (slot)] = clock64();                       \
  } while (0)
// wall-clock variant (%globaltimer, ns): comparable across CTAs and across consecutive launches
#define SKV_STAMP_NS(row, slot)                                                          \
  do {                                                                                   \
    if ((row) >= 0 && (row) < 256) {                                                     \
      unsigned long long t_;                                                             \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                             \
      g_skv_trace[(row) * 32 + (slot)] = (long long)t_;                                  \
    }                                                                                    \
  } while (0)
#define SKV_TRACE_READER(unit)                                                           \
  extern "C" int skv_debug_read_trace_##unit(long long *host_out, int n) {               \
    cudaDeviceSynchronize();                                                             \
    return (int)cudaMemcpyFromSymbol(host_out, skv::g_skv_trace, sizeof(long long) * (n < 8192 ? n : 8192)); \
  }
#else
#define SKV_STAMP(cta, slot) do { } while (0)
#define SKV_STAMP_ANY(cta, slot) do { } while (0)
#define SKV_STAMP_NS(row, slot) do { } while (0)
#define SKV_TRACE_READER(unit)
#endif

inline cudaStream_t as_stream(skv_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// Collect the launch status without synchronising.
inline int launch_status(const char *what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  return SKV_OK;
}

inline bool valid_map_size(int s) { return s == 128 || s == 256 || s == 512 || s == 1024; }

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------
// The kernels of the step form a chain in one stream.  Launched with the programmatic-serialization attribute,
// a kernel may START while its predecessor still runs: its prologue (barrier init, TMEM allocation, loads of
// data the predecessor does not produce, e.g. the landmark stream) overlaps the predecessor's tail, and the
// launch latency disappears from the critical path.  pdl_wait() blocks until the predecessor has completed and
// its writes are visible; pdl_launch_dependents() lets the successor be scheduled.  Both are no-ops when the
// kernel was launched normally.  SKV_PDL=0 in the environment disables the attribute (A/B measurements).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool pdl_enabled();

// Opt-in for > 48 KiB of dynamic shared memory, remembered per (kernel, device).  Returns SKV_OK or the CUDA error.
int ensure_dynamic_smem(const void *kernel, int bytes, const char *what);

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                                 Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl && pdl_enabled()) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- 128-bit streaming loads / stores --------------------------------------------------------
__device__ __forceinline__ int4 ld_stream(const int4 *p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(int4 *p, const int4 &v) {
  asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}

// ---- bf16 helpers -----------------------------------------------------------------------------
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint16_t f32_to_bf16_bits(float f) {
  return __bfloat16_as_ushort(__float2bfloat16_rn(f));
}
__device__ __forceinline__ float bf16_bits_to_f32(uint16_t b) { return __uint_as_float(uint32_t(b) << 16); }

// RoPE element in bf16, in the exact operation order ptxas gives the reference kernels on
// sm_100a (reference kernels/rope_new.cu:366-367; SASS: one HMUL2.BF16 for the sin product, then
// one HFMA2.BF16 that folds the cos product into the add).  Written with explicit fma so the
// result does not depend on contraction heuristics.
//   first  = x1*c + bf16(-x2*s)     second = x2*c + bf16(x1*s)
__device__ __forceinline__ void rope_pair_bf16(__nv_bfloat16 x1, __nv_bfloat16 x2, __nv_bfloat16 c,
                                               __nv_bfloat16 s, __nv_bfloat16 &o1, __nv_bfloat16 &o2) {
  o1 = __hfma(x1, c, __hmul(__hneg(x2), s));
  o2 = __hfma(x2, c, __hmul(x1, s));
}

__device__ __forceinline__ uint32_t pack_bf16(__nv_bfloat16 lo, __nv_bfloat16 hi) {
  return uint32_t(__bfloat16_as_ushort(lo)) | (uint32_t(__bfloat16_as_ushort(hi)) << 16);
}

// ---- warp / block primitives -------------------------------------------------------------------
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_incl_scan(int v) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int n = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += n;
  }
  return v;
}

// Exclusive block scan of one int per thread (blockDim.x <= 1024, multiple of 32).
// `warp_tot` is 33 ints of shared scratch; returns the exclusive prefix, `total` the block sum.
__device__ __forceinline__ int block_excl_scan(int v, int *warp_tot, int &total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  int incl = warp_incl_scan(v);
  __syncthreads();                      // protect warp_tot from a previous use
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = lane < nwarps ? warp_tot[lane] : 0;
    int ws = warp_incl_scan(w);
    warp_tot[lane] = ws - w;            // exclusive per-warp base
    if (lane == 31) warp_tot[32] = ws;
  }
  __syncthreads();
  total = warp_tot[32];
  return warp_tot[warp] + incl - v;
}

}  // namespace skv
