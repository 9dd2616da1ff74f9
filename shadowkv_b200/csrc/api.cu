// Error reporting and identification for the C-ABI.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <utility>

#include "common.cuh"

namespace skv {
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
// Opt-in for more than 48 KiB of dynamic shared memory.  cudaFuncSetAttribute applies to the CURRENT device, so the
// largest size already granted is remembered per (kernel, device); thread safe.
int ensure_dynamic_smem(const void *kernel, int bytes, const char *what) {
  static std::mutex mu;
  static std::map<std::pair<const void *, int>, int> granted;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { set_error("%s: cudaGetDevice: %s", what, cudaGetErrorString(e)); return int(e); }
  std::lock_guard<std::mutex> lock(mu);
  int &cur = granted[std::make_pair(kernel, dev)];
  if (bytes > cur) {
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) { set_error("%s: %d bytes of dynamic shared memory: %s", what, bytes, cudaGetErrorString(e)); return int(e); }
    cur = bytes;
  }
  return SKV_OK;
}
bool pdl_enabled() {
  static const bool on = [] { const char *e = getenv("SKV_PDL"); return !(e && e[0] == '0'); }();
  return on;
}
}  // namespace skv

extern "C" {
const char *skv_version(void) { return "shadowkv_b200 0.2 (sm_100a; tcgen05 scoring / K rebuild / dense stream-K linears, TMA rings)"; }
const char *skv_last_error(void) { return skv::g_err; }
// Which implementation every stage of the decode step runs in THIS process.  The cross-check kernels that ship inside
// the library (SIMT scoring / attention / rebuild) and the A/B knobs are selected by environment variables read here
// and at the launch sites; a bench or test line that prints this string cannot pick one of them up silently.
const char *skv_active_config(void) {
  static thread_local char buf[384];
  auto on = [](const char *name, char v) { const char *e = getenv(name); return e && e[0] == v; };
  const char *ctas = getenv("SKV_SCORE_CTAS");
  snprintf(buf, sizeof(buf),
           "score=%s select=%s rebuild=%s attention=%s attend_fusion=%s pdl=%s score_ring=%s score_ctas=%s",
           on("SKV_SCORE_IMPL", '1') ? "SIMT(cross-check)" : "tcgen05+TMA",
           (on("SKV_FUSED_SELECT", '1') && !on("SKV_SCORE_IMPL", '1')) ? "fused-with-scoring(experiment)"
           : (on("SKV_SELECT_FAST", '1') && !on("SKV_SELECT_HINT", '0')) ? "own-launch(1-CTA-hinted+cluster8-radix, experiment)"
                                                                        : "own-launch(cluster8-radix)",
           on("SKV_REBUILD_IMPL", '1') ? "SIMT(cross-check)" : "tcgen05",
           on("SKV_ATTN_IMPL", '1') ? "SIMT(cross-check)" : "mma.sync+TMA", on("SKV_FUSED_ATTEND", '0') ? "off" : "on",
           skv::pdl_enabled() ? "on" : "off", on("SKV_SCORE_STAGES", '2') ? "2x3" : "3x2", (ctas && *ctas) ? ctas : "default");
  return buf;
}
}
