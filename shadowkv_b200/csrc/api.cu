// Error reporting and identification for the C-ABI.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

namespace skv {
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
bool pdl_enabled() {
  static const bool on = [] { const char *e = getenv("SKV_PDL"); return !(e && e[0] == '0'); }();
  return on;
}
}  // namespace skv

extern "C" {
const char *skv_version(void) { return "shadowkv_b200 0.2 (sm_100a; tcgen05 scoring / K rebuild / dense stream-K linears, TMA rings)"; }
const char *skv_last_error(void) { return skv::g_err; }
}
