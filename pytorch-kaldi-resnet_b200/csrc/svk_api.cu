// Library-level entry points: version, error string, launch counter.
#include "svk_common.cuh"
#include <stdarg.h>
#include <stdlib.h>
#include <atomic>

static thread_local char g_err[512] = "no error";
static std::atomic<long long> g_launches{0};

void svk_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
bool svk_pdl_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("SVK_DISABLE_PDL"); on = (e && e[0] == '1') ? 0 : 1; }
  return on == 1;
}
void svk_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

SVK_API int svk_version(void) { return SVK_VERSION; }
SVK_API const char* svk_last_error_string(void) { return g_err; }
SVK_API long long svk_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
