// pvk_abi.cu -- error reporting and version of the C ABI (include/pvk.h).
#include <stdarg.h>

#include "pvk_common.cuh"

namespace pvk {
static thread_local char g_err[512] = "";

static long long g_launches = 0;
void count_launch() { __atomic_fetch_add(&g_launches, 1, __ATOMIC_RELAXED); }

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace pvk

extern "C" int pvk_version(void) { return 5; }

extern "C" const char *pvk_last_error(void) { return pvk::g_err; }

extern "C" int64_t pvk_launch_count(void) { return (int64_t)__atomic_load_n(&pvk::g_launches, __ATOMIC_RELAXED); }
