// error.cu -- thread-local error string + ABI version for libmsmd_b200.
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

namespace msmd {
static thread_local char g_error[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

static std::atomic<unsigned long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
}  // namespace msmd

extern "C" MSMD_API const char* msmd_last_error(void) { return msmd::g_error; }
extern "C" MSMD_API int msmd_abi_version(void) { return MSMD_ABI_VERSION; }
extern "C" MSMD_API unsigned long long msmd_launch_count(void) {
  return msmd::g_launches.load(std::memory_order_relaxed);
}
