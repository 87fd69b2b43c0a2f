// error.cu -- thread-local error string + ABI version for libmsmd_b200.
#include <stdarg.h>

#include "common.cuh"

namespace msmd {
static thread_local char g_error[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}
}  // namespace msmd

extern "C" MSMD_API const char* msmd_last_error(void) { return msmd::g_error; }
extern "C" MSMD_API int msmd_abi_version(void) { return MSMD_ABI_VERSION; }
