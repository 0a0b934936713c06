// libmcn runtime: thread-local error text, launch accounting, version.
#include <cstring>

#include "mcn_common.cuh"

namespace mcn {

static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

}  // namespace mcn

extern "C" const char* mcn_last_error(void) { return mcn::g_err; }
extern "C" int mcn_version(void) { return 100; }
extern "C" long long mcn_launch_count(void) {
  return mcn::g_launches.load(std::memory_order_relaxed);
}
