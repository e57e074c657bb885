// Error string, version and launch accounting for liblfp_sg2.
#include <stdarg.h>
#include <atomic>
#include "common.cuh"

namespace lfp {
static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
}  // namespace lfp

extern "C" const char* lfp_last_error(void) { return lfp::g_err; }
extern "C" int lfp_version(void) { return 100; }
extern "C" uint64_t lfp_launch_count(void) { return lfp::g_launches.load(); }
