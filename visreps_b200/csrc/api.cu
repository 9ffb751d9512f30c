// Library-level entry points: ABI version, thread-local error text, launch counter.
#include "common.cuh"
#include <atomic>

namespace vr {
static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(unsigned n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace vr

extern "C" int vr_abi_version(void) { return VR_ABI_VERSION; }
extern "C" const char* vr_last_error_string(void) { return vr::g_err; }
extern "C" unsigned long long vr_launch_count(void) { return vr::g_launches.load(std::memory_order_relaxed); }
