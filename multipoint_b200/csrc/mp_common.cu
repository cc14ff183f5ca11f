// Error string, launch counter and version of the C ABI.
#include <stdarg.h>

#include <atomic>

#include "mp_common.cuh"

namespace mp {
static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(unsigned n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace mp

extern "C" int mp_version(void) { return 100; }
extern "C" const char *mp_last_error_string(void) { return mp::g_err; }
extern "C" unsigned long long mp_launch_count(void) { return mp::g_launches.load(std::memory_order_relaxed); }
