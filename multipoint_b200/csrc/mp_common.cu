// Error string, launch counter and version of the C ABI.
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

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

// ---- per-kernel timing with CUDA events on the launching stream ----
struct ProfRec { const char *name; cudaEvent_t ev; cudaStream_t stream; };
static std::atomic<int> g_prof_on{0};
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;

void prof_mark(const char *name, cudaStream_t stream) {
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    ProfRec r{name, nullptr, stream};
    if (cudaEventCreate(&r.ev) != cudaSuccess) return;
    cudaEventRecord(r.ev, stream);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.push_back(r);
}
}  // namespace mp

extern "C" int mp_profile_begin(void) {
    std::lock_guard<std::mutex> lk(mp::g_prof_mu);
    for (auto &r : mp::g_prof) cudaEventDestroy(r.ev);
    mp::g_prof.clear();
    mp::g_prof_on.store(1);
    return MP_OK;
}

// Writes a JSON object {"kernel": {"launches": n, "total_ms": t}, ...} into buf (NUL terminated,
// truncated to cap) and returns the length that was needed.  Synchronises the recorded events.
extern "C" size_t mp_profile_end(char *buf, size_t cap) {
    mp::g_prof_on.store(0);
    std::lock_guard<std::mutex> lk(mp::g_prof_mu);
    std::map<std::string, std::pair<long, double>> agg;
    std::map<cudaStream_t, cudaEvent_t> prev;
    for (auto &r : mp::g_prof) {
        cudaEventSynchronize(r.ev);
        auto it = prev.find(r.stream);
        if (r.name != nullptr && it != prev.end()) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, it->second, r.ev) == cudaSuccess) {
                auto &a = agg[r.name];
                a.first += 1;
                a.second += ms;
            }
        }
        prev[r.stream] = r.ev;
    }
    std::string js = "{";
    bool first = true;
    for (auto &kv : agg) {
        char tmp[256];
        snprintf(tmp, sizeof(tmp), "%s\"%s\": {\"launches\": %ld, \"total_ms\": %.6f}", first ? "" : ", ", kv.first.c_str(),
                 kv.second.first, kv.second.second);
        js += tmp;
        first = false;
    }
    js += "}";
    for (auto &r : mp::g_prof) cudaEventDestroy(r.ev);
    mp::g_prof.clear();
    if (buf != nullptr && cap > 0) {
        const size_t n = js.size() < cap - 1 ? js.size() : cap - 1;
        memcpy(buf, js.data(), n);
        buf[n] = 0;
    }
    return js.size() + 1;
}

extern "C" int mp_version(void) { return 100; }
extern "C" const char *mp_last_error_string(void) { return mp::g_err; }
extern "C" unsigned long long mp_launch_count(void) { return mp::g_launches.load(std::memory_order_relaxed); }
