// Error reporting and launch accounting shared by every entry point of libstb200.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace stb200 {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return STB200_ERR_CUDA;
    }
    return STB200_OK;
}

// ---- per-kernel CUDA-event profiler -------------------------------------------------------------------
// When enabled, every kernel launch of the library is bracketed by two events recorded on the launching
// stream; stb200_profile_dump() synchronises, sums durations and algorithmic bytes per kernel name.
struct ProfRec {
    const char *name;
    double bytes;
    cudaEvent_t e0, e1;
};
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static std::vector<cudaEvent_t> g_event_pool;

static cudaEvent_t take_event() {
    if (!g_event_pool.empty()) {
        cudaEvent_t e = g_event_pool.back();
        g_event_pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}

KernelScope::KernelScope(const char *name, double algorithmic_bytes, cudaStream_t stream) : stream_(stream), idx_(-1) {
    count_launch();
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    ProfRec r{name, algorithmic_bytes, take_event(), take_event()};
    cudaEventRecord(r.e0, stream);
    idx_ = (int)g_prof.size();
    g_prof.push_back(r);
}

KernelScope::~KernelScope() {
    if (idx_ < 0) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    cudaEventRecord(g_prof[idx_].e1, stream_);
}

}  // namespace stb200

extern "C" {
void stb200_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(stb200::g_prof_mu);
    stb200::g_prof_on = on != 0;
}

// JSON: {"kernel": {"launches": n, "ms": total, "bytes": total algorithmic bytes}, ...}; clears the records.
size_t stb200_profile_dump(char *buf, size_t cap) {
    using namespace stb200;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (!buf || !cap) return 320 * (g_prof.size() + 1);  // size query only: upper bound, records are kept
    struct Agg { long long n = 0; double ms = 0, bytes = 0; };
    std::map<std::string, Agg> agg;
    for (auto &r : g_prof) {
        cudaEventSynchronize(r.e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.e0, r.e1);
        Agg &a = agg[r.name];
        a.n += 1; a.ms += ms; a.bytes += r.bytes;
        g_event_pool.push_back(r.e0);
        g_event_pool.push_back(r.e1);
    }
    g_prof.clear();
    std::string out = "{";
    bool first = true;
    for (auto &kv : agg) {
        char line[256];
        snprintf(line, sizeof(line), "%s\"%s\": {\"launches\": %lld, \"ms\": %.6f, \"bytes\": %.1f}", first ? "" : ", ",
                 kv.first.c_str(), kv.second.n, kv.second.ms, kv.second.bytes);
        out += line;
        first = false;
    }
    out += "}";
    if (buf && cap) {
        const size_t n = out.size() < cap - 1 ? out.size() : cap - 1;
        memcpy(buf, out.data(), n);
        buf[n] = 0;
    }
    return out.size() + 1;
}

const char *stb200_last_error(void) { return stb200::g_err; }
long long stb200_launch_count(void) { return stb200::g_launches.load(std::memory_order_relaxed); }
int stb200_version(void) { return 103; }   // 101: stb200_index gained len_order / t_len_order; 102: fused plan + fused attention; 103: qkv split / merge
}
