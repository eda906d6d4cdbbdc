// Error reporting and launch accounting shared by every entry point of libstb200.
#include <atomic>
#include <cstdarg>
#include <cstdio>

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

}  // namespace stb200

extern "C" {
const char *stb200_last_error(void) { return stb200::g_err; }
long long stb200_launch_count(void) { return stb200::g_launches.load(std::memory_order_relaxed); }
int stb200_version(void) { return 100; }
}
