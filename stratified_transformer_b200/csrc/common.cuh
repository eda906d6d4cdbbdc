// Shared device/host helpers for libstb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/stb200.h"

namespace stb200 {

constexpr int kWarp = 32;
constexpr int kNumSMs = 148;  // B200; grids of persistent kernels are sized in multiples of this

void set_error(const char *fmt, ...);
void count_launch(int n = 1);
int check_launch(const char *what);  // cudaGetLastError -> STB200_ERR_CUDA

// Declared around every kernel launch: counts it and, when profiling is on, brackets it with CUDA events on the
// launching stream.  algorithmic_bytes = what this launch must read + write once ("Roofline accounting").
class KernelScope {
  public:
    KernelScope(const char *name, double algorithmic_bytes, cudaStream_t stream);
    ~KernelScope();
  private:
    cudaStream_t stream_;
    int idx_;
};

#define STB200_REQUIRE(cond, code, ...)        \
    do {                                       \
        if (!(cond)) {                         \
            ::stb200::set_error(__VA_ARGS__);  \
            return (code);                     \
        }                                      \
    } while (0)

// ---- loads -----------------------------------------------------------------------------------
// Pair-ordered arrays (attn, index1, rel_idx).  A lane group re-reads neighbouring words of the same sector
// (3 rel_idx ints, h attn floats), so they go through L1 like everything else; kept as a separate name so the
// cache policy of the streamed arrays can be changed in one place.
__device__ __forceinline__ float ld_stream(const float *p) { return __ldg(p); }
__device__ __forceinline__ int ld_stream(const int *p) { return __ldg(p); }
// Gathered rows (q/k/v/grad_out) are re-used across heads and neighbouring queries: read-only path, L1 allocate.
__device__ __forceinline__ float4 ld_row4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }

__device__ __forceinline__ float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4_fma(float s, float4 a, float4 c) {
    return make_float4(fmaf(s, a.x, c.x), fmaf(s, a.y, c.y), fmaf(s, a.z, c.z), fmaf(s, a.w, c.w));
}
__device__ __forceinline__ float f4_dot(float4 a, float4 b, float acc) {
    acc = fmaf(a.x, b.x, acc);
    acc = fmaf(a.y, b.y, acc);
    acc = fmaf(a.z, b.z, acc);
    return fmaf(a.w, b.w, acc);
}

template <int WIDTH>
__device__ __forceinline__ float group_sum(float v) {  // sum over aligned groups of WIDTH lanes
#pragma unroll
    for (int o = WIDTH / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

inline int largest_head_group(int h, int cap) {
    for (int g = cap; g > 1; --g)
        if (h % g == 0) return g;
    return 1;
}

}  // namespace stb200
