// The elementwise passes of WindowAttention.forward around the pair ops, fused into one kernel each way.
//
// Reference (/root/reference/model/stratified_transformer.py:172-175):
//     qkv = self.qkv(feats).reshape(N, 3, h, C // h).permute(1, 0, 2, 3).contiguous();  query = query * self.scale
// followed by `.float()` at every pointops call under AMP (lines 193-216).  As torch ops that is, per block, a bias add inside the
// GEMM epilogue, an N x 3C permute copy, a multiply, three casts forward and three casts + a concatenation + a bias
// reduction backward.  Here the projection GEMM runs without a bias and ONE kernel turns its [N, 3C] output (fp32, bf16 or
// fp16) into the three contiguous fp32 [N, h, d] operands (bias added, scale already folded into the weights by the caller);
// the backward kernel writes the [N, 3C] gradient in the GEMM's dtype and the bias gradient (column sums) in the same pass.
//
// HBM-bound streaming: 16 B loads / stores, grid = a multiple of the SM count, every byte touched once.
#include "common.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace stb200 {

template <typename T> struct Vec8;   // 8 consecutive elements <-> 8 floats
template <> struct Vec8<float> {
    static __device__ __forceinline__ void load(const float *p, float (&x)[8]) {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(p)), b = __ldg(reinterpret_cast<const float4 *>(p) + 1);
        x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
    }
    static __device__ __forceinline__ void store(float *p, const float (&x)[8]) {
        reinterpret_cast<float4 *>(p)[0] = make_float4(x[0], x[1], x[2], x[3]);
        reinterpret_cast<float4 *>(p)[1] = make_float4(x[4], x[5], x[6], x[7]);
    }
};
template <> struct Vec8<__nv_bfloat16> {
    static __device__ __forceinline__ void load(const __nv_bfloat16 *p, float (&x)[8]) {
        const uint4 u = __ldg(reinterpret_cast<const uint4 *>(p));
        const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&u);
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); x[2 * i] = f.x; x[2 * i + 1] = f.y; }
    }
    static __device__ __forceinline__ void store(__nv_bfloat16 *p, const float (&x)[8]) {
        uint4 u;
        __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&u);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(x[2 * i], x[2 * i + 1]);
        *reinterpret_cast<uint4 *>(p) = u;
    }
};
template <> struct Vec8<__half> {
    static __device__ __forceinline__ void load(const __half *p, float (&x)[8]) {
        const uint4 u = __ldg(reinterpret_cast<const uint4 *>(p));
        const __half2 *h = reinterpret_cast<const __half2 *>(&u);
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(h[i]); x[2 * i] = f.x; x[2 * i + 1] = f.y; }
    }
    static __device__ __forceinline__ void store(__half *p, const float (&x)[8]) {
        uint4 u;
        __half2 *h = reinterpret_cast<__half2 *>(&u);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(x[2 * i], x[2 * i + 1]);
        *reinterpret_cast<uint4 *>(p) = u;
    }
};

constexpr int kQkvThreads = 256;

// Work item = 8 consecutive columns of one row of the [N, 3C] matrix; G = 3C/8 items per row.  A thread keeps its column
// group for the whole kernel (the item stride is a multiple of G), so the backward kernel can sum its columns in registers.
template <typename T>
__global__ void __launch_bounds__(kQkvThreads) qkv_split_kernel(int N, int C, const T *__restrict__ qkv, const float *__restrict__ bias,
                                                                float *__restrict__ q, float *__restrict__ k, float *__restrict__ v,
                                                                int rows_per_step) {
    const int G = 3 * C / 8;
    const int cg = threadIdx.x % G, lr = threadIdx.x / G;   // threads beyond rows_per_step * G idle (G need not divide 256)
    if (lr >= rows_per_step) return;
    const int col = cg * 8, part = col / C, pc = col - part * C;
    float *dst = part == 0 ? q : (part == 1 ? k : v);
    float b[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (bias) Vec8<float>::load(bias + col, b);
    for (long long r = (long long)blockIdx.x * rows_per_step + lr; r < N; r += (long long)gridDim.x * rows_per_step) {
        float x[8];
        Vec8<T>::load(qkv + r * 3 * C + col, x);
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] += b[i];
        Vec8<float>::store(dst + r * C + pc, x);
    }
}

template <typename T>
__global__ void __launch_bounds__(kQkvThreads) qkv_merge_kernel(int N, int C, const float *__restrict__ gq, const float *__restrict__ gk,
                                                                const float *__restrict__ gv, T *__restrict__ g_qkv,
                                                                float *__restrict__ bias_partial, int rows_per_step) {
    __shared__ float red[kQkvThreads * 8];
    const int G = 3 * C / 8;
    const int cg = threadIdx.x % G, lr = threadIdx.x / G;
    const bool active = lr < rows_per_step;
    const int col = cg * 8, part = col / C, pc = col - part * C;
    const float *src = part == 0 ? gq : (part == 1 ? gk : gv);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (active) {
        for (long long r = (long long)blockIdx.x * rows_per_step + lr; r < N; r += (long long)gridDim.x * rows_per_step) {
            float x[8];
            Vec8<float>::load(src + r * C + pc, x);
            Vec8<T>::store(g_qkv + r * 3 * C + col, x);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] += x[i];
        }
    }
    if (!bias_partial) return;
    // deterministic column sums: per-CTA partial [3C] (rows of this CTA in a fixed order), the caller adds the partials
#pragma unroll
    for (int i = 0; i < 8; ++i) red[threadIdx.x * 8 + i] = active ? acc[i] : 0.f;
    __syncthreads();
    for (int c = threadIdx.x; c < 3 * C; c += kQkvThreads) {
        const int g = c / 8, i = c % 8;
        float s = 0.f;
        for (int l = 0; l < rows_per_step; ++l) s += red[(l * G + g) * 8 + i];
        bias_partial[(size_t)blockIdx.x * 3 * C + c] = s;
    }
}

static int qkv_grid(int N, int rows_per_step) {
    const long long need = ((long long)N + rows_per_step - 1) / rows_per_step;
    const long long cap = 8LL * kNumSMs;   // 8 resident CTAs of 256 threads per SM
    return (int)(need < cap ? (need < 1 ? 1 : need) : cap);
}


// ---- LayerNorm over short rows -------------------------------------------------------------------------------------------
// nn.LayerNorm(C) on [N, C] with C = 48 ... 384 is what every block of the model runs twice (SwinTransformerBlock.norm1 / norm2,
// model/stratified_transformer.py:227,233) plus TransitionDown / Upsample; torch's kernel gives a whole warp (or more) to each
// 48-float row and reaches ~0.6 TB/s forward, less backward (12.8 + 15.3 ms per step of the full model on 8 x 80k points).
// Here a group of G = 8 / 16 / 32 lanes owns a row (at most 12 elements per lane, kept in registers), statistics by shuffles
// inside the group, and the backward pass accumulates the gamma / beta gradients of its columns in registers across all the
// rows a thread sees, reduces them through shared memory once per CTA and writes per-CTA partials (deterministic).
constexpr int kLnThreads = 256;
constexpr int kLnMaxPer = 12;   // elements per lane

template <int G>
__device__ __forceinline__ float group_sum_ln(float v) {
#pragma unroll
    for (int o = G / 2; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int G>
__global__ void __launch_bounds__(kLnThreads) layer_norm_fwd_kernel(long long N, int C, float eps, const float *__restrict__ x,
                                                                    const float *__restrict__ gamma, const float *__restrict__ beta,
                                                                    float *__restrict__ y, float *__restrict__ mean, float *__restrict__ rstd) {
    const int lane = threadIdx.x % G, grp = threadIdx.x / G, groups = kLnThreads / G;
    const int per = (C + G - 1) / G;
    float gm[kLnMaxPer], bt[kLnMaxPer];
#pragma unroll
    for (int t = 0; t < kLnMaxPer; ++t) {
        const int c = lane + t * G;
        gm[t] = (t < per && c < C && gamma) ? __ldg(gamma + c) : 1.f;
        bt[t] = (t < per && c < C && beta) ? __ldg(beta + c) : 0.f;
    }
    const float inv_c = 1.0f / (float)C;
    // the row base is CTA-uniform, so every lane of a warp runs the same number of iterations and takes part in the shuffles; rows
    // past the end are masked (a loop bound per group would leave the full-mask shuffles of the other groups waiting forever)
    for (long long base = (long long)blockIdx.x * groups; base < N; base += (long long)gridDim.x * groups) {
        const long long r = base + grp;
        const bool row_on = r < N;
        float v[kLnMaxPer];
        float s = 0.f;
#pragma unroll
        for (int t = 0; t < kLnMaxPer; ++t) {
            const int c = lane + t * G;
            v[t] = (row_on && t < per && c < C) ? __ldg(x + r * C + c) : 0.f;
            s += v[t];
        }
        const float mu = group_sum_ln<G>(s) * inv_c;
        float q = 0.f;
#pragma unroll
        for (int t = 0; t < kLnMaxPer; ++t) {
            const int c = lane + t * G;
            const float d = (t < per && c < C) ? v[t] - mu : 0.f;
            q = fmaf(d, d, q);
        }
        const float rs = rsqrtf(group_sum_ln<G>(q) * inv_c + eps);
#pragma unroll
        for (int t = 0; t < kLnMaxPer; ++t) {
            const int c = lane + t * G;
            if (row_on && t < per && c < C) y[r * C + c] = fmaf((v[t] - mu) * rs, gm[t], bt[t]);
        }
        if (row_on && lane == 0) { mean[r] = mu; rstd[r] = rs; }
    }
}

template <int G>
__global__ void __launch_bounds__(kLnThreads) layer_norm_bwd_kernel(long long N, int C, const float *__restrict__ g, const float *__restrict__ x,
                                                                    const float *__restrict__ gamma, const float *__restrict__ mean,
                                                                    const float *__restrict__ rstd, float *__restrict__ gx,
                                                                    float *__restrict__ partial /* [grid, 2C]: dgamma | dbeta */) {
    extern __shared__ float ln_red[];   // [groups][2C]
    const int lane = threadIdx.x % G, grp = threadIdx.x / G, groups = kLnThreads / G;
    const int per = (C + G - 1) / G;
    float gm[kLnMaxPer], dg[kLnMaxPer], db[kLnMaxPer];
#pragma unroll
    for (int t = 0; t < kLnMaxPer; ++t) {
        const int c = lane + t * G;
        gm[t] = (t < per && c < C && gamma) ? __ldg(gamma + c) : 1.f;
        dg[t] = 0.f;
        db[t] = 0.f;
    }
    const float inv_c = 1.0f / (float)C;
    for (long long base = (long long)blockIdx.x * groups; base < N; base += (long long)gridDim.x * groups) {   // uniform, see forward
        const long long r = base + grp;
        const bool row_on = r < N;
        const float mu = row_on ? __ldg(mean + r) : 0.f, rs = row_on ? __ldg(rstd + r) : 0.f;
        float xh[kLnMaxPer], a[kLnMaxPer];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int t = 0; t < kLnMaxPer; ++t) {
            const int c = lane + t * G;
            const bool on = row_on && t < per && c < C;
            const float gv = on ? __ldg(g + r * C + c) : 0.f;
            xh[t] = on ? (__ldg(x + r * C + c) - mu) * rs : 0.f;
            a[t] = gv * gm[t];
            s1 += a[t];
            s2 = fmaf(a[t], xh[t], s2);
            dg[t] = fmaf(gv, xh[t], dg[t]);
            db[t] += gv;
        }
        s1 = group_sum_ln<G>(s1) * inv_c;
        s2 = group_sum_ln<G>(s2) * inv_c;
#pragma unroll
        for (int t = 0; t < kLnMaxPer; ++t) {
            const int c = lane + t * G;
            if (row_on && t < per && c < C) gx[r * C + c] = rs * (a[t] - s1 - xh[t] * s2);
        }
    }
    if (!partial) return;
#pragma unroll
    for (int t = 0; t < kLnMaxPer; ++t) {
        const int c = lane + t * G;
        if (t < per && c < C) {
            ln_red[(size_t)grp * 2 * C + c] = dg[t];
            ln_red[(size_t)grp * 2 * C + C + c] = db[t];
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < 2 * C; c += kLnThreads) {
        float s = 0.f;
        for (int gI = 0; gI < groups; ++gI) s += ln_red[(size_t)gI * 2 * C + c];
        partial[(size_t)blockIdx.x * 2 * C + c] = s;
    }
}

static int ln_group(int C) { return C <= 96 ? (C <= 48 ? 8 : 16) : 32; }   // <= 12 elements per lane up to C = 384
static int ln_grid(long long N, int G) {
    const long long need = (N + kLnThreads / G - 1) / (kLnThreads / G);
    const long long cap = 8LL * kNumSMs;
    return (int)(need < cap ? (need < 1 ? 1 : need) : cap);
}


// ---- KPConv neighbourhood aggregation (stem, outside the hot path) ---------------------------------------------------------
// Rigid kernel-point convolution, linear influence, sum aggregation (torch_points3d KPConvLayer, third party: DESIGN.md section 4):
//     weighted[i, k, :] = sum_j max(0, 1 - |x_{n(i,j)} - x_i - K_k| / extent) * f[n(i,j), :]          (this kernel)
//     out[i, :]         = sum_k weighted[i, k, :] W_k                                                  (one GEMM, library)
// As torch operators the first line materialises [n, 34, 15, 3] differences and [n, 34, 15] weights (several GB at 640k points)
// forward and backward.  Here a warp takes a query point: lane = neighbour slot (34 slots: two rounds), each lane computes its
// neighbour's 15 influences, and the [15, C] sums are reduced over the lanes with shuffles.  Neighbour index < 0 or >= n_support
// is the radius search's padding: no contribution.  Backward scatters grad_f with vector-free float atomics (order-dependent sums
// in the last bits; the stem is not a parity surface).
constexpr int kKpMaxK = 16;   // kernel points
constexpr int kKpMaxC = 16;   // input channels handled by this kernel (the stem has 6 and 12)

template <bool BWD>
__global__ void __launch_bounds__(256) kpconv_weighted_kernel(int n, int n_sup, int nn, int K, int C, float inv_extent,
                                                              const float *__restrict__ q_xyz, const float *__restrict__ s_xyz,
                                                              const long long *__restrict__ nbr, const float *__restrict__ kpts,
                                                              const float *__restrict__ feats, float *__restrict__ weighted,
                                                              const float *__restrict__ g_weighted, float *__restrict__ g_feats) {
    __shared__ float skp[kKpMaxK * 3];
    if (threadIdx.x < K * 3) skp[threadIdx.x] = kpts[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x % 32, wpb = blockDim.x / 32;
    for (long long i = (long long)blockIdx.x * wpb + threadIdx.x / 32; i < n; i += (long long)gridDim.x * wpb) {   // warp-uniform
        const float qx = __ldg(q_xyz + i * 3), qy = __ldg(q_xyz + i * 3 + 1), qz = __ldg(q_xyz + i * 3 + 2);
        float acc[BWD ? 1 : kKpMaxC];   // forward: this lane's share of one kernel point's [C] sum, reduced below
        for (int k = 0; k < K; ++k) {
            if (!BWD)
#pragma unroll
                for (int c = 0; c < kKpMaxC; ++c) acc[c] = 0.f;
            for (int j = lane; j < nn; j += 32) {
                const long long p = __ldg(nbr + i * nn + j);
                if (p < 0 || p >= n_sup) continue;
                const float dx = __ldg(s_xyz + p * 3) - qx - skp[k * 3], dy = __ldg(s_xyz + p * 3 + 1) - qy - skp[k * 3 + 1],
                            dz = __ldg(s_xyz + p * 3 + 2) - qz - skp[k * 3 + 2];
                const float w = fmaxf(1.f - sqrtf(dx * dx + dy * dy + dz * dz) * inv_extent, 0.f);
                if (w == 0.f) continue;
                if (!BWD) {
#pragma unroll
                    for (int c = 0; c < kKpMaxC; ++c)
                        if (c < C) acc[c] = fmaf(w, __ldg(feats + p * C + c), acc[c]);
                } else {
                    for (int c = 0; c < C; ++c) atomicAdd(g_feats + p * C + c, w * __ldg(g_weighted + (i * K + k) * C + c));
                }
            }
            if (!BWD) {
#pragma unroll
                for (int c = 0; c < kKpMaxC; ++c) {
                    if (c < C) {   // C is warp-uniform
                        float v = acc[c];
#pragma unroll
                        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                        if (lane == 0) weighted[(i * K + k) * C + c] = v;
                    }
                }
            }
        }
    }
}

}  // namespace stb200

using namespace stb200;

extern "C" {

int stb200_qkv_partial_rows(int N, int C) {
    if (N <= 0 || C <= 0 || C % 8 || 3 * C / 8 > kQkvThreads) return 0;
    return qkv_grid(N, kQkvThreads / (3 * C / 8));
}

int stb200_qkv_split(int N, int C, int dtype, const void *qkv, const float *bias, float *q, float *k, float *v, void *stream) {
    STB200_REQUIRE(N > 0 && C > 0 && C % 8 == 0 && 3 * C / 8 <= kQkvThreads, STB200_ERR_ARG,
                   "qkv_split: C must be a multiple of 8 with 3C <= %d (got %d)", 8 * kQkvThreads, C);
    STB200_REQUIRE(qkv && q && k && v, STB200_ERR_ARG, "null pointer");
    const int rps = kQkvThreads / (3 * C / 8), grid = qkv_grid(N, rps);
    cudaStream_t s = (cudaStream_t)stream;
    const double esz = dtype == 0 ? 4.0 : 2.0;
    KernelScope ks("qkv_split", (double)N * 3 * C * (esz + 4.0), s);
    if (dtype == 0) qkv_split_kernel<float><<<grid, kQkvThreads, 0, s>>>(N, C, (const float *)qkv, bias, q, k, v, rps);
    else if (dtype == 1) qkv_split_kernel<__nv_bfloat16><<<grid, kQkvThreads, 0, s>>>(N, C, (const __nv_bfloat16 *)qkv, bias, q, k, v, rps);
    else if (dtype == 2) qkv_split_kernel<__half><<<grid, kQkvThreads, 0, s>>>(N, C, (const __half *)qkv, bias, q, k, v, rps);
    else { set_error("qkv_split: dtype %d (0 = fp32, 1 = bf16, 2 = fp16)", dtype); return STB200_ERR_ARG; }
    return check_launch("qkv_split");
}

int stb200_qkv_merge(int N, int C, int dtype, const float *gq, const float *gk, const float *gv, void *g_qkv, float *bias_partial,
                     void *stream) {
    STB200_REQUIRE(N > 0 && C > 0 && C % 8 == 0 && 3 * C / 8 <= kQkvThreads, STB200_ERR_ARG,
                   "qkv_merge: C must be a multiple of 8 with 3C <= %d (got %d)", 8 * kQkvThreads, C);
    STB200_REQUIRE(gq && gk && gv && g_qkv, STB200_ERR_ARG, "null pointer");
    const int rps = kQkvThreads / (3 * C / 8), grid = qkv_grid(N, rps);
    cudaStream_t s = (cudaStream_t)stream;
    const double esz = dtype == 0 ? 4.0 : 2.0;
    KernelScope ks("qkv_merge", (double)N * 3 * C * (esz + 4.0), s);
    if (dtype == 0) qkv_merge_kernel<float><<<grid, kQkvThreads, 0, s>>>(N, C, gq, gk, gv, (float *)g_qkv, bias_partial, rps);
    else if (dtype == 1) qkv_merge_kernel<__nv_bfloat16><<<grid, kQkvThreads, 0, s>>>(N, C, gq, gk, gv, (__nv_bfloat16 *)g_qkv, bias_partial, rps);
    else if (dtype == 2) qkv_merge_kernel<__half><<<grid, kQkvThreads, 0, s>>>(N, C, gq, gk, gv, (__half *)g_qkv, bias_partial, rps);
    else { set_error("qkv_merge: dtype %d (0 = fp32, 1 = bf16, 2 = fp16)", dtype); return STB200_ERR_ARG; }
    return check_launch("qkv_merge");
}

}  // extern "C"

extern "C" {

int stb200_layer_norm_partial_rows(long long N, int C) {
    if (N <= 0 || C <= 0 || C > 32 * kLnMaxPer) return 0;
    return ln_grid(N, ln_group(C));
}

int stb200_layer_norm_forward(long long N, int C, float eps, const float *x, const float *gamma, const float *beta, float *y, float *mean,
                              float *rstd, void *stream) {
    STB200_REQUIRE(N > 0 && C > 0 && C <= 32 * kLnMaxPer, STB200_ERR_ARG, "layer_norm: C must be in 1..%d (got %d)", 32 * kLnMaxPer, C);
    STB200_REQUIRE(x && y && mean && rstd, STB200_ERR_ARG, "null pointer");
    const int G = ln_group(C), grid = ln_grid(N, G);
    cudaStream_t s = (cudaStream_t)stream;
    KernelScope ks("layer_norm_fwd", 8.0 * N * C + 8.0 * N, s);
    if (G == 8) layer_norm_fwd_kernel<8><<<grid, kLnThreads, 0, s>>>(N, C, eps, x, gamma, beta, y, mean, rstd);
    else if (G == 16) layer_norm_fwd_kernel<16><<<grid, kLnThreads, 0, s>>>(N, C, eps, x, gamma, beta, y, mean, rstd);
    else layer_norm_fwd_kernel<32><<<grid, kLnThreads, 0, s>>>(N, C, eps, x, gamma, beta, y, mean, rstd);
    return check_launch("layer_norm_fwd");
}

int stb200_layer_norm_backward(long long N, int C, const float *grad_y, const float *x, const float *gamma, const float *mean,
                               const float *rstd, float *grad_x, float *partial, void *stream) {
    STB200_REQUIRE(N > 0 && C > 0 && C <= 32 * kLnMaxPer, STB200_ERR_ARG, "layer_norm: C must be in 1..%d (got %d)", 32 * kLnMaxPer, C);
    STB200_REQUIRE(grad_y && x && mean && rstd && grad_x, STB200_ERR_ARG, "null pointer");
    const int G = ln_group(C), grid = ln_grid(N, G);
    const size_t smem = (size_t)(kLnThreads / G) * 2 * C * sizeof(float);
    cudaStream_t s = (cudaStream_t)stream;
    KernelScope ks("layer_norm_bwd", 12.0 * N * C + 8.0 * N, s);
    if (G == 8) layer_norm_bwd_kernel<8><<<grid, kLnThreads, smem, s>>>(N, C, grad_y, x, gamma, mean, rstd, grad_x, partial);
    else if (G == 16) layer_norm_bwd_kernel<16><<<grid, kLnThreads, smem, s>>>(N, C, grad_y, x, gamma, mean, rstd, grad_x, partial);
    else layer_norm_bwd_kernel<32><<<grid, kLnThreads, smem, s>>>(N, C, grad_y, x, gamma, mean, rstd, grad_x, partial);
    return check_launch("layer_norm_bwd");
}

}  // extern "C"

extern "C" int stb200_kpconv_weighted(int n, int n_sup, int nn, int K, int C, float extent, const float *q_xyz, const float *s_xyz,
                                      const long long *nbr, const float *kpts, const float *feats, float *weighted, void *stream) {
    STB200_REQUIRE(n >= 0 && n_sup > 0 && nn > 0 && K > 0 && K <= kKpMaxK && C > 0 && C <= kKpMaxC && extent > 0.f, STB200_ERR_ARG,
                   "kpconv: K <= %d kernel points, C <= %d channels (got %d, %d)", kKpMaxK, kKpMaxC, K, C);
    if (n == 0) return STB200_OK;
    STB200_REQUIRE(q_xyz && s_xyz && nbr && kpts && feats && weighted, STB200_ERR_ARG, "null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    KernelScope ks("kpconv_weighted_fwd", 0.0, s);
    const int grid = (int)min((long long)kNumSMs * 8, ((long long)n + 7) / 8);
    kpconv_weighted_kernel<false><<<grid, 256, 0, s>>>(n, n_sup, nn, K, C, 1.0f / extent, q_xyz, s_xyz, nbr, kpts, feats, weighted, nullptr, nullptr);
    return check_launch("kpconv_weighted");
}

extern "C" int stb200_kpconv_weighted_backward(int n, int n_sup, int nn, int K, int C, float extent, const float *q_xyz, const float *s_xyz,
                                               const long long *nbr, const float *kpts, const float *grad_weighted, float *grad_feats,
                                               void *stream) {
    STB200_REQUIRE(n >= 0 && n_sup > 0 && nn > 0 && K > 0 && K <= kKpMaxK && C > 0 && C <= kKpMaxC && extent > 0.f, STB200_ERR_ARG,
                   "kpconv: K <= %d kernel points, C <= %d channels (got %d, %d)", kKpMaxK, kKpMaxC, K, C);
    if (n == 0) return STB200_OK;
    STB200_REQUIRE(q_xyz && s_xyz && nbr && kpts && grad_weighted && grad_feats, STB200_ERR_ARG, "null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    KernelScope ks("kpconv_weighted_bwd", 0.0, s);
    const int grid = (int)min((long long)kNumSMs * 8, ((long long)n + 7) / 8);
    kpconv_weighted_kernel<true><<<grid, 256, 0, s>>>(n, n_sup, nn, K, C, 1.0f / extent, q_xyz, s_xyz, nbr, kpts, nullptr, nullptr, grad_weighted,
                                                     grad_feats);
    return check_launch("kpconv_weighted_backward");
}
