// v1 operators: explicit, possibly unsorted (index0, index1) pair lists.  API completeness only
// (reachable from WindowAttention branches that no shipped config takes), so these are
// simple pair-parallel kernels, not tuned.  Semantics follow
//   /root/reference/lib/pointops2/src/attention/attention_cuda_kernel.cu:7-87
//   /root/reference/lib/pointops2/src/rpe/relative_pos_encoding_cuda_kernel.cu:7-118
// The reference runs one thread per (pair[, axis], head, channel) with a scalar atomicAdd each; here a group
// of d/4 lanes owns a (pair, head) item, reads 16 B per lane and issues one vector red.global.add.v4.f32
// per lane where a scatter is unavoidable (the pair list is unsorted, so there is no segment to own).
#include "common.cuh"

namespace stb200 {

constexpr int kLThreads = 256;

struct PairParams {
    int M, h, L;
    const float *X, *Y, *w, *T;
    const int *index0, *index1, *rel_idx;
    float *out0, *out1, *out2;
};

__device__ __forceinline__ void red_add4(float *dst, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

// table value for channels 4g..4g+3 of head hh straight from the [L,h,D,3] global table
template <int D>
__device__ __forceinline__ float4 table_sum_global(const float *T, const int *rel_idx, size_t m, int L, int h, int hh,
                                                   int g) {
    float e[4];
    int r[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) r[a] = min(max(__ldg(rel_idx + 3 * m + a), 0), L - 1);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int ch = 4 * g + c;
        e[c] = (__ldg(T + ((size_t)(r[0] * h + hh) * D + ch) * 3 + 0) + __ldg(T + ((size_t)(r[1] * h + hh) * D + ch) * 3 + 1)) +
               __ldg(T + ((size_t)(r[2] * h + hh) * D + ch) * 3 + 2);
    }
    return make_float4(e[0], e[1], e[2], e[3]);
}

enum PairOp { kStep1Fwd, kStep1Bwd, kStep2Fwd, kStep2Bwd, kRpeFwd, kRpeBwd, kStep2RpvFwd, kStep2RpvBwd };

template <int D, int OP>
__global__ void __launch_bounds__(kLThreads) pair_kernel(const PairParams p) {
    constexpr int G = D / 4;
    const int h = p.h, C = h * D;
    const long long items = (long long)p.M * h;
    const int g = threadIdx.x % G;
    for (long long e = (blockIdx.x * (long long)blockDim.x + threadIdx.x) / G; e < (items + kWarp / G - 1) / (kWarp / G) * (kWarp / G);
         e += (long long)gridDim.x * blockDim.x / G) {
        const bool active = e < items;
        const long long ee = active ? e : items - 1;
        const size_t m = (size_t)(ee / h);
        const int hh = (int)(ee % h);
        const size_t col = (size_t)hh * D + 4 * g;
        if (OP == kStep1Fwd) {  // attn[m,h] = <q[i0], k[i1]>
            const float4 a = ld_row4(p.X + (size_t)__ldg(p.index0 + m) * C + col);
            const float4 b = ld_row4(p.Y + (size_t)__ldg(p.index1 + m) * C + col);
            const float s = group_sum<G>(f4_dot(a, b, 0.f));
            if (active && g == 0) p.out0[m * h + hh] = s;
        } else if (OP == kStep1Bwd) {  // grad_q[i0] += g*k[i1]; grad_k[i1] += g*q[i0]
            const size_t i0 = __ldg(p.index0 + m), i1 = __ldg(p.index1 + m);
            const float gv = active ? __ldg(p.w + m * h + hh) : 0.f;
            const float4 a = ld_row4(p.X + i0 * C + col), b = ld_row4(p.Y + i1 * C + col);
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            if (active) {
                red_add4(p.out0 + i0 * C + col, f4_fma(gv, b, z));
                red_add4(p.out1 + i1 * C + col, f4_fma(gv, a, z));
            }
        } else if (OP == kStep2Fwd || OP == kStep2RpvFwd) {  // out[i0] += attn * (v[i1] (+ E))
            const size_t i0 = __ldg(p.index0 + m), i1 = __ldg(p.index1 + m);
            float4 val = ld_row4(p.Y + i1 * C + col);
            if (OP == kStep2RpvFwd) val = f4_add(table_sum_global<D>(p.T, p.rel_idx, m, p.L, h, hh, g), val);
            const float a = __ldg(p.w + m * h + hh);
            if (active) red_add4(p.out0 + i0 * C + col, f4_fma(a, val, make_float4(0.f, 0.f, 0.f, 0.f)));
        } else if (OP == kStep2Bwd || OP == kStep2RpvBwd) {  // grad_attn = <g[i0], v[i1] (+E)>; grad_v[i1] += attn*g[i0]; grad_T += attn*g
            const size_t i0 = __ldg(p.index0 + m), i1 = __ldg(p.index1 + m);
            const float4 go = ld_row4(p.X + i0 * C + col);
            float4 val = ld_row4(p.Y + i1 * C + col);
            if (OP == kStep2RpvBwd) val = f4_add(table_sum_global<D>(p.T, p.rel_idx, m, p.L, h, hh, g), val);
            const float s = group_sum<G>(f4_dot(go, val, 0.f));
            const float a = __ldg(p.w + m * h + hh);
            if (active) {
                if (g == 0) p.out0[m * h + hh] = s;
                const float4 c = f4_fma(a, go, make_float4(0.f, 0.f, 0.f, 0.f));
                red_add4(p.out1 + i1 * C + col, c);
                if (OP == kStep2RpvBwd) {
                    const float cc[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
                    for (int ax = 0; ax < 3; ++ax) {
                        const int r = min(max(__ldg(p.rel_idx + 3 * m + ax), 0), p.L - 1);
#pragma unroll
                        for (int ch = 0; ch < 4; ++ch)
                            atomicAdd(p.out2 + ((size_t)(r * h + hh) * D + 4 * g + ch) * 3 + ax, cc[ch]);
                    }
                }
            }
        } else if (OP == kRpeFwd) {  // out[m,h] = <x[index[m]], E(m)>
            const float4 x = ld_row4(p.X + (size_t)__ldg(p.index0 + m) * C + col);
            const float s = group_sum<G>(f4_dot(x, table_sum_global<D>(p.T, p.rel_idx, m, p.L, h, hh, g), 0.f));
            if (active && g == 0) p.out0[m * h + hh] = s;
        } else if (OP == kRpeBwd) {  // grad_x[index] += g*E ; grad_T[r_a,h,c,a] += g*x
            const size_t i0 = __ldg(p.index0 + m);
            const float gv = __ldg(p.w + m * h + hh);
            const float4 x = ld_row4(p.X + i0 * C + col);
            const float4 E = table_sum_global<D>(p.T, p.rel_idx, m, p.L, h, hh, g);
            if (active) {
                red_add4(p.out0 + i0 * C + col, f4_fma(gv, E, make_float4(0.f, 0.f, 0.f, 0.f)));
                const float cc[4] = {gv * x.x, gv * x.y, gv * x.z, gv * x.w};
#pragma unroll
                for (int ax = 0; ax < 3; ++ax) {
                    const int r = min(max(__ldg(p.rel_idx + 3 * m + ax), 0), p.L - 1);
#pragma unroll
                    for (int ch = 0; ch < 4; ++ch)
                        atomicAdd(p.out1 + ((size_t)(r * h + hh) * D + 4 * g + ch) * 3 + ax, cc[ch]);
                }
            }
        }
    }
}

template <int OP>
static int launch_pair(int D, const PairParams &p, cudaStream_t s) {
    if (p.M == 0) return STB200_OK;
    const long long threads = (long long)p.M * p.h * (D / 4);
    const int blocks = (int)max(1LL, min((threads + kLThreads - 1) / kLThreads, (long long)kNumSMs * 16));
    {
        KernelScope ks("pair_kernel[v1]", 0.0, s);
        if (D == 16) pair_kernel<16, OP><<<blocks, kLThreads, 0, s>>>(p);
        else pair_kernel<32, OP><<<blocks, kLThreads, 0, s>>>(p);
    }
    return check_launch("pair_kernel");
}

static int check_v1(int M, int h, int D) {
    STB200_REQUIRE(M >= 0 && h > 0, STB200_ERR_ARG, "bad sizes M=%d h=%d", M, h);
    STB200_REQUIRE(D == 16 || D == 32, STB200_ERR_HEAD_DIM, "d != 16 and d != 32 (got %d)", D);
    return STB200_OK;
}

}  // namespace stb200

using namespace stb200;

extern "C" {

int stb200_attention_step1_forward(int, int M, int h, int C, const float *q, const float *k, const int *index0,
                                   const int *index1, float *attn, void *stream) {
    if (int rc = check_v1(M, h, C / h)) return rc;
    PairParams p{};
    p.M = M; p.h = h; p.X = q; p.Y = k; p.index0 = index0; p.index1 = index1; p.out0 = attn;
    return launch_pair<kStep1Fwd>(C / h, p, (cudaStream_t)stream);
}

int stb200_attention_step1_backward(int, int M, int h, int C, const float *grad_out, const int *index0,
                                    const int *index1, const float *q, const float *k, float *grad_q, float *grad_k,
                                    void *stream) {
    if (int rc = check_v1(M, h, C / h)) return rc;
    PairParams p{};
    p.M = M; p.h = h; p.X = q; p.Y = k; p.w = grad_out; p.index0 = index0; p.index1 = index1;
    p.out0 = grad_q; p.out1 = grad_k;
    return launch_pair<kStep1Bwd>(C / h, p, (cudaStream_t)stream);
}

int stb200_attention_step2_forward(int, int M, int h, int C, const float *attn, const float *v, const int *index0,
                                   const int *index1, float *output, void *stream) {
    if (int rc = check_v1(M, h, C / h)) return rc;
    PairParams p{};
    p.M = M; p.h = h; p.Y = v; p.w = attn; p.index0 = index0; p.index1 = index1; p.out0 = output;
    return launch_pair<kStep2Fwd>(C / h, p, (cudaStream_t)stream);
}

int stb200_attention_step2_backward(int, int M, int h, int C, const float *grad_out, const int *index0,
                                    const int *index1, const float *attn, const float *v, float *grad_attn,
                                    float *grad_v, void *stream) {
    if (int rc = check_v1(M, h, C / h)) return rc;
    PairParams p{};
    p.M = M; p.h = h; p.X = grad_out; p.Y = v; p.w = attn; p.index0 = index0; p.index1 = index1;
    p.out0 = grad_attn; p.out1 = grad_v;
    return launch_pair<kStep2Bwd>(C / h, p, (cudaStream_t)stream);
}

int stb200_dot_prod_with_idx_forward(int, int M, int h, int hdim, int L, const float *q, const int *index,
                                     const float *table, const int *rel_idx, float *output, void *stream) {
    if (int rc = check_v1(M, h, hdim)) return rc;
    STB200_REQUIRE(L > 0, STB200_ERR_ARG, "L<=0");
    PairParams p{};
    p.M = M; p.h = h; p.L = L; p.X = q; p.T = table; p.index0 = index; p.rel_idx = rel_idx; p.out0 = output;
    return launch_pair<kRpeFwd>(hdim, p, (cudaStream_t)stream);
}

int stb200_dot_prod_with_idx_backward(int, int M, int h, int hdim, int L, const float *grad_out, const float *q,
                                      const int *index, const float *table, const int *rel_idx, float *grad_q,
                                      float *grad_table, void *stream) {
    if (int rc = check_v1(M, h, hdim)) return rc;
    STB200_REQUIRE(L > 0, STB200_ERR_ARG, "L<=0");
    PairParams p{};
    p.M = M; p.h = h; p.L = L; p.X = q; p.w = grad_out; p.T = table; p.index0 = index; p.rel_idx = rel_idx;
    p.out0 = grad_q; p.out1 = grad_table;
    return launch_pair<kRpeBwd>(hdim, p, (cudaStream_t)stream);
}

int stb200_attention_step2_with_rel_pos_value_forward(int, int M, int h, int hdim, int L, const float *attn,
                                                      const float *v, const int *index0, const int *index1,
                                                      const float *table, const int *rel_idx, float *output,
                                                      void *stream) {
    if (int rc = check_v1(M, h, hdim)) return rc;
    STB200_REQUIRE(L > 0, STB200_ERR_ARG, "L<=0");
    PairParams p{};
    p.M = M; p.h = h; p.L = L; p.Y = v; p.w = attn; p.T = table; p.index0 = index0; p.index1 = index1;
    p.rel_idx = rel_idx; p.out0 = output;
    return launch_pair<kStep2RpvFwd>(hdim, p, (cudaStream_t)stream);
}

int stb200_attention_step2_with_rel_pos_value_backward(int, int M, int h, int hdim, int L, const float *grad_out,
                                                       const int *index0, const int *index1, const float *attn,
                                                       const float *v, const float *table, const int *rel_idx,
                                                       float *grad_attn, float *grad_v, float *grad_table,
                                                       void *stream) {
    if (int rc = check_v1(M, h, hdim)) return rc;
    STB200_REQUIRE(L > 0, STB200_ERR_ARG, "L<=0");
    PairParams p{};
    p.M = M; p.h = h; p.L = L; p.X = grad_out; p.Y = v; p.w = attn; p.T = table; p.index0 = index0;
    p.index1 = index1; p.rel_idx = rel_idx; p.out0 = grad_attn; p.out1 = grad_v; p.out2 = grad_table;
    return launch_pair<kStep2RpvBwd>(hdim, p, (cudaStream_t)stream);
}

}  // extern "C"
