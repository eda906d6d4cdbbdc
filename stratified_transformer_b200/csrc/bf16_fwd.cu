// bf16-storage forward path (inference; BASELINE config 3 "ScanNetv2 inference ... bf16").
//
// Same math as the fp32 per-pair forward kernels (seg_dot logits, seg_reduce aggregation in attention_ops.cu), with
// q / k / v rows and the staged rel-pos tables held as bf16: a head row is 32 B instead of 64 B, a table look-up is
// an 8-byte LDS per lane instead of 16, i.e. both limiters of the fp32 kernels (L1 wavefronts of the row gathers,
// shared-memory bandwidth of the table look-ups) are halved.  Products and sums are fp32; the logits, the softmax
// and the output stay fp32.  Stated tolerance against the fp32 oracle: 2e-2 of the output scale (bf16 has 8 bits of
// mantissa).  The tables arrive as the fp32 parameters and are rounded once while being staged.
// Forward only: training keeps the fp32 path.
#include <cuda_bf16.h>

#include "common.cuh"

namespace stb200 {

constexpr int kBThreads = 512;
constexpr int kBD = 16;            // head dim
constexpr int kBG = kBD / 4;       // lanes per (pair, head) item: 4 channels each
constexpr int kBNS = kWarp / kBG;  // items per warp step

struct BfParams {
    int N, h, L;
    const __nv_bfloat16 *q, *k, *v;   // [N, h, 16]
    const float *tq, *tk, *tv;        // [L, h, 16, 3] fp32 parameters
    const int *offsets, *index1, *row_order;
    const unsigned *packed;           // [M] bins, 10 bits each
    const float *attn;                // [M, h] probabilities (aggregate)
    float *out;                       // logits [M, h] or output [N, h, 16]
};

__device__ __forceinline__ float4 bf4_to_f4(uint2 u) {
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162 *>(&u.x), b = *reinterpret_cast<const __nv_bfloat162 *>(&u.y);
    const float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ float4 ld_row4_bf(const __nv_bfloat16 *p) { return bf4_to_f4(__ldg(reinterpret_cast<const uint2 *>(p))); }

// one head group of a [L,h,16,3] fp32 table -> shared memory [axis][l][hh][c] as bf16
template <int HG>
__device__ __forceinline__ void stage_table_bf(__nv_bfloat16 *dst, const float *__restrict__ src, int L, int h, int h0) {
    const int total = 3 * L * HG * kBD;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int c = i % kBD, hh = (i / kBD) % HG, l = (i / (kBD * HG)) % L, a = i / (kBD * HG * L);
        dst[i] = __float2bfloat16_rn(__ldg(src + ((size_t)(l * h + h0 + hh) * kBD + c) * 3 + a));
    }
}
template <int HG>
__device__ __forceinline__ float4 table_sum4_bf(const __nv_bfloat16 *ts, int L, int r0, int r1, int r2, int hh, int g) {
    const uint2 *t = reinterpret_cast<const uint2 *>(ts);
    const float4 a = bf4_to_f4(t[((0 * L + r0) * HG + hh) * kBG + g]);
    const float4 b = bf4_to_f4(t[((1 * L + r1) * HG + hh) * kBG + g]);
    const float4 c = bf4_to_f4(t[((2 * L + r2) * HG + hh) * kBG + g]);
    return f4_add(f4_add(a, b), c);
}

// logits[m, h] = <q, k> + <q, Eq(m)> + <k, Ek(m)>
template <int HG>
__global__ void __launch_bounds__(kBThreads) bf16_logits_kernel(const BfParams p) {
    extern __shared__ float4 smem4[];
    __nv_bfloat16 *tx = reinterpret_cast<__nv_bfloat16 *>(smem4);
    const int L = p.L, h = p.h, C = h * kBD, h0 = blockIdx.y * HG, tsz = 3 * L * HG * kBD;
    __nv_bfloat16 *ty = tx + tsz;
    float4 *xs = reinterpret_cast<float4 *>(ty + tsz);
    stage_table_bf<HG>(tx, p.tq, L, h, h0);
    stage_table_bf<HG>(ty, p.tk, L, h, h0);
    __syncthreads();
    const int warp = threadIdx.x / kWarp, lane = threadIdx.x % kWarp, nwarps = blockDim.x / kWarp;
    const int grp = lane / kBG, g = lane % kBG;
    float4 *xw = xs + warp * HG * kBG;
    for (int base_n = blockIdx.x * 64; base_n < p.N; base_n += gridDim.x * 64) {
        const int end_n = min(p.N, base_n + 64);
        for (int nn = base_n + warp; nn < end_n; nn += nwarps) {
            const int n = p.row_order ? __ldg(p.row_order + nn) : nn;
            const int start = __ldg(p.offsets + n), len = __ldg(p.offsets + n + 1) - start;
            if (len <= 0) continue;
            __syncwarp();
            if (lane < HG * kBG) xw[lane] = ld_row4_bf(p.q + (size_t)n * C + h0 * kBD + 4 * lane);
            __syncwarp();
            for (int c0 = 0; c0 < len; c0 += kWarp) {
                const int cnt = min(kWarp, len - c0);
                const int mt = start + c0 + min(lane, cnt - 1);
                const int j_l = __ldg(p.index1 + mt);
                const unsigned pk_l = __ldg(p.packed + mt);
                const int items = cnt * HG;
                for (int e0 = 0; e0 < items; e0 += kBNS * 4) {
                    float4 y4[4];
                    int pl[4], hh[4];
                    bool act[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int e = e0 + u * kBNS + grp;
                        act[u] = e < items;
                        const int ee = act[u] ? e : items - 1;
                        pl[u] = ee / HG;
                        hh[u] = ee - pl[u] * HG;
                        const int j = __shfl_sync(0xffffffffu, j_l, pl[u]);
                        y4[u] = ld_row4_bf(p.k + (size_t)j * C + (h0 + hh[u]) * kBD + 4 * g);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (e0 + u * kBNS >= items) break;
                        const float4 x4 = xw[hh[u] * kBG + g];
                        const unsigned pk = __shfl_sync(0xffffffffu, pk_l, pl[u]);
                        const int r0 = min((int)(pk & 0x3ff), L - 1), r1 = min((int)((pk >> 10) & 0x3ff), L - 1), r2 = min((int)(pk >> 20), L - 1);
                        float acc = f4_dot(x4, y4[u], 0.f);
                        acc = f4_dot(x4, table_sum4_bf<HG>(tx, L, r0, r1, r2, hh[u], g), acc);
                        acc = f4_dot(y4[u], table_sum4_bf<HG>(ty, L, r0, r1, r2, hh[u], g), acc);
                        acc = group_sum<kBG>(acc);
                        if (act[u] && g == 0) p.out[(size_t)(start + c0 + pl[u]) * h + h0 + hh[u]] = acc;
                    }
                }
            }
        }
    }
}

// out[n, h, :] = sum_seg attn[m, h] * (v[i1[m], h, :] + Ev(m, h, :))
template <int HG>
__global__ void __launch_bounds__(kBThreads) bf16_aggregate_kernel(const BfParams p) {
    extern __shared__ float4 smem4[];
    __nv_bfloat16 *ts = reinterpret_cast<__nv_bfloat16 *>(smem4);
    const int L = p.L, h = p.h, C = h * kBD, h0 = blockIdx.y * HG;
    stage_table_bf<HG>(ts, p.tv, L, h, h0);
    __syncthreads();
    const int warp = threadIdx.x / kWarp, lane = threadIdx.x % kWarp, nwarps = blockDim.x / kWarp;
    const int slot = lane / kBG, g = lane % kBG;
    for (int base_n = blockIdx.x * 64; base_n < p.N; base_n += gridDim.x * 64) {
        const int end_n = min(p.N, base_n + 64);
        for (int nn = base_n + warp; nn < end_n; nn += nwarps) {
            const int n = p.row_order ? __ldg(p.row_order + nn) : nn;
            const int start = __ldg(p.offsets + n), end = __ldg(p.offsets + n + 1);
            float4 acc[HG];
#pragma unroll
            for (int hh = 0; hh < HG; ++hh) acc[hh] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int c0 = start; c0 < end; c0 += kWarp) {
                const int cnt = min(kWarp, end - c0);
                const int tl = c0 + min(lane, cnt - 1);
                const int j_l = __ldg(p.index1 + tl);
                const unsigned pk_l = __ldg(p.packed + tl);
                for (int s0 = 0; s0 < cnt; s0 += kBNS * 2) {
                    float4 val[2][HG];
                    float wv[2][HG];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int pl = s0 + u * kBNS + slot;
                        const bool act = pl < cnt;
                        const int pc = act ? pl : cnt - 1;
                        const int j = __shfl_sync(0xffffffffu, j_l, pc);
                        const unsigned pk = __shfl_sync(0xffffffffu, pk_l, pc);
                        const int r0 = min((int)(pk & 0x3ff), L - 1), r1 = min((int)((pk >> 10) & 0x3ff), L - 1), r2 = min((int)(pk >> 20), L - 1);
#pragma unroll
                        for (int hh = 0; hh < HG; ++hh) {
                            wv[u][hh] = act ? __ldg(p.attn + (size_t)(c0 + pc) * h + h0 + hh) : 0.f;
                            val[u][hh] = f4_add(table_sum4_bf<HG>(ts, L, r0, r1, r2, hh, g),
                                                ld_row4_bf(p.v + (size_t)j * C + (h0 + hh) * kBD + 4 * g));
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 2; ++u)
#pragma unroll
                        for (int hh = 0; hh < HG; ++hh) acc[hh] = f4_fma(wv[u][hh], val[u][hh], acc[hh]);
                }
            }
#pragma unroll
            for (int hh = 0; hh < HG; ++hh) {
#pragma unroll
                for (int o = kBG; o < kWarp; o <<= 1) {
                    acc[hh].x += __shfl_xor_sync(0xffffffffu, acc[hh].x, o);
                    acc[hh].y += __shfl_xor_sync(0xffffffffu, acc[hh].y, o);
                    acc[hh].z += __shfl_xor_sync(0xffffffffu, acc[hh].z, o);
                    acc[hh].w += __shfl_xor_sync(0xffffffffu, acc[hh].w, o);
                }
            }
            if (slot == 0) {
#pragma unroll
                for (int hh = 0; hh < HG; ++hh)
                    *reinterpret_cast<float4 *>(p.out + (size_t)n * C + (h0 + hh) * kBD + 4 * g) = acc[hh];
            }
        }
    }
}

template <int HG>
static int launch_bf16(const BfParams &p, int M, bool aggregate, cudaStream_t s) {
    const size_t tsz = (size_t)3 * p.L * HG * kBD * sizeof(__nv_bfloat16);
    const size_t smem = aggregate ? tsz : 2 * tsz + (size_t)(kBThreads / kWarp) * HG * kBD * sizeof(float);
    cudaError_t e = aggregate
        ? cudaFuncSetAttribute(bf16_aggregate_kernel<HG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
        : cudaFuncSetAttribute(bf16_logits_kernel<HG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        set_error("bf16 forward smem attribute: %s", cudaGetErrorString(e));
        return STB200_ERR_CUDA;
    }
    const int groups = p.h / HG;
    const int chunks = (p.N + 63) / 64;
    const int ctas_per_sm = (int)max((size_t)1, min((size_t)4, (size_t)(220 * 1024) / max(smem, (size_t)1)));
    dim3 grid(max(1, min(chunks, (kNumSMs * ctas_per_sm + groups - 1) / groups)), groups);
    const double C = (double)p.h * kBD;
    if (aggregate) {
        KernelScope ks("bf16[aggregate_fwd]", 4.0 * M * p.h + 8.0 * M + 2.0 * p.N * C + 4.0 * p.N * C + 4.0 * p.N, s);
        bf16_aggregate_kernel<HG><<<grid, kBThreads, smem, s>>>(p);
    } else {
        KernelScope ks("bf16[logits_fwd]", 4.0 * M * p.h + 8.0 * M + 4.0 * p.N * C + 4.0 * p.N, s);
        bf16_logits_kernel<HG><<<grid, kBThreads, smem, s>>>(p);
    }
    return check_launch("bf16 forward");
}

static int dispatch_bf16(const BfParams &p, int M, bool aggregate, cudaStream_t s) {
    int cap = 4;
    while (cap > 1 && (size_t)(aggregate ? 1 : 2) * 3 * p.L * cap * kBD * 2 > 100 * 1024) --cap;
    switch (largest_head_group(p.h, cap)) {
        case 4: return launch_bf16<4>(p, M, aggregate, s);
        case 3: return launch_bf16<3>(p, M, aggregate, s);
        case 2: return launch_bf16<2>(p, M, aggregate, s);
        default: return launch_bf16<1>(p, M, aggregate, s);
    }
}

}  // namespace stb200

using namespace stb200;

extern "C" {

int stb200_window_logits_forward_bf16(const stb200_index *ix, int h, int hdim, int L, const void *q_bf16, const void *k_bf16,
                                      const float *table_q, const float *table_k, float *logits, void *stream) {
    STB200_REQUIRE(ix && ix->N >= 0 && ix->M >= 0 && h > 0, STB200_ERR_ARG, "bad sizes");
    STB200_REQUIRE(hdim == kBD, STB200_ERR_HEAD_DIM, "bf16 forward supports head dim 16 only (got %d)", hdim);
    if (ix->M == 0 || ix->N == 0) return STB200_OK;
    STB200_REQUIRE(L > 0 && L <= 1024 && ix->index0_offsets && ix->index1 && ix->rel_packed && q_bf16 && k_bf16 && table_q && table_k &&
                       logits, STB200_ERR_ARG, "null pointer (rel_packed is required) or bad L");
    BfParams p{};
    p.N = ix->N; p.h = h; p.L = L; p.q = (const __nv_bfloat16 *)q_bf16; p.k = (const __nv_bfloat16 *)k_bf16;
    p.tq = table_q; p.tk = table_k; p.offsets = ix->index0_offsets; p.index1 = ix->index1; p.row_order = ix->row_order;
    p.packed = ix->rel_packed; p.out = logits;
    return dispatch_bf16(p, ix->M, false, (cudaStream_t)stream);
}

int stb200_window_aggregate_forward_bf16(const stb200_index *ix, int h, int hdim, int L, const float *attn, const void *v_bf16,
                                         const float *table_v, float *output, void *stream) {
    STB200_REQUIRE(ix && ix->N >= 0 && ix->M >= 0 && h > 0, STB200_ERR_ARG, "bad sizes");
    STB200_REQUIRE(hdim == kBD, STB200_ERR_HEAD_DIM, "bf16 forward supports head dim 16 only (got %d)", hdim);
    if (ix->N == 0) return STB200_OK;
    STB200_REQUIRE(L > 0 && L <= 1024 && ix->index0_offsets && output && (ix->M == 0 || (ix->index1 && ix->rel_packed && attn && v_bf16 && table_v)),
                   STB200_ERR_ARG, "null pointer (rel_packed is required) or bad L");
    BfParams p{};
    p.N = ix->N; p.h = h; p.L = L; p.v = (const __nv_bfloat16 *)v_bf16; p.tv = table_v; p.offsets = ix->index0_offsets;
    p.index1 = ix->index1; p.row_order = ix->row_order; p.packed = ix->rel_packed; p.attn = attn; p.out = output;
    return dispatch_bf16(p, ix->M, true, (cudaStream_t)stream);
}

}  // extern "C"
