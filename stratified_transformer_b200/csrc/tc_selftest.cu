// Self-test of the tcgen05 primitives (tc_umma.cuh): one CTA computes a small 3xTF32 GEMM through TMEM and dumps the whole
// accumulator.  tests/test_gpu_tc.py compares it with an fp64 product for every operand interpretation the fused kernels
// rely on (K-major and MN-major views of the same chunked bytes, M = 128 and M = 64).
#include "common.cuh"
#include "tc_umma.cuh"

namespace stb200 {

// mode 0: A [M x K] row major, B [N x K] row major -> D = A B^T      (both operands K-major)
// mode 1: A [K x M] row major, B [K x N] row major -> D = A^T B      (both operands MN-major)
// mode 2: A [M x K] row major (K-major), B [K x N] row major (MN-major) -> D = A B
// out [128 x N]: TMEM lanes 0..127 of the accumulator, whatever M is.
__global__ void __launch_bounds__(128, 1) tc_selftest_kernel(int mode, int variant, int M, int N, int K, const float *__restrict__ A,
                                                            const float *__restrict__ B, float *__restrict__ out, int *status) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
    const bool a_mn = mode == 1, b_mn = mode == 1 || mode == 2;
    // chunked matrices: stored [rows][cols] with rows = leading index of the global array
    const int a_rows = a_mn ? K : M, a_cols = a_mn ? M : K;
    const int b_rows = b_mn ? K : N, b_cols = b_mn ? N : K;
    const uint32_t a_cq = 128, a_ro = (uint32_t)(a_cols / 4) * 128;
    const uint32_t b_cq = 128, b_ro = (uint32_t)(b_cols / 4) * 128;
    const uint32_t a_bytes = (uint32_t)(a_rows / 8) * a_ro, b_bytes = (uint32_t)(b_rows / 8) * b_ro;
    // variant bits (development probes): 1 = hi*hi term only, 2 = B buffers first, 4 = 1 KB gaps between the buffers
    const uint32_t gap = (variant & 4) ? 1024u : 0u;
    unsigned char *a_hi, *a_lo, *b_hi, *b_lo;
    if (variant & 2) { b_hi = smem; b_lo = b_hi + b_bytes + gap; a_hi = b_lo + b_bytes + gap; a_lo = a_hi + a_bytes + gap; }
    else { a_hi = smem; a_lo = a_hi + a_bytes + gap; b_hi = a_lo + a_bytes + gap; b_lo = b_hi + b_bytes + gap; }
    int tcols = 32;
    while (tcols < N) tcols *= 2;
    if (warp == 0) tc::tmem_alloc(&tmem_slot, tcols);
    if (tid == 0) tc::mbar_init(&bar, 1);
    for (int e = tid; e < a_rows * a_cols; e += blockDim.x) {
        const int r = e / a_cols, c = e % a_cols;
        const float x = A[e], h = tc::tf32_hi(x);
        const uint32_t o = tc::chunked_off(r, c, a_ro, a_cq);
        *reinterpret_cast<float *>(a_hi + o) = h;
        *reinterpret_cast<float *>(a_lo + o) = x - h;
    }
    for (int e = tid; e < b_rows * b_cols; e += blockDim.x) {
        const int r = e / b_cols, c = e % b_cols;
        const float x = B[e], h = tc::tf32_hi(x);
        const uint32_t o = tc::chunked_off(r, c, b_ro, b_cq);
        *reinterpret_cast<float *>(b_hi + o) = h;
        *reinterpret_cast<float *>(b_lo + o) = x - h;
    }
    tc::fence_smem_to_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tbase = tmem_slot;
    if (tid == 0) {
        auto view = [&](unsigned char *p, bool mn, uint32_t ro, uint32_t cq) {
            return mn ? tc::mn_major_view(tc::smem_u32(p), ro, cq) : tc::k_major_view(tc::smem_u32(p), ro, cq);
        };
        if (variant & 1) {
            const tc::OperandView va = view(a_hi, a_mn, a_ro, a_cq), vb = view(b_hi, b_mn, b_ro, b_cq);
            const uint32_t idesc = tc::make_idesc_tf32(M, N, va.mn_major, vb.mn_major);
            for (int ks = 0; ks < K / 8; ++ks)
                tc::mma_tf32(tbase, tc::make_smem_desc(va, ks * (va.mn_major ? va.k_stride : 2 * va.k_stride)),
                             tc::make_smem_desc(vb, ks * (vb.mn_major ? vb.k_stride : 2 * vb.k_stride)), idesc, ks > 0);
        } else {
            tc::gemm_3xtf32(tbase, view(a_hi, a_mn, a_ro, a_cq), view(a_lo, a_mn, a_ro, a_cq), view(b_hi, b_mn, b_ro, b_cq),
                            view(b_lo, b_mn, b_ro, b_cq), M, N, K, false);
        }
        tc::mma_commit(&bar);
    }
    const bool ok = tc::mbar_wait(&bar, 0);
    tc::fence_after_sync();
    if (!ok && tid == 0) atomicExch(status, 1);
    if (ok) {
        for (int c0 = 0; c0 < N; c0 += 8) {
            float v[8];
            tc::tmem_ld8(tbase + ((uint32_t)(32 * warp) << 16) + (uint32_t)c0, v);
            for (int i = 0; i < 8; ++i) out[(size_t)(32 * warp + lane) * N + c0 + i] = v[i];
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tbase, tcols);
}

}  // namespace stb200

using namespace stb200;

extern "C" int stb200_tc_selftest(int mode_variant, int M, int N, int K, const float *A, const float *B, float *out, int *status, void *stream) {
    const int mode = mode_variant & 0xff, variant = mode_variant >> 8;
    STB200_REQUIRE((M == 64 || M == 128) && N % 8 == 0 && N >= 8 && N <= 256 && K % 8 == 0 && K > 0 && mode >= 0 && mode <= 2,
                   STB200_ERR_ARG, "tc selftest: unsupported shape M=%d N=%d K=%d mode=%d", M, N, K, mode);
    STB200_REQUIRE(A && B && out && status, STB200_ERR_ARG, "null pointer");
    const size_t smem = 2 * ((size_t)M * K + (size_t)N * K) * sizeof(float) + 4096;
    STB200_REQUIRE(smem <= 200 * 1024, STB200_ERR_ARG, "tc selftest: operands too large");
    cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    KernelScope ks("tc_selftest", 0.0, (cudaStream_t)stream);
    tc_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(mode, variant, M, N, K, A, B, out, status);
    return check_launch("tc_selftest");
}
