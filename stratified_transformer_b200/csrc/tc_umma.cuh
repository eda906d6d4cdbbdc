// tcgen05 (5th-generation tensor core) primitives for sm_100a, hand-written PTX: TMEM allocation, shared-memory matrix
// descriptors for the un-swizzled canonical layouts, single-thread MMA issue (kind::tf32), commit to an mbarrier, TMEM loads.
//
// Operand layout used throughout ("chunked" = the UMMA canonical layout without swizzle): a matrix X[rows][cols] of 32-bit
// words is stored as 16-byte chunks of 4 consecutive columns; the chunks of 8 consecutive rows are adjacent (one 128-byte
// "core matrix"), core matrices of neighbouring column-quads are `cq_stride` bytes apart, those of neighbouring row-octets
// `ro_stride` bytes apart:
//        byte(r, c) = (r / 8) * ro_stride + (c / 4) * cq_stride + (r % 8) * 16 + (c % 4) * 4
// The SAME bytes serve two descriptor interpretations:
//   * K-major  operand with MN = rows, K = cols : leading byte offset (K direction)  = cq_stride, stride byte offset (MN) = ro_stride
//   * MN-major operand with MN = cols, K = rows : leading byte offset (K direction)  = ro_stride, stride byte offset (MN) = cq_stride
// (cute/atom/mma_traits_sm100.hpp, make_umma_desc: K-major INTERLEAVE ((8,n),2):((1,SBO),LBO), MN-major INTERLEAVE
// ((1,n),(8,k)):((X,SBO),(1,LBO)), in 16-byte units.)  That is what lets one staged copy of q / k / g / v rows, of a table
// and of a histogram feed every GEMM of the fused kernels without re-layout.
//
// fp32 accuracy: kind::tf32 reads the upper 19 bits of each 32-bit operand word.  Every operand is staged twice, hi =
// x with the low 13 mantissa bits cleared and lo = x - hi (exact), and a product is issued as hi*hi + lo*hi + hi*lo
// (3xTF32): the dropped lo*lo term is 2^-22 relative.
//
// Compiled for the host (FW_HOST_EMU) the same entry points operate on plain arrays that stand in for TMEM and shared
// memory, so tests/emu checks the layouts, the descriptor arithmetic and the issue sequence on the CPU; what only the GPU
// can confirm (that the hardware reads the descriptors the way this header writes them) is covered by stb200_tc_selftest.
#pragma once
#include <stdint.h>

namespace stb200 {
namespace tc {

struct OperandView {     // one chunked matrix in shared memory, as an MMA operand
    uint32_t addr;       // shared-memory byte address (device: 32-bit shared window; host emulation: offset into the smem array)
    uint32_t k_stride;   // bytes between core matrices along K
    uint32_t mn_stride;  // bytes between core matrices along M / N
    int mn_major;        // 0: K-major (K = the 4-word chunks' direction), 1: MN-major
};

// byte offset of word (r, c) of a chunked matrix
#if defined(__CUDACC__) && !defined(FW_HOST_EMU)
__host__ __device__
#endif
inline uint32_t chunked_off(int r, int c, uint32_t ro_stride, uint32_t cq_stride) {
    return (uint32_t)(r >> 3) * ro_stride + (uint32_t)(c >> 2) * cq_stride + (uint32_t)(r & 7) * 16u + (uint32_t)(c & 3) * 4u;
}

inline
#if defined(__CUDACC__) && !defined(FW_HOST_EMU)
__host__ __device__
#endif
OperandView k_major_view(uint32_t addr, uint32_t ro_stride, uint32_t cq_stride) { return OperandView{addr, cq_stride, ro_stride, 0}; }
inline
#if defined(__CUDACC__) && !defined(FW_HOST_EMU)
__host__ __device__
#endif
OperandView mn_major_view(uint32_t addr, uint32_t ro_stride, uint32_t cq_stride) { return OperandView{addr, ro_stride, cq_stride, 1}; }

// instruction descriptor, kind::tf32, fp32 accumulate (cute/arch/mma_sm100_desc.hpp, InstrDescriptor)
inline
#if defined(__CUDACC__) && !defined(FW_HOST_EMU)
__host__ __device__
#endif
uint32_t make_idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) /* D = f32 */ | (2u << 7) /* A = tf32 */ | (2u << 10) /* B = tf32 */ | ((uint32_t)a_mn_major << 15) |
           ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

#if defined(__CUDACC__) && !defined(FW_HOST_EMU)
// ------------------------------------------------------------------------------------------------ device
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_smem_desc(const OperandView &v, uint32_t k_byte_advance) {
    // start address, leading (K) byte offset, stride (MN) byte offset in 16-byte units; version 1 (sm_100); no swizzle
    const uint32_t a = v.addr + k_byte_advance;
    return (uint64_t)((a >> 4) & 0x3fffu) | ((uint64_t)((v.k_stride >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((v.mn_stride >> 4) & 0x3fffu) << 32) | (1ull << 46);
}

// TMEM: allocate `cols` (power of two >= 32) columns; the base address lands in *slot (shared memory).  One full warp.
__device__ __forceinline__ void tmem_alloc(uint32_t *slot, int cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t base, int cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy writes to shared memory (st.shared) -> visible to the async proxy (the MMA unit reads operands through it)
__device__ __forceinline__ void fence_smem_to_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// bounded wait: returns false when the phase did not complete (a descriptor / issue bug), so a broken build fails instead of hanging the box
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    for (int spin = 0; spin < (1 << 22); ++spin) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
        if (done) return true;
    }
    return false;
}

// D[tmem] (+)= A * B, one instruction (K = 8 words), issued by ONE thread
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    const uint32_t acc = accumulate ? 1u : 0u;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
// all MMAs issued so far by this thread arrive on `bar` when they complete
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 3xTF32 GEMM  D[M x N] (+)= A[M x K] * B[N x K]^T  from chunked hi / lo operands; K multiple of 8.  One thread.
// k advances by one MMA step = 8 words: 2 column-quads for a K-major operand, one row-octet for an MN-major one.
__device__ __forceinline__ void gemm_3xtf32(uint32_t d_tmem, const OperandView &a_hi, const OperandView &a_lo, const OperandView &b_hi,
                                            const OperandView &b_lo, int M, int N, int K, bool accumulate) {
    const uint32_t idesc = make_idesc_tf32(M, N, a_hi.mn_major, b_hi.mn_major);
    const uint32_t a_step = a_hi.mn_major ? a_hi.k_stride : 2 * a_hi.k_stride;
    const uint32_t b_step = b_hi.mn_major ? b_hi.k_stride : 2 * b_hi.k_stride;
    for (int ks = 0; ks < K / 8; ++ks) {
        const uint64_t ah = make_smem_desc(a_hi, ks * a_step), al = make_smem_desc(a_lo, ks * a_step);
        const uint64_t bh = make_smem_desc(b_hi, ks * b_step), bl = make_smem_desc(b_lo, ks * b_step);
        mma_tf32(d_tmem, al, bh, idesc, accumulate || ks > 0);   // small terms first
        mma_tf32(d_tmem, ah, bl, idesc, true);
        mma_tf32(d_tmem, ah, bh, idesc, true);
    }
}

// TMEM -> registers: 16 consecutive columns of this thread's lane (lane = 32 * (warp % 4) + laneid)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
#else
// ------------------------------------------------------------------------------------------------ host emulation
inline float tf32_hi(float x) {
    uint32_t u;
    __builtin_memcpy(&u, &x, 4);
    u &= 0xffffe000u;
    float y;
    __builtin_memcpy(&y, &u, 4);
    return y;
}
#endif

}  // namespace tc
}  // namespace stb200
