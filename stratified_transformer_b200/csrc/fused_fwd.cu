// Fused window-attention forward: logits + rel-pos bias + softmax + aggregation for one window at a
// time, with no M-sized intermediate except the attention probabilities the backward pass needs.
//
// Math = the three reference operators chained as in WindowAttention.forward
// (/root/reference/model/stratified_transformer.py:183-208; kernels: attention_v2/attention_cuda_kernel_v2.cu:39-45,
// rpe_v2/relative_pos_encoding_cuda_kernel_v2.cu:272-281,422-437).  What changes is where the rel-pos tables meet the
// data.  In a window all queries share one key list, so with
//     QT[q][(a,l)] = <q_q, T_q[l,:,a]>      KT[k][(a,l)] = <k_k, T_k[l,:,a]>
// the bias of pair (q,k) is six scalar look-ups  QT[q][r_a] + KT[k][r_a]  instead of six 64-byte table rows, and
//     sum_k p[q,k] * Ev(q,k) = Ph[q] . T_v,   Ph[q][(a,l)] = sum_{k: r_a(q,k)=l} p[q,k]
// so the value-table term is a per-query histogram times the table.  QT, KT, q.k^T, p.V and Ph.T_v are small dense
// GEMMs per window: they run on the tensor cores (mma.sync m16n8k8 TF32, 3-term split = fp32-level accuracy) out of
// shared memory, where K, V, KT live for the whole window.
//
// A window is eligible when every query of it has exactly the same key list (always true unless the reference's two
// window-id roundings disagree for one of its points) and the list fits the shared-memory tile; all other
// windows are left to the per-pair kernels (the caller runs them on the complementary row list).
#include "common.cuh"

namespace stb200 {

constexpr int kFThreads = 256;
constexpr int kFWarps = kFThreads / kWarp;
constexpr int kQT = 32;      // queries per tile
constexpr int kNKMax = 96;   // keys per window handled by the fused kernel
constexpr int kD = 16;       // head dim (all shipped configs)
constexpr int kPA = 20;      // pitch of Q / K tiles   (A operand, and B operand in [n][k] form): 20 = 4 mod 32
constexpr int kPV = 24;      // pitch of the V tile    (B operand in [k][n] form): 24 = 8 * 3 mod 32
constexpr int kSP = kNKMax + 4;   // pitch of the score tile (A operand): 100 = 4 mod 32

__device__ __forceinline__ void split_tf32_f(float x, unsigned &hi, unsigned &lo) {
    hi = __float_as_uint(x);
    lo = __float_as_uint(x - __uint_as_float(hi & 0xffffe000u));
}
__device__ __forceinline__ void mma_tf32_f(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// c += A[m0:m0+16, 0:K] * B[0:K, n0:n0+8].  A is [m][lda] (k contiguous); B is [k][ldb] when B_KN, else [n][ldb].
template <bool B_KN>
__device__ __forceinline__ void mma_tile(float (&c)[4], const float *A, int lda, const float *B, int ldb, int m0, int n0,
                                         int K, int gid, int tig) {
    const float *a_lo = A + (m0 + gid) * lda + tig, *a_hi = a_lo + 8 * lda;
    for (int k0 = 0; k0 < K; k0 += 8) {
        unsigned ah[4], al[4], bh[2], bl[2];
        split_tf32_f(a_lo[k0], ah[0], al[0]);
        split_tf32_f(a_hi[k0], ah[1], al[1]);
        split_tf32_f(a_lo[k0 + 4], ah[2], al[2]);
        split_tf32_f(a_hi[k0 + 4], ah[3], al[3]);
        const float b0 = B_KN ? B[(k0 + tig) * ldb + n0 + gid] : B[(n0 + gid) * ldb + k0 + tig];
        const float b1 = B_KN ? B[(k0 + tig + 4) * ldb + n0 + gid] : B[(n0 + gid) * ldb + k0 + tig + 4];
        split_tf32_f(b0, bh[0], bl[0]);
        split_tf32_f(b1, bh[1], bl[1]);
        mma_tf32_f(c, al, bh);
        mma_tf32_f(c, ah, bl);
        mma_tf32_f(c, ah, bh);
    }
}

// store a C fragment (rows gid / gid+8, cols 2*tig, 2*tig+1 of the tile) to shared memory
__device__ __forceinline__ void store_tile(float *C, int ldc, int m0, int n0, const float (&c)[4], int gid, int tig) {
    float *r0 = C + (m0 + gid) * ldc + n0 + 2 * tig, *r1 = r0 + 8 * ldc;
    r0[0] = c[0]; r0[1] = c[1];
    r1[0] = c[2]; r1[1] = c[3];
}

struct FusedParams {
    int N, h, L, n_win;
    const int *offsets, *index1, *row_order, *win_offsets;
    const unsigned char *win_flags;
    const unsigned *packed;     // [M] bins 10 bits each, query-segment order
    const float *q, *k, *v, *tq, *tk, *tv;
    float *out, *p;
};

// flags[w] = 1 iff all queries of window w have the same key list and it has at most kNKMax keys
__global__ void classify_windows_kernel(int n_win, const int *__restrict__ win_offsets, const int *__restrict__ row_order,
                                        const int *__restrict__ offsets, const int *__restrict__ index1,
                                        unsigned char *__restrict__ flags) {
    const int lane = threadIdx.x % kWarp;
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) / kWarp, nw = gridDim.x * blockDim.x / kWarp;
    for (int w = wid; w < n_win; w += nw) {
        const int w0 = win_offsets[w], nq = win_offsets[w + 1] - w0;
        const int q0 = row_order[w0];
        const int s0 = offsets[q0], nk = offsets[q0 + 1] - s0;
        bool ok = nk <= kNKMax && nk > 0;
        for (int i = 1; i < nq && ok; ++i) {
            const int qi = row_order[w0 + i];
            const int si = offsets[qi];
            bool same = offsets[qi + 1] - si == nk;
            if (same)
                for (int j = lane; j < nk; j += kWarp) same &= index1[si + j] == index1[s0 + j];
            ok = __all_sync(0xffffffffu, same);
        }
        if (lane == 0) flags[w] = ok ? 1 : 0;
    }
}

// rows of the windows the fused kernel does not take (for the per-pair fallback): out_rows[0..count)
__global__ void fallback_rows_kernel(int n_win, const int *__restrict__ win_offsets, const int *__restrict__ row_order,
                                     const unsigned char *__restrict__ flags, int *__restrict__ out_rows, int *counter) {
    const int lane = threadIdx.x % kWarp;
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) / kWarp, nw = gridDim.x * blockDim.x / kWarp;
    for (int w = wid; w < n_win; w += nw) {
        if (flags[w]) continue;
        const int w0 = win_offsets[w], nq = win_offsets[w + 1] - w0;
        int base = 0;
        if (lane == 0) base = atomicAdd(counter, nq);
        base = __shfl_sync(0xffffffffu, base, 0);
        for (int i = lane; i < nq; i += kWarp) out_rows[base + i] = row_order[w0 + i];
    }
}

__global__ void __launch_bounds__(kFThreads, 1) fused_forward_kernel(const FusedParams p, int Rpad) {
    extern __shared__ float4 smem4[];
    float *sm = reinterpret_cast<float *>(smem4);
    const int L = p.L, h = p.h, C = h * kD, R = 3 * L;
    const int head = blockIdx.y;
    const int PQK = Rpad + 8;    // pitch of Tq^T / Tk^T ([c][row], B operand in [k][n] form): 8 or 24 mod 32
    const int PT = Rpad + 4;     // pitch of Tv^T ([c][row], B operand in [n][k] form) and of QT / KT / Ph rows: 4 mod 32
    float *TqT = sm;                              // [16][PQK]
    float *TkT = TqT + kD * PQK;                  // [16][PQK]
    float *TvT = TkT + kD * PQK;                  // [16][PT]
    float *KTs = TvT + kD * PT;                   // [kNKMax][PT]
    float *QTs = KTs + kNKMax * PT;               // [kQT][PT]      (re-used as the histogram Ph after the logits)
    float *Ss = QTs + kQT * PT;                   // [kQT][kSP]
    float *Qs = Ss + kQT * kSP;                   // [kQT][kPA]
    float *Ks = Qs + kQT * kPA;                   // [kNKMax][kPA]
    float *Vs = Ks + kNKMax * kPA;                // [kNKMax][kPV]
    float *Op = Vs + kNKMax * kPV;                // [4 tiles][16][8] partial outputs of the second k-half
    unsigned *Rs = reinterpret_cast<unsigned *>(Op + 4 * 128);   // [kQT][kNKMax] packed bins
    int *kid = reinterpret_cast<int *>(Rs + kQT * kNKMax);       // [kNKMax] key ids
    int *qid = kid + kNKMax;                      // [kQT] query ids
    int *qoff = qid + kQT;                        // [kQT] first pair of each query

    const int tid = threadIdx.x, lane = tid % kWarp, warp = tid / kWarp;
    const int gid = lane >> 2, tig = lane & 3;

    // ---- tables of this head, transposed to [c][(axis, l)], zero beyond the last row
    for (int i = tid; i < kD * (PQK + PQK + PT); i += kFThreads) sm[i] = 0.f;
    __syncthreads();
    for (int i = tid; i < L * kD * 3; i += kFThreads) {
        const int l = i / (kD * 3), e = i - l * (kD * 3);
        const int c = e / 3, a = e - c * 3;
        const size_t g = (size_t)(l * h + head) * (kD * 3) + e;
        TqT[c * PQK + a * L + l] = __ldg(p.tq + g);
        TkT[c * PQK + a * L + l] = __ldg(p.tk + g);
        TvT[c * PT + a * L + l] = __ldg(p.tv + g);
    }
    __syncthreads();

    for (int w = blockIdx.x; w < p.n_win; w += gridDim.x) {
        if (!p.win_flags[w]) continue;
        const int w0 = p.win_offsets[w], nq = p.win_offsets[w + 1] - w0;
        const int q0 = p.row_order[w0];
        const int s0 = p.offsets[q0], nk = p.offsets[q0 + 1] - s0;
        const int nkp = (nk + 15) & ~15;   // key rows used by the GEMMs (multiple of 16)

        // ---- keys of the window: ids, K and V rows of this head (zero padded)
        __syncthreads();   // previous window done with Ks / Vs / KTs
        for (int j = tid; j < nkp; j += kFThreads) kid[j] = j < nk ? __ldg(p.index1 + s0 + j) : -1;
        __syncthreads();
        for (int i = tid; i < nkp * 4; i += kFThreads) {
            const int j = i >> 2, c4 = i & 3;
            float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
            if (kid[j] >= 0) {
                const size_t g = (size_t)kid[j] * C + head * kD + 4 * c4;
                kv = ld_row4(p.k + g);
                vv = ld_row4(p.v + g);
            }
            *reinterpret_cast<float4 *>(Ks + j * kPA + 4 * c4) = kv;
            *reinterpret_cast<float4 *>(Vs + j * kPV + 4 * c4) = vv;
        }
        __syncthreads();
        // ---- KT = K . Tk^T : [nkp x 16] x [16 x Rpad]
        {
            const int nts = Rpad / 8, tiles = (nkp / 16) * nts;
            for (int t = warp; t < tiles; t += kFWarps) {
                const int mt = t / nts, nt = t - mt * nts;
                float c[4] = {0.f, 0.f, 0.f, 0.f};
                mma_tile<true>(c, Ks, kPA, TkT, PQK, mt * 16, nt * 8, kD, gid, tig);
                store_tile(KTs, PT, mt * 16, nt * 8, c, gid, tig);
            }
        }

        for (int qt = 0; qt < nq; qt += kQT) {
            const int nqt = min(kQT, nq - qt);
            __syncthreads();   // previous query tile done with Qs / Ss / QTs / Rs; KTs complete
            if (tid < kQT) {
                const int n = tid < nqt ? __ldg(p.row_order + w0 + qt + tid) : -1;
                qid[tid] = n;
                qoff[tid] = n >= 0 ? __ldg(p.offsets + n) : 0;
            }
            __syncthreads();
            for (int i = tid; i < kQT * 4; i += kFThreads) {
                const int r = i >> 2, c4 = i & 3;
                float4 qv = make_float4(0.f, 0.f, 0.f, 0.f);
                if (qid[r] >= 0) qv = ld_row4(p.q + (size_t)qid[r] * C + head * kD + 4 * c4);
                *reinterpret_cast<float4 *>(Qs + r * kPA + 4 * c4) = qv;
            }
            for (int i = tid; i < nqt * nk; i += kFThreads) {
                const int r = i / nk, j = i - r * nk;
                Rs[r * kNKMax + j] = __ldg(p.packed + qoff[r] + j);
            }
            __syncthreads();
            // ---- QT = Q . Tq^T  and  S = Q . K^T
            {
                const int nts = Rpad / 8, t_qt = 2 * nts, t_s = 2 * (nkp / 8);
                for (int t = warp; t < t_qt + t_s; t += kFWarps) {
                    float c[4] = {0.f, 0.f, 0.f, 0.f};
                    if (t < t_qt) {
                        const int mt = t / nts, nt = t - mt * nts;
                        mma_tile<true>(c, Qs, kPA, TqT, PQK, mt * 16, nt * 8, kD, gid, tig);
                        store_tile(QTs, PT, mt * 16, nt * 8, c, gid, tig);
                    } else {
                        const int u = t - t_qt, mt = u / (nkp / 8), nt = u - mt * (nkp / 8);
                        mma_tile<false>(c, Qs, kPA, Ks, kPA, mt * 16, nt * 8, kD, gid, tig);
                        store_tile(Ss, kSP, mt * 16, nt * 8, c, gid, tig);
                    }
                }
            }
            __syncthreads();
            // ---- bias + softmax, one warp per query row; probabilities go to Ss (zero padded) and to global p
            for (int r = warp; r < nqt; r += kFWarps) {
                float s[kNKMax / kWarp];
                float mx = -INFINITY;
#pragma unroll
                for (int u = 0; u < kNKMax / kWarp; ++u) {
                    const int j = lane + u * kWarp;
                    s[u] = -INFINITY;
                    if (j < nk) {
                        const unsigned pk = Rs[r * kNKMax + j];
                        const int r0 = pk & 0x3ff, r1 = L + ((pk >> 10) & 0x3ff), r2 = 2 * L + (pk >> 20);
                        const float *qt_row = QTs + r * PT, *kt_row = KTs + j * PT;
                        s[u] = Ss[r * kSP + j] + ((qt_row[r0] + qt_row[r1]) + qt_row[r2]) + ((kt_row[r0] + kt_row[r1]) + kt_row[r2]);
                    }
                    mx = fmaxf(mx, s[u]);
                }
                mx = warp_max(mx);
                float sum = 0.f;
#pragma unroll
                for (int u = 0; u < kNKMax / kWarp; ++u) {
                    s[u] = (lane + u * kWarp < nk) ? expf(s[u] - mx) : 0.f;
                    sum += s[u];
                }
                sum = group_sum<kWarp>(sum);
                float *gp = p.p + (size_t)qoff[r] * h + head;
#pragma unroll
                for (int u = 0; u < kNKMax / kWarp; ++u) {
                    const int j = lane + u * kWarp;
                    const float pv = s[u] / sum;
                    if (j < nkp) Ss[r * kSP + j] = j < nk ? pv : 0.f;
                    if (j < nk) gp[(size_t)j * h] = pv;
                }
            }
            __syncthreads();
            // ---- Ph[q][(a,l)] = sum_k [r_a(q,k) = l] p[q,k]   (QT no longer needed: same buffer)
            for (int i = tid; i < kQT * PT; i += kFThreads) QTs[i] = 0.f;
            __syncthreads();
            if (tid < 3 * kQT) {
                const int r = tid % kQT, a = tid / kQT;
                if (r < nqt) {
                    float *ph = QTs + r * PT + a * L;
                    const int sh = 10 * a;
                    for (int j = 0; j < nk; ++j) ph[(Rs[r * kNKMax + j] >> sh) & 0x3ffu] += Ss[r * kSP + j];
                }
            }
            __syncthreads();
            // ---- out = P . V + Ph . Tv : 4 output tiles (2 x 2), the k range split over two warp groups
            {
                const int t = warp & 3, part = warp >> 2;
                const int mt = t >> 1, nt = t & 1;
                float c[4] = {0.f, 0.f, 0.f, 0.f};
                const int half = (Rpad / 16) * 8;   // table rows of the first half (multiple of 8)
                if (part == 0) {
                    mma_tile<true>(c, Ss, kSP, Vs, kPV, mt * 16, nt * 8, nkp, gid, tig);
                    mma_tile<false>(c, QTs, PT, TvT, PT, mt * 16, nt * 8, half, gid, tig);
                } else {
                    mma_tile<false>(c, QTs + half, PT, TvT + half, PT, mt * 16, nt * 8, Rpad - half, gid, tig);
                    float *o = Op + t * 128 + gid * 8 + 2 * tig;
                    o[0] = c[0]; o[1] = c[1]; o[64] = c[2]; o[65] = c[3];
                }
                __syncthreads();
                if (part == 0) {
                    const float *o = Op + t * 128 + gid * 8 + 2 * tig;
                    c[0] += o[0]; c[1] += o[1]; c[2] += o[64]; c[3] += o[65];
                    const int ra = mt * 16 + gid, rb = ra + 8, col = head * kD + nt * 8 + 2 * tig;
                    if (ra < nqt) *reinterpret_cast<float2 *>(p.out + (size_t)qid[ra] * C + col) = make_float2(c[0], c[1]);
                    if (rb < nqt) *reinterpret_cast<float2 *>(p.out + (size_t)qid[rb] * C + col) = make_float2(c[2], c[3]);
                }
            }
        }
    }
}

static size_t fused_smem_bytes(int Rpad) {
    const int PQK = Rpad + 8, PT = Rpad + 4;
    size_t f = (size_t)kD * (2 * PQK + PT) + (size_t)kNKMax * PT + (size_t)kQT * PT + kQT * kSP + kQT * kPA + kNKMax * kPA +
               kNKMax * kPV + 4 * 128 + (size_t)kQT * kNKMax + kNKMax + 2 * kQT;
    return f * sizeof(float);
}

}  // namespace stb200

using namespace stb200;

extern "C" {

int stb200_fused_max_keys(void) { return kNKMax; }

int stb200_classify_windows(int n_win, const int *win_offsets, const int *row_order, const int *index0_offsets,
                            const int *index1, unsigned char *flags, int *fallback_rows, int *fallback_count, void *stream) {
    STB200_REQUIRE(n_win >= 0, STB200_ERR_ARG, "bad n_win");
    if (n_win == 0) return STB200_OK;
    STB200_REQUIRE(win_offsets && row_order && index0_offsets && index1 && flags && fallback_rows && fallback_count, STB200_ERR_ARG,
                   "null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    const int blocks = max(1, min((n_win + 7) / 8, kNumSMs * 8));
    KernelScope ks("classify_windows[3 launches]", 0.0, s);
    count_launch(2);
    cudaMemsetAsync(fallback_count, 0, sizeof(int), s);
    classify_windows_kernel<<<blocks, 256, 0, s>>>(n_win, win_offsets, row_order, index0_offsets, index1, flags);
    fallback_rows_kernel<<<blocks, 256, 0, s>>>(n_win, win_offsets, row_order, flags, fallback_rows, fallback_count);
    return check_launch("classify_windows");
}

int stb200_window_attention_forward_fused(const stb200_index *ix, int n_win, const int *win_offsets,
                                          const unsigned char *win_flags, int h, int hdim, int L, const float *q,
                                          const float *k, const float *v, const float *table_q, const float *table_k,
                                          const float *table_v, float *output, float *attn, void *stream) {
    STB200_REQUIRE(ix && ix->N >= 0 && ix->M >= 0 && n_win >= 0 && h > 0, STB200_ERR_ARG, "bad sizes");
    STB200_REQUIRE(hdim == kD, STB200_ERR_HEAD_DIM, "fused forward supports head dim 16 only (got %d)", hdim);
    STB200_REQUIRE(L > 0 && 3 * L <= 256, STB200_ERR_ARG, "fused forward: table length %d > 85", L);
    if (n_win == 0 || ix->M == 0) return STB200_OK;
    STB200_REQUIRE(ix->index0_offsets && ix->index1 && ix->rel_packed && ix->row_order && win_offsets && win_flags && q && k &&
                       v && table_q && table_k && table_v && output && attn, STB200_ERR_ARG, "null pointer (rel_packed and row_order are required)");
    const int Rpad = (3 * L + 15) / 16 * 16;
    const size_t smem = fused_smem_bytes(Rpad);
    STB200_REQUIRE(smem <= 227 * 1024, STB200_ERR_ARG, "fused forward needs %zu B of shared memory", smem);
    cudaError_t e = cudaFuncSetAttribute(fused_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        set_error("fused forward smem attribute: %s", cudaGetErrorString(e));
        return STB200_ERR_CUDA;
    }
    FusedParams p{};
    p.N = ix->N; p.h = h; p.L = L; p.n_win = n_win;
    p.offsets = ix->index0_offsets; p.index1 = ix->index1; p.row_order = ix->row_order; p.win_offsets = win_offsets;
    p.win_flags = win_flags; p.packed = ix->rel_packed;
    p.q = q; p.k = k; p.v = v; p.tq = table_q; p.tk = table_k; p.tv = table_v; p.out = output; p.p = attn;
    {
        // reads q, k, v rows + packed bins + key ids + offsets, writes out and the probabilities
        const double bytes = 4.0 * ((double)ix->N * h * kD * 4 + (double)ix->M * (2 + h) + 2.0 * ix->N) + 36.0 * L * h * kD;
        KernelScope ks("fused_window_forward", bytes, (cudaStream_t)stream);
        fused_forward_kernel<<<dim3(min(n_win, kNumSMs), h), kFThreads, smem, (cudaStream_t)stream>>>(p, Rpad);
    }
    return check_launch("fused_window_forward");
}

}  // extern "C"
