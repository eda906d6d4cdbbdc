// Window-centric fused attention, tensor-core version: the CTA bodies of fused_phases.cuh with every large GEMM moved
// to the 5th-generation tensor cores (tcgen05.mma kind::tf32, 3xTF32 split for fp32 accuracy, accumulators in TMEM):
//   * per-point table products  QT = Q T_q^T, KT = K T_k^T (, GT = G T_v^T)        [rows x 16] x [16 x R]
//   * the q.k and g.v tiles of every block                                           [rows x 16] x [16 x keys]
//   * table gradients           gT_q += Sq^T Q, gT_k += Sk^T K, gT_v += Ph^T G        [R x rows] x [rows x 16], accumulated in
//     TMEM across all items of the CTA and flushed once
// The element-wise pair phase (6 scalar look-ups per pair), the histogram builds and the small [rows x 16] output GEMMs
// (p V, gs K, gs^T Q, p^T G, histogram x table) stay on the FMA pipe.  Same work items, same pass structure, same
// results (within fp32 rounding) as fused_phases.cuh, which remains the reference implementation for A/B measurements.
//
// All MMA operands are K-major "chunked" matrices (tc_umma.cuh): rows = the M / N index, 16-byte chunks along K.
//   row arrays   R-form [row][16]   : A of the products and of the tiles, B of the tiles        (K = channel)
//   row arrays   T-form [16][row]   : B of the table-gradient GEMMs                             (K = row)
//   tables              [bin][16]   : B of the products                                          (K = channel)
//   histograms          [bin][row]  : A of the table-gradient GEMMs                             (K = row)
// each as a hi part (low 13 mantissa bits cleared) and a lo part (x - hi).  The histograms are split in place: the hi
// parts are multiplied first while every thread keeps its lo values in registers, then the lo parts take the same buffer.
//
// Host emulation (FW_HOST_EMU): the MMA, TMEM and barrier primitives have functional stand-ins below, so tests/emu runs
// this file on the CPU too (layouts, descriptor arithmetic, TMEM row -> lane maps as verified by stb200_tc_selftest).
#pragma once
#include "fused_phases.cuh"
#include "tc_umma.cuh"

#if defined(__CUDACC__) && !defined(FW_HOST_EMU)
#define TC_PHASE_BEGIN {
#define TC_PHASE_END } __syncthreads();
#define TC_PER_THREAD(type, name, n) type name[n]
#define TC_PER_THREAD_USE(type, name)
#else
#define TC_PHASE_BEGIN for (int fw_tid_ = 0; fw_tid_ < NTC; ++fw_tid_) {
#define TC_PHASE_END }
#define TC_PER_THREAD(type, name, n) static thread_local type name##_all[NTC][n]
#define TC_PER_THREAD_USE(type, name) type *name = name##_all[fw_tid_]
#endif

// per-phase clock stamps of one CTA (development: tools/fused_phase_prof.py): cycles since the previous stamp are added to slot k
#if defined(__CUDACC__) && !defined(FW_HOST_EMU)
#define TC_STAMP(k)                                                                  \
    do {                                                                             \
        if (prof_on) {                                                               \
            const long long now_ = clock64();                                        \
            if (threadIdx.x == 0) atomicAdd((unsigned long long *)(P.prof + (k)), (unsigned long long)(now_ - last_stamp)); \
            last_stamp = now_;                                                       \
        }                                                                            \
    } while (0)
#define TC_STAMP_DECL const bool prof_on = P.prof != nullptr && cta == 0 && head == 0; long long last_stamp = prof_on ? clock64() : 0
#else
#define TC_STAMP(k) do { } while (0)
#define TC_STAMP_DECL do { } while (0)
#endif

namespace stb200 {
namespace fwtc {

using namespace fw;

constexpr int NTC = 512;          // threads per CTA
constexpr int KSF = 8;            // K-slices of the forward output GEMM
constexpr uint32_t CQ = 128;      // bytes between the core matrices of neighbouring column quads (all chunked matrices)
constexpr int TMEM_COLS = 512;
constexpr int C_GT = 0;           // TMEM columns [0,48): table-gradient accumulators (bins 0..127), [48,96): bins 128..255
constexpr int C_WORK = 96;        // product slot, then the two tiles

struct TcCtx {
#if defined(__CUDACC__) && !defined(FW_HOST_EMU)
    uint32_t tmem, smem_u32;
    uint64_t *bar;
    uint32_t parity;
#else
    float *tmem;            // [128 lanes][512 columns]
    unsigned char *smem;
#endif
};

FW_HD int round8(int x) { return (x + 7) & ~7; }
FW_HD int round16(int x) { return (x + 15) & ~15; }
FW_HD int m64_lane(int r) { return 32 * (r >> 4) + (r & 15); }   // TMEM lane of row r of an M = 64 accumulator (measured)

// ---- one-thread GEMM issue:  D[M x N] at TMEM column d_col (+)= A[M x K] * B[N x K]^T -------------------------------
// a_* / b_* : byte offsets of chunked operands (already advanced to their first row), ro = row-octet stride in bytes.
// terms: 1 = a_hi*b_hi, 2 = a_hi*b_lo, 4 = a_lo*b_hi  (7 = 3xTF32).
FW_FN void tc_gemm(TcCtx &c, int d_col, int M, int N, int K, uint32_t a_hi, uint32_t a_lo, uint32_t a_ro, uint32_t b_hi, uint32_t b_lo,
                   uint32_t b_ro, int terms, bool accumulate) {
#if defined(__CUDACC__) && !defined(FW_HOST_EMU)
    const uint32_t idesc = tc::make_idesc_tf32(M, N, 0, 0);
    const uint32_t d = c.tmem + (uint32_t)d_col;
    bool acc = accumulate;
    for (int ks = 0; ks < K / 8; ++ks) {
        const uint32_t adv = (uint32_t)ks * 2u * CQ;
        const tc::OperandView ah = tc::k_major_view(c.smem_u32 + a_hi + adv, a_ro, CQ), al = tc::k_major_view(c.smem_u32 + a_lo + adv, a_ro, CQ);
        const tc::OperandView bh = tc::k_major_view(c.smem_u32 + b_hi + adv, b_ro, CQ), bl = tc::k_major_view(c.smem_u32 + b_lo + adv, b_ro, CQ);
        if (terms & 4) { tc::mma_tf32(d, tc::make_smem_desc(al, 0), tc::make_smem_desc(bh, 0), idesc, acc); acc = true; }
        if (terms & 2) { tc::mma_tf32(d, tc::make_smem_desc(ah, 0), tc::make_smem_desc(bl, 0), idesc, acc); acc = true; }
        if (terms & 1) { tc::mma_tf32(d, tc::make_smem_desc(ah, 0), tc::make_smem_desc(bh, 0), idesc, acc); acc = true; }
    }
#else
    auto at = [&](uint32_t base, uint32_t ro, int r, int k) {
        float x = *reinterpret_cast<const float *>(c.smem + base + tc::chunked_off(r, k, ro, CQ));
        return tc::tf32_hi(x);   // the tensor core reads the upper 19 bits
    };
    for (int m = 0; m < M; ++m) {
        const int lane = M == 64 ? m64_lane(m) : m;
        for (int n = 0; n < N; ++n) {
            float s = accumulate ? c.tmem[lane * TMEM_COLS + d_col + n] : 0.f;
            for (int k = 0; k < K; ++k) {
                if (terms & 4) s += at(a_lo, a_ro, m, k) * at(b_hi, b_ro, n, k);
                if (terms & 2) s += at(a_hi, a_ro, m, k) * at(b_lo, b_ro, n, k);
                if (terms & 1) s += at(a_hi, a_ro, m, k) * at(b_hi, b_ro, n, k);
            }
            c.tmem[lane * TMEM_COLS + d_col + n] = s;
        }
    }
#endif
}

// thread 0 before its tc_gemm calls of a phase (orders them after the TMEM reads / barrier that preceded)
FW_FN void tc_begin_issue() {
#if defined(__CUDACC__) && !defined(FW_HOST_EMU)
    tc::fence_after_sync();
#endif
}
// thread 0 after its tc_gemm calls; then every thread of the CTA
FW_FN void tc_commit(TcCtx &c) {
#if defined(__CUDACC__) && !defined(FW_HOST_EMU)
    tc::mma_commit(c.bar);
#else
    (void)c;
#endif
}
FW_FN void tc_wait(TcCtx &c) {
#if defined(__CUDACC__) && !defined(FW_HOST_EMU)
    tc::mbar_wait(c.bar, c.parity);
    c.parity ^= 1u;
    tc::fence_after_sync();
#else
    (void)c;
#endif
}
// writes to operand buffers (generic proxy) -> visible to the tensor core; call in the phase that wrote them
FW_FN void tc_publish_smem() {
#if defined(__CUDACC__) && !defined(FW_HOST_EMU)
    tc::fence_smem_to_async();
#endif
}
// before a barrier that separates TMEM reads (tcgen05.ld) from later MMAs overwriting the same columns
FW_FN void tc_tmem_reads_done() {
#if defined(__CUDACC__) && !defined(FW_HOST_EMU)
    tc::fence_before_sync();
#endif
}
// 16 consecutive columns of this thread's TMEM lane (lane = 32 * (warp % 4) + laneid); whole warps only
FW_FN void tc_load16(TcCtx &c, int tid, int col, float (&v)[16]) {
#if defined(__CUDACC__) && !defined(FW_HOST_EMU)
    tc::tmem_ld16(c.tmem + ((uint32_t)(32 * ((tid >> 5) & 3)) << 16) + (uint32_t)col, v);
#else
    const int lane = 32 * ((tid >> 5) & 3) + (tid & 31);
    for (int i = 0; i < 16; ++i) v[i] = c.tmem[lane * TMEM_COLS + col + i];
#endif
}

// ---- shared-memory plan ------------------------------------------------------------------------------------------------
struct TcLayout {    // byte offsets
    int Rpad, RP, PS, HR, BQm;
    uint32_t ro_row, ro_tab, ro_tq, ro_tk;                 // row-octet strides: R-form rows / tables, T-form (q side / k side)
    uint32_t ro_hq, ro_hk;                                 // histograms (query side [bin][BQ], key side [bin][BK])
    uint32_t tab[3][2];                                    // tables chunked [Rpad][16] hi, lo
    uint32_t tvT;                                          // forward: T_v transposed plain [16][Rpad]
    uint32_t qR[2], gR[2], kR[2], vR[2];                   // R-form rows hi, lo (64 rows each)
    uint32_t qT[2], gT[2], kT[2];                          // T-form rows hi, lo
    uint32_t vplain;                                       // forward: v rows plain [BK][16]
    uint32_t B1, B2, B3;                                   // big buffers: products [rows][RP] / histograms [HR][rows]
    uint32_t P, GS, U;                                     // tiles [BQ][PS] fp32, union rel tile / partial sums
    uint32_t rowinfo, keyid, red, mrow, lrow, drow, slot;
    uint32_t total;
};

FW_HD int hist_rows(int Rpad) { return Rpad <= 128 ? 128 : (Rpad <= 192 ? 192 : 256); }

FW_HD TcLayout make_tc_layout(int BQ, int BK, int Rpad, bool bwd) {
    TcLayout y;
    y.Rpad = Rpad;
    y.RP = Rpad + 4;
    y.PS = BK + 1;
    y.HR = hist_rows(Rpad);
    y.BQm = 64;                                            // rows of an M = 64 A operand
    y.ro_row = 4 * CQ; y.ro_tab = 4 * CQ;
    y.ro_tq = (uint32_t)(BQ / 4) * CQ; y.ro_tk = (uint32_t)(BK / 4) * CQ;
    y.ro_hq = y.ro_tq; y.ro_hk = y.ro_tk;
    uint32_t o = 0;
    auto take = [&](uint32_t bytes) { uint32_t r = o; o += (bytes + 127u) & ~127u; return r; };
    const uint32_t tab_bytes = (uint32_t)(Rpad / 8) * y.ro_tab;
    for (int t = 0; t < 3; ++t)
        for (int p = 0; p < 2; ++p) y.tab[t][p] = (bwd || t < 2) ? take(tab_bytes) : 0;
    y.tvT = bwd ? 0 : take((uint32_t)HD * Rpad * 4);
    // R-form rows: contiguous, so that buffer B3 (the key-side histogram of the backward pass) can reuse the whole region
    const uint32_t rrow = 8 * y.ro_row;   // 64 rows
    y.qR[0] = take(rrow); y.qR[1] = take(rrow);
    y.kR[0] = take(rrow); y.kR[1] = take(rrow);
    y.gR[0] = y.gR[1] = y.vR[0] = y.vR[1] = 0;
    if (bwd) { y.gR[0] = take(rrow); y.gR[1] = take(rrow); y.vR[0] = take(rrow); y.vR[1] = take(rrow); }
    const uint32_t hist_k = (uint32_t)(y.HR / 8) * y.ro_hk;
    y.B3 = y.qR[0];
    if (bwd && hist_k > 8 * rrow) take(hist_k - 8 * rrow);
    const uint32_t tq_bytes = 2 * y.ro_tq, tk_bytes = 2 * y.ro_tk;   // 16 channels = 2 row octets
    for (int p = 0; p < 2; ++p) { y.qT[p] = bwd ? take(tq_bytes) : 0; y.gT[p] = bwd ? take(tq_bytes) : 0; y.kT[p] = bwd ? take(tk_bytes) : 0; }
    y.vplain = bwd ? 0 : take((uint32_t)BK * HD * 4);
    const uint32_t prod_q = (uint32_t)BQ * y.RP * 4, prod_k = (uint32_t)BK * y.RP * 4, hist_q = (uint32_t)(y.HR / 8) * y.ro_hq;
    if (bwd) {
        const uint32_t big = prod_q > hist_q ? (prod_q > prod_k ? prod_q : prod_k) : (hist_q > prod_k ? hist_q : prod_k);
        y.B1 = take(big); y.B2 = take(big);
    } else {
        const uint32_t phist = (uint32_t)Rpad * (BQ + 4) * 4;   // forward: plain histogram [Rpad][BQ+4], reuses the QT buffer
        y.B1 = take(prod_q > phist ? prod_q : phist); y.B2 = take(prod_k);
    }
    y.P = take((uint32_t)BQ * y.PS * 4);
    y.GS = bwd ? take((uint32_t)BQ * y.PS * 4) : 0;
    const uint32_t opart = bwd ? 2u * (BQ + 2 * BK) * HD * 4 : (uint32_t)KSF * BQ * HD * 4;
    const uint32_t rel = (uint32_t)BQ * y.PS * 4;
    y.U = take(opart > rel ? opart : rel);
    y.rowinfo = take(BQ * 16); y.keyid = take(BK * 16);
    y.red = take((BQ > BK ? BQ : BK) * 8 * 4); y.mrow = take(BQ * 4); y.lrow = take(BQ * 4); y.drow = take(BQ * 4);
    y.slot = take(64);
    y.total = o;
    return y;
}

FW_FN void put_hilo(unsigned char *sm, uint32_t hi, uint32_t lo, uint32_t off, float x) {
    const float h = tc::tf32_hi(x);
    *reinterpret_cast<float *>(sm + hi + off) = h;
    *reinterpret_cast<float *>(sm + lo + off) = x - h;
}
FW_FN void put_hilo4(unsigned char *sm, uint32_t hi, uint32_t lo, uint32_t off, float4 x) {
    const float4 h = make_float4(tc::tf32_hi(x.x), tc::tf32_hi(x.y), tc::tf32_hi(x.z), tc::tf32_hi(x.w));
    *reinterpret_cast<float4 *>(sm + hi + off) = h;
    *reinterpret_cast<float4 *>(sm + lo + off) = make_float4(x.x - h.x, x.y - h.y, x.z - h.z, x.w - h.w);
}

// table of one head, chunked [bin column][16] hi / lo, restricted to the staged bins (0 elsewhere)
FW_FN void stage_table_chunked(unsigned char *sm, uint32_t hi, uint32_t lo, uint32_t ro, const float *table, int head, int h, int L, int blo,
                               int RB, int Rpad, int tid) {
    for (int e = tid; e < Rpad * HD; e += NTC) {
        const int col = e / HD, ch = e - col * HD;
        const int a = col / RB, l = blo + (col - a * RB);
        float x = 0.f;
        if (a < 3 && l >= 0 && l < L) x = table[(((size_t)l * h + head) * HD + ch) * 3 + a];
        put_hilo(sm, hi, lo, tc::chunked_off(col, ch, ro, CQ), x);
    }
}
FW_FN void stage_table_plainT(float *dst, const float *table, int head, int h, int L, int blo, int RB, int Rpad, int tid) {
    for (int e = tid; e < HD * Rpad; e += NTC) {
        const int c = e / Rpad, col = e - c * Rpad;
        const int a = col / RB, l = blo + (col - a * RB);
        float x = 0.f;
        if (a < 3 && l >= 0 && l < L) x = table[(((size_t)l * h + head) * HD + c) * 3 + a];
        dst[e] = x;
    }
}

template <int BQ, int BK>
FW_FN void describe_item_tc(const PassParams &P, const Item &it, int4 *rowinfo, int4 *keyid, int tid) {
    for (int r = tid; r < BQ + BK; r += NTC) {
        if (r < BQ) {
            int4 ri = make_int4(-1, 0, 0, 0);
            if (r < it.nq) {
                const int pos = it.q_pos + r;
                ri.x = P.q_order[pos];
                if (it.flags & F_PACKED) {
                    const int win = P.pos_win[pos];
                    const int ws = P.wstart[win], we = P.wstart[win + 1];
                    ri.y = ws - it.k_pos; ri.z = we - it.k_pos;
                    ri.w = P.tile_base[win] + (pos - ws) * (we - ws);
                } else {
                    ri.y = 0; ri.z = it.nk;
                    ri.w = it.rel_off + r * it.rel_pitch;
                }
            }
            rowinfo[r] = ri;
        } else {
            const int j = r - BQ;
            int4 ki = make_int4(-1, 0, 0, 0);
            if (j < it.nk) {
                const int pos = it.k_pos + j;
                ki.x = P.k_order[pos];
                if (it.flags & F_PACKED) {
                    const int win = P.pos_win[pos];
                    ki.y = P.wstart[win] - it.q_pos; ki.z = P.wstart[win + 1] - it.q_pos;
                } else {
                    ki.y = 0; ki.z = it.nq;
                }
            }
            keyid[j] = ki;
        }
    }
}

// copy `ncols` accumulator columns starting at TMEM column c0 of an M = 64 accumulator into dst[row][ld] (row < rows).
// Warp w serves TMEM sub-partition w % 4 (rows 16 * (w % 4) .. + 15 in lanes 0..15) and the 16-column chunks w / 4, w / 4 + 4, ...
FW_FN void tmem_to_smem_m64(TcCtx &c, int tid, int c0, int ncols, int rows, float *dst, int ld, bool vec) {
    const int w = tid >> 5, l = tid & 31;
    const int r = 16 * (w & 3) + l;
    for (int ch = w >> 2; ch * 16 < ncols; ch += NTC / 128) {
        float v[16];
        tc_load16(c, tid, c0 + ch * 16, v);
        if (l < 16 && r < rows) {
            float *d = dst + (size_t)r * ld + ch * 16;
            if (vec) {
                FW_UNROLL
                for (int q = 0; q < 4; ++q) *reinterpret_cast<float4 *>(d + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            } else {
                FW_UNROLL
                for (int q = 0; q < 16; ++q)
                    if (ch * 16 + q < ncols) d[q] = v[q];
            }
        }
    }
}

// acc[r][n] += sum_{bins in [c0,c1)} H[bin][row0+r] * T[bin][4*cq+n]    H chunked [bin][rows], T chunked hi + lo [bin][16]
FW_FN void hist_x_table_ck(float (&acc)[4][4], const unsigned char *sm, uint32_t hist, uint32_t ro_h, int row0, uint32_t t_hi, uint32_t t_lo,
                           uint32_t ro_t, int cq, int c0, int c1) {
    for (int b = c0; b < c1; ++b) {
        const float4 hv = *reinterpret_cast<const float4 *>(sm + hist + tc::chunked_off(b, row0, ro_h, CQ));
        const uint32_t to = tc::chunked_off(b, 4 * cq, ro_t, CQ);
        const float4 th = *reinterpret_cast<const float4 *>(sm + t_hi + to), tl = *reinterpret_cast<const float4 *>(sm + t_lo + to);
        const float t[4] = {th.x + tl.x, th.y + tl.y, th.z + tl.z, th.w + tl.w};
        const float hr[4] = {hv.x, hv.y, hv.z, hv.w};
        FW_UNROLL
        for (int r = 0; r < 4; ++r)
            FW_UNROLL
            for (int n = 0; n < 4; ++n) acc[r][n] = fmaf(hr[r], t[n], acc[r][n]);
    }
}
// value of a T-form row array (hi + lo) at (channel c, row i)
FW_FN float tform(const unsigned char *sm, uint32_t hi, uint32_t lo, uint32_t ro, int c, int i) {
    const uint32_t o = tc::chunked_off(c, i, ro, CQ);
    return *reinterpret_cast<const float *>(sm + hi + o) + *reinterpret_cast<const float *>(sm + lo + o);
}

// ============================================================================================ forward
template <int BQ, int BK>
FW_FN void forward_cta_tc(const PassParams &P, int head, int cta, int n_cta, unsigned char *smb, TcCtx &ctx) {
    const TcLayout y = make_tc_layout(BQ, BK, P.Rpad, false);
    float *tvT = reinterpret_cast<float *>(smb + y.tvT), *vR = reinterpret_cast<float *>(smb + y.vplain);
    float *QT = reinterpret_cast<float *>(smb + y.B1), *KT = reinterpret_cast<float *>(smb + y.B2), *S = reinterpret_cast<float *>(smb + y.P);
    float *PhT = QT;
    const int PH = BQ + 4;
    unsigned *REL = reinterpret_cast<unsigned *>(smb + y.U);
    float *Opart = reinterpret_cast<float *>(smb + y.U);
    int4 *rowinfo = reinterpret_cast<int4 *>(smb + y.rowinfo);
    int4 *keyid = reinterpret_cast<int4 *>(smb + y.keyid);
    float *red = reinterpret_cast<float *>(smb + y.red), *mrow = reinterpret_cast<float *>(smb + y.mrow), *lrow = reinterpret_cast<float *>(smb + y.lrow);
    const int h = P.h, lo = P.bin_lo, RB = P.RB, Rpad = P.Rpad;
    const int c_qt = C_WORK, c_kt = C_WORK + Rpad, c_s = C_WORK + 2 * Rpad;
    TC_STAMP_DECL;

    TC_PHASE_BEGIN
        const int tid = FW_TID;
        stage_table_chunked(smb, y.tab[0][0], y.tab[0][1], y.ro_tab, P.tq, head, h, P.L, lo, RB, Rpad, tid);
        stage_table_chunked(smb, y.tab[1][0], y.tab[1][1], y.ro_tab, P.tk, head, h, P.L, lo, RB, Rpad, tid);
        stage_table_plainT(tvT, P.tv, head, h, P.L, lo, RB, Rpad, tid);
        tc_publish_smem();
    TC_PHASE_END

    for (int ii = cta; ii < P.n_items; ii += n_cta) {
        const Item it = P.items[ii];
        const int nq4 = round4(it.nq), nk4 = round4(it.nk);

        TC_PHASE_BEGIN   // ---- 0
            describe_item_tc<BQ, BK>(P, it, rowinfo, keyid, FW_TID);
        TC_PHASE_END
        TC_STAMP(0);

        TC_PHASE_BEGIN   // ---- 1: stage q, k as chunked hi / lo rows (64 rows each, padding rows zero), v plain, the rel tile
            const int tid = FW_TID;
            for (int e = tid; e < 64 * 4; e += NTC) {
                const int r = e >> 2, c4 = e & 3;
                const int pid = r < BQ ? rowinfo[r].x : -1;
                put_hilo4(smb, y.qR[0], y.qR[1], tc::chunked_off(r, 4 * c4, y.ro_row, CQ), load_row_quarter(P.q, pid, h, head, c4));
                if (pid >= 0) {
                    const int4 ri = rowinfo[r];
                    for (int t = c4; t < ri.z - ri.y; t += 4) REL[r * y.PS + ri.y + t] = P.rel[ri.w + t];
                }
            }
            for (int e = tid; e < 64 * 4; e += NTC) {
                const int j = e >> 2, c4 = e & 3;
                const int pid = j < BK ? keyid[j].x : -1;
                put_hilo4(smb, y.kR[0], y.kR[1], tc::chunked_off(j, 4 * c4, y.ro_row, CQ), load_row_quarter(P.k, pid, h, head, c4));
                if (j < BK) *reinterpret_cast<float4 *>(vR + j * HD + 4 * c4) = load_row_quarter(P.v, pid, h, head, c4);
            }
            tc_publish_smem();
        TC_PHASE_END
        TC_STAMP(1);

        TC_PHASE_BEGIN   // ---- 2: products and the q.k tile on the tensor cores
            const int tid = FW_TID;
            if (tid == 0) {
                tc_begin_issue();
                tc_gemm(ctx, c_qt, 64, Rpad, HD, y.qR[0], y.qR[1], y.ro_row, y.tab[0][0], y.tab[0][1], y.ro_tab, 7, false);
                tc_gemm(ctx, c_kt, 64, Rpad, HD, y.kR[0], y.kR[1], y.ro_row, y.tab[1][0], y.tab[1][1], y.ro_tab, 7, false);
                tc_gemm(ctx, c_s, 64, round16(BK), HD, y.qR[0], y.qR[1], y.ro_row, y.kR[0], y.kR[1], y.ro_row, 7, false);
                tc_commit(ctx);
            }
        TC_PHASE_END
        TC_STAMP(2);
        tc_wait(ctx);
        TC_STAMP(3);

        TC_PHASE_BEGIN   // ---- 3: accumulators -> shared memory
            const int tid = FW_TID;
            tmem_to_smem_m64(ctx, tid, c_qt, Rpad, nq4, QT, y.RP, true);
            tmem_to_smem_m64(ctx, tid, c_kt, Rpad, nk4, KT, y.RP, true);
            tmem_to_smem_m64(ctx, tid, c_s, nk4, nq4, S, y.PS, false);
            tc_tmem_reads_done();
        TC_PHASE_END
        TC_STAMP(4);

        TC_PHASE_BEGIN   // ---- 4: logits of the valid pairs, -inf elsewhere; partial row maxima (8 stripes per row)
            const int tid = FW_TID;
            for (int e = tid; e < nq4 * 8; e += NTC) {
                const int i = e >> 3, st = e & 7;
                const int4 ri = rowinfo[i];
                float mx = -INFINITY;
                for (int j = st; j < nk4; j += 8) {
                    float s = -INFINITY;
                    if (ri.x >= 0 && j >= ri.y && j < ri.z) {
                        const unsigned w = REL[i * y.PS + j];
                        if (!(w & REL_INVALID)) {
                            const int c0 = rel_col(w, 0, lo, RB), c1 = rel_col(w, 1, lo, RB), c2 = rel_col(w, 2, lo, RB);
                            const float *qt = QT + (size_t)i * y.RP, *kt = KT + (size_t)j * y.RP;
                            s = S[i * y.PS + j] + ((qt[c0] + qt[c1]) + qt[c2]) + ((kt[c0] + kt[c1]) + kt[c2]);
                        }
                    }
                    S[i * y.PS + j] = s;
                    mx = fmaxf(mx, s);
                }
                red[i * 8 + st] = mx;
            }
        TC_PHASE_END
        TC_STAMP(5);

        TC_PHASE_BEGIN   // ---- 5: row maxima; clear the histogram (the query products are dead now)
            const int tid = FW_TID;
            for (int i = tid; i < nq4; i += NTC) {
                float mx = red[i * 8];
                FW_UNROLL
                for (int s = 1; s < 8; ++s) mx = fmaxf(mx, red[i * 8 + s]);
                mrow[i] = mx;
            }
            for (int e = tid; e < Rpad * PH / 4; e += NTC) reinterpret_cast<float4 *>(PhT)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
        TC_PHASE_END
        TC_STAMP(6);

        TC_PHASE_BEGIN   // ---- 6: p = exp(s - max) in place, partial row sums
            const int tid = FW_TID;
            for (int e = tid; e < nq4 * 8; e += NTC) {
                const int i = e >> 3, st = e & 7;
                const float mx = mrow[i];
                float sum = 0.f;
                for (int j = st; j < nk4; j += 8) {
                    const float s = S[i * y.PS + j];
                    const float p = (s == -INFINITY) ? 0.f : expf(s - mx);
                    S[i * y.PS + j] = p;
                    sum += p;
                }
                red[i * 8 + st] = sum;
            }
        TC_PHASE_END
        TC_STAMP(7);

        TC_PHASE_BEGIN   // ---- 7: histogram Ph[bin][i] += p_ij (one thread per (row, axis)); row sums
            const int tid = FW_TID;
            for (int e = tid; e < nq4 * 3; e += NTC) {
                const int i = e / 3, a = e - 3 * i;
                const int4 ri = rowinfo[i];
                if (ri.x >= 0) {
                    for (int j = ri.y; j < ri.z; ++j) {
                        const float p = S[i * y.PS + j];
                        if (p != 0.f) PhT[(size_t)rel_col(REL[i * y.PS + j], a, lo, RB) * PH + i] += p;
                    }
                }
                if (a == 0) {
                    float l = 0.f;
                    FW_UNROLL
                    for (int s = 0; s < 8; ++s) l += red[i * 8 + s];
                    lrow[i] = l;
                }
            }
        TC_PHASE_END
        TC_STAMP(8);

        TC_PHASE_BEGIN   // ---- 8: out tile = P V + Ph T_v in KSF K-slices
            const int tid = FW_TID;
            const int nrq = nq4 / 4;
            for (int t = tid; t < nrq * 4 * KSF; t += NTC) {
                const int ks = t % KSF, cq = (t / KSF) & 3, rg = t / (KSF * 4);
                float acc[4][4];
                FW_UNROLL
                for (int r = 0; r < 4; ++r)
                    FW_UNROLL
                    for (int n = 0; n < 4; ++n) acc[r][n] = 0.f;
                const int jper = round4((nk4 + KSF - 1) / KSF);
                const int j0 = fw_min(ks * jper, nk4), j1 = fw_min(j0 + jper, nk4);
                for (int j = j0; j < j1; ++j) {
                    const float4 vv = *reinterpret_cast<const float4 *>(vR + j * HD + 4 * cq);
                    FW_UNROLL
                    for (int r = 0; r < 4; ++r) {
                        const float p = S[(4 * rg + r) * y.PS + j];
                        acc[r][0] = fmaf(p, vv.x, acc[r][0]);
                        acc[r][1] = fmaf(p, vv.y, acc[r][1]);
                        acc[r][2] = fmaf(p, vv.z, acc[r][2]);
                        acc[r][3] = fmaf(p, vv.w, acc[r][3]);
                    }
                }
                const int cper = round4((Rpad + KSF - 1) / KSF);
                const int c0 = fw_min(ks * cper, Rpad), c1 = fw_min(c0 + cper, Rpad);
                hist_x_table(acc, PhT, PH, 4 * rg, tvT, Rpad, cq, c0, c1);
                FW_UNROLL
                for (int r = 0; r < 4; ++r)
                    *reinterpret_cast<float4 *>(Opart + ((size_t)ks * BQ + 4 * rg + r) * HD + 4 * cq) =
                        make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
            }
            if (!(it.flags & F_FIRST))
                for (int i = tid; i < it.nq; i += NTC) {
                    const size_t rowh = (size_t)rowinfo[i].x * h + head;
                    red[i * 8] = P.m[rowh];
                    red[i * 8 + 1] = P.l[rowh];
                }
        TC_PHASE_END
        TC_STAMP(9);

        TC_PHASE_BEGIN   // ---- 9: merge with the partial of an earlier pass, normalise on the final pass, store
            const int tid = FW_TID;
            for (int e = tid; e < it.nq * 4; e += NTC) {
                const int i = e >> 2, c4 = e & 3;
                const int pid = rowinfo[i].x;
                float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                FW_UNROLL
                for (int ks = 0; ks < KSF; ++ks) {
                    const float4 x = *reinterpret_cast<const float4 *>(Opart + ((size_t)ks * BQ + i) * HD + 4 * c4);
                    o.x += x.x; o.y += x.y; o.z += x.z; o.w += x.w;
                }
                float mx = mrow[i], l = lrow[i];
                const size_t rowh = (size_t)pid * h + head;
                float *dst = P.out + rowh * HD + 4 * c4;
                if (!(it.flags & F_FIRST)) {
                    const float m0 = red[i * 8], l0 = red[i * 8 + 1];
                    const float4 o0 = *reinterpret_cast<const float4 *>(dst);
                    const float mn = fmaxf(m0, mx);
                    const float a0 = (m0 == -INFINITY) ? 0.f : expf(m0 - mn);
                    const float a1 = (mx == -INFINITY) ? 0.f : expf(mx - mn);
                    o = make_float4(a0 * o0.x + a1 * o.x, a0 * o0.y + a1 * o.y, a0 * o0.z + a1 * o.z, a0 * o0.w + a1 * o.w);
                    l = a0 * l0 + a1 * l;
                    mx = mn;
                }
                if (it.flags & F_FINAL) {
                    const float inv = l > 0.f ? 1.f / l : 0.f;
                    o = make_float4(o.x * inv, o.y * inv, o.z * inv, o.w * inv);
                }
                *reinterpret_cast<float4 *>(dst) = o;
                if (c4 == 0) {
                    P.m[rowh] = (it.flags & F_FINAL) ? mx + logf(l) : mx;
                    P.l[rowh] = l;
                }
            }
        TC_PHASE_END
        TC_STAMP(10);
    }
}

// ============================================================================================ backward
// HRT = histogram rows the instantiation is built for (128, 192 or 256 >= hist_rows(P.Rpad)): fixes how many histogram
// chunks a thread keeps in registers across the split
template <int BQ, int BK, int HRT>
FW_FN void backward_cta_tc(const PassParams &P, int head, int cta, int n_cta, unsigned char *smb, TcCtx &ctx) {
    const TcLayout y = make_tc_layout(BQ, BK, P.Rpad, true);
    float *B1f = reinterpret_cast<float *>(smb + y.B1), *B2f = reinterpret_cast<float *>(smb + y.B2);
    float *Pm = reinterpret_cast<float *>(smb + y.P), *GS = reinterpret_cast<float *>(smb + y.GS);
    unsigned *REL = reinterpret_cast<unsigned *>(smb + y.U);
    float *OQ = reinterpret_cast<float *>(smb + y.U), *OK = OQ + 2 * BQ * HD, *OV = OK + 2 * BK * HD;
    int4 *rowinfo = reinterpret_cast<int4 *>(smb + y.rowinfo);
    int4 *keyid = reinterpret_cast<int4 *>(smb + y.keyid);
    float *red = reinterpret_cast<float *>(smb + y.red), *lse = reinterpret_cast<float *>(smb + y.mrow), *drow = reinterpret_cast<float *>(smb + y.drow);
    const int h = P.h, lo = P.bin_lo, RB = P.RB, Rpad = P.Rpad;
    const int c_prod = C_WORK, c_s = C_WORK + Rpad, c_gv = c_s + round16(BK);
    // histogram chunks (float4) each thread splits / keeps the lo part of
    constexpr int HQ_CH = (HRT * BQ / 4 + NTC - 1) / NTC, HK_CH = (HRT * BK / 4 + NTC - 1) / NTC;
    TC_PER_THREAD(float, lo_sq, HQ_CH * 4);
    TC_PER_THREAD(float, lo_ph, HQ_CH * 4);
    TC_PER_THREAD(float, lo_sk, HK_CH * 4);
    bool q_acc_live = false, k_acc_live = false;   // table-gradient accumulators already hold a product (else the first MMA overwrites)
    const uint32_t hq_bytes = (uint32_t)(y.HR / 8) * y.ro_hq, hk_bytes = (uint32_t)(y.HR / 8) * y.ro_hk;
    TC_STAMP_DECL;

    TC_PHASE_BEGIN
        const int tid = FW_TID;
        stage_table_chunked(smb, y.tab[0][0], y.tab[0][1], y.ro_tab, P.tq, head, h, P.L, lo, RB, Rpad, tid);
        stage_table_chunked(smb, y.tab[1][0], y.tab[1][1], y.ro_tab, P.tk, head, h, P.L, lo, RB, Rpad, tid);
        stage_table_chunked(smb, y.tab[2][0], y.tab[2][1], y.ro_tab, P.tv, head, h, P.L, lo, RB, Rpad, tid);
        tc_publish_smem();
    TC_PHASE_END

    for (int ii = cta; ii < P.n_items; ii += n_cta) {
        const Item it = P.items[ii];
        const int nq4 = round4(it.nq), nk4 = round4(it.nk), nq8 = round8(it.nq), nk8 = round8(it.nk);

        TC_PHASE_BEGIN   // ---- 0
            describe_item_tc<BQ, BK>(P, it, rowinfo, keyid, FW_TID);
        TC_PHASE_END
        TC_STAMP(0);

        TC_PHASE_BEGIN   // ---- 1: stage q, g, k (R-form and T-form), v (R-form), all hi / lo; g.out partials, LSE, rel tile
            const int tid = FW_TID;
            for (int e = tid; e < 64 * 4; e += NTC) {
                const int r = e >> 2, c4 = e & 3;
                const int4 ri = r < BQ ? rowinfo[r] : make_int4(-1, 0, 0, 0);
                const float4 qq = load_row_quarter(P.q, ri.x, h, head, c4), gg = load_row_quarter(P.g, ri.x, h, head, c4);
                const uint32_t ro = tc::chunked_off(r, 4 * c4, y.ro_row, CQ);
                put_hilo4(smb, y.qR[0], y.qR[1], ro, qq);
                put_hilo4(smb, y.gR[0], y.gR[1], ro, gg);
                if (r < BQ) {
                    const float qa[4] = {qq.x, qq.y, qq.z, qq.w}, ga[4] = {gg.x, gg.y, gg.z, gg.w};
                    FW_UNROLL
                    for (int c = 0; c < 4; ++c) {
                        const uint32_t to = tc::chunked_off(4 * c4 + c, r, y.ro_tq, CQ);
                        put_hilo(smb, y.qT[0], y.qT[1], to, qa[c]);
                        put_hilo(smb, y.gT[0], y.gT[1], to, ga[c]);
                    }
                    const float4 oo = load_row_quarter(P.out, ri.x, h, head, c4);
                    red[r * 8 + c4] = (gg.x * oo.x + gg.y * oo.y) + (gg.z * oo.z + gg.w * oo.w);
                    if (c4 == 0) lse[r] = ri.x >= 0 ? P.lse[(size_t)ri.x * h + head] : 0.f;
                    if (ri.x >= 0)
                        for (int t = c4; t < ri.z - ri.y; t += 4) REL[r * y.PS + ri.y + t] = P.rel[ri.w + t];
                }
            }
            for (int e = tid; e < 64 * 4; e += NTC) {
                const int j = e >> 2, c4 = e & 3;
                const int pid = j < BK ? keyid[j].x : -1;
                const float4 kk = load_row_quarter(P.k, pid, h, head, c4);
                const uint32_t ro = tc::chunked_off(j, 4 * c4, y.ro_row, CQ);
                put_hilo4(smb, y.kR[0], y.kR[1], ro, kk);
                put_hilo4(smb, y.vR[0], y.vR[1], ro, load_row_quarter(P.v, pid, h, head, c4));
                if (j < BK) {
                    const float ka[4] = {kk.x, kk.y, kk.z, kk.w};
                    FW_UNROLL
                    for (int c = 0; c < 4; ++c) put_hilo(smb, y.kT[0], y.kT[1], tc::chunked_off(4 * c4 + c, j, y.ro_tk, CQ), ka[c]);
                }
            }
            tc_publish_smem();
        TC_PHASE_END
        TC_STAMP(1);

        TC_PHASE_BEGIN   // ---- 2: QT product and both tiles
            const int tid = FW_TID;
            if (tid == 0) {
                tc_begin_issue();
                tc_gemm(ctx, c_prod, 64, Rpad, HD, y.qR[0], y.qR[1], y.ro_row, y.tab[0][0], y.tab[0][1], y.ro_tab, 7, false);
                tc_gemm(ctx, c_s, 64, round16(BK), HD, y.qR[0], y.qR[1], y.ro_row, y.kR[0], y.kR[1], y.ro_row, 7, false);
                tc_gemm(ctx, c_gv, 64, round16(BK), HD, y.gR[0], y.gR[1], y.ro_row, y.vR[0], y.vR[1], y.ro_row, 7, false);
                tc_commit(ctx);
            }
            for (int i = tid; i < BQ; i += NTC) drow[i] = (red[i * 8] + red[i * 8 + 1]) + (red[i * 8 + 2] + red[i * 8 + 3]);
        TC_PHASE_END
        TC_STAMP(2);
        tc_wait(ctx);
        TC_STAMP(3);

        TC_PHASE_BEGIN   // ---- 3
            const int tid = FW_TID;
            tmem_to_smem_m64(ctx, tid, c_prod, Rpad, nq4, B1f, y.RP, true);
            tmem_to_smem_m64(ctx, tid, c_s, nk4, nq4, Pm, y.PS, false);
            tmem_to_smem_m64(ctx, tid, c_gv, nk4, nq4, GS, y.PS, false);
            tc_tmem_reads_done();
        TC_PHASE_END
        TC_STAMP(4);

        TC_PHASE_BEGIN   // ---- 4: KT product
            if (FW_TID == 0) {
                tc_begin_issue();
                tc_gemm(ctx, c_prod, 64, Rpad, HD, y.kR[0], y.kR[1], y.ro_row, y.tab[1][0], y.tab[1][1], y.ro_tab, 7, false);
                tc_commit(ctx);
            }
        TC_PHASE_END
        TC_STAMP(5);
        tc_wait(ctx);
        TC_STAMP(6);

        TC_PHASE_BEGIN   // ---- 5
            tmem_to_smem_m64(ctx, FW_TID, c_prod, Rpad, nk4, B2f, y.RP, true);
            tc_tmem_reads_done();
        TC_PHASE_END
        TC_STAMP(7);

        TC_PHASE_BEGIN   // ---- 6: GT product in flight while p = exp(s - LSE) is computed
            const int tid = FW_TID;
            if (tid == 0) {
                tc_begin_issue();
                tc_gemm(ctx, c_prod, 64, Rpad, HD, y.gR[0], y.gR[1], y.ro_row, y.tab[2][0], y.tab[2][1], y.ro_tab, 7, false);
                tc_commit(ctx);
            }
            for (int e = tid; e < nq4 * 8; e += NTC) {
                const int i = e >> 3, st = e & 7;
                const int4 ri = rowinfo[i];
                const float ls = lse[i];
                for (int j = st; j < nk4; j += 8) {
                    float p = 0.f;
                    if (ri.x >= 0 && j >= ri.y && j < ri.z) {
                        const unsigned w = REL[i * y.PS + j];
                        if (!(w & REL_INVALID)) {
                            const int c0 = rel_col(w, 0, lo, RB), c1 = rel_col(w, 1, lo, RB), c2 = rel_col(w, 2, lo, RB);
                            const float *qt = B1f + (size_t)i * y.RP, *kt = B2f + (size_t)j * y.RP;
                            p = expf(Pm[i * y.PS + j] + ((qt[c0] + qt[c1]) + qt[c2]) + ((kt[c0] + kt[c1]) + kt[c2]) - ls);
                        }
                    }
                    Pm[i * y.PS + j] = p;
                }
            }
        TC_PHASE_END
        TC_STAMP(8);
        tc_wait(ctx);
        TC_STAMP(9);

        TC_PHASE_BEGIN   // ---- 7
            tmem_to_smem_m64(ctx, FW_TID, c_prod, Rpad, nq4, B1f, y.RP, true);
            tc_tmem_reads_done();
        TC_PHASE_END
        TC_STAMP(10);

        TC_PHASE_BEGIN   // ---- 8: gs = p (g.v + GT look-ups - g.out)
            const int tid = FW_TID;
            for (int e = tid; e < nq4 * 8; e += NTC) {
                const int i = e >> 3, st = e & 7;
                const float dd = drow[i];
                for (int j = st; j < nk4; j += 8) {
                    float gs = 0.f;
                    const float p = Pm[i * y.PS + j];
                    if (p != 0.f) {
                        const unsigned w = REL[i * y.PS + j];
                        const float *gt = B1f + (size_t)i * y.RP;
                        const float gp = GS[i * y.PS + j] + ((gt[rel_col(w, 0, lo, RB)] + gt[rel_col(w, 1, lo, RB)]) + gt[rel_col(w, 2, lo, RB)]);
                        gs = p * (gp - dd);
                    }
                    GS[i * y.PS + j] = gs;
                }
            }
        TC_PHASE_END
        TC_STAMP(11);

        TC_PHASE_BEGIN   // ---- 9: clear the three histograms: Sq -> B1, Ph -> B2 ([bin][query row]), Sk -> B3 ([bin][key row], over the dead R-form rows)
            const int tid = FW_TID;
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            for (uint32_t e = tid; e < hq_bytes / 16; e += NTC) {
                reinterpret_cast<float4 *>(smb + y.B1)[e] = z;
                reinterpret_cast<float4 *>(smb + y.B2)[e] = z;
            }
            for (uint32_t e = tid; e < hk_bytes / 16; e += NTC) reinterpret_cast<float4 *>(smb + y.B3)[e] = z;
        TC_PHASE_END
        TC_STAMP(12);

        TC_PHASE_BEGIN   // ---- 10: build them (thread per (query row, axis), thread per (key row, axis))
            const int tid = FW_TID;
            const int wq = nq4 * 3, wk = nk4 * 3;
            for (int e = tid; e < wq + wk; e += NTC) {
                if (e < wq) {
                    const int i = e / 3, a = e - 3 * i;
                    const int4 ri = rowinfo[i];
                    if (ri.x < 0) continue;
                    for (int j = ri.y; j < ri.z; ++j) {
                        const float p = Pm[i * y.PS + j];
                        if (p == 0.f) continue;
                        const uint32_t o = tc::chunked_off(rel_col(REL[i * y.PS + j], a, lo, RB), i, y.ro_hq, CQ);
                        *reinterpret_cast<float *>(smb + y.B1 + o) += GS[i * y.PS + j];
                        *reinterpret_cast<float *>(smb + y.B2 + o) += p;
                    }
                } else {
                    const int j = (e - wq) / 3, a = (e - wq) - 3 * j;
                    const int4 ki = keyid[j];
                    if (ki.x < 0) continue;
                    for (int i = ki.y; i < ki.z; ++i) {
                        if (Pm[i * y.PS + j] == 0.f) continue;
                        *reinterpret_cast<float *>(smb + y.B3 + tc::chunked_off(rel_col(REL[i * y.PS + j], a, lo, RB), j, y.ro_hk, CQ)) += GS[i * y.PS + j];
                    }
                }
            }
        TC_PHASE_END
        TC_STAMP(13);

        TC_PHASE_BEGIN   // ---- 11: gq, gk, gv tiles on the FMA pipe (two K-halves each); the rel tile is dead, its space takes the partials
            const int tid = FW_TID;
            const int nrq = nq4 / 4, nrk = nk4 / 4;
            const int uq = nrq * 4 * 2, uk = nrk * 4 * 2;
            for (int t = tid; t < uq + 2 * uk; t += NTC) {
                float acc[4][4];
                FW_UNROLL
                for (int r = 0; r < 4; ++r)
                    FW_UNROLL
                    for (int n = 0; n < 4; ++n) acc[r][n] = 0.f;
                float *dst;
                if (t < uq) {            // gq[i] = sum_j gs_ij k_j + Sq[i] . T_q
                    const int kh = t & 1, cq = (t >> 1) & 3, rg = t >> 3;
                    const int jh = round4(nk4 / 2), ch = round4(Rpad / 2);
                    for (int j = kh ? jh : 0; j < (kh ? nk4 : jh); ++j) {
                        float x[4];
                        FW_UNROLL
                        for (int n = 0; n < 4; ++n) x[n] = tform(smb, y.kT[0], y.kT[1], y.ro_tk, 4 * cq + n, j);
                        FW_UNROLL
                        for (int r = 0; r < 4; ++r) {
                            const float w = GS[(4 * rg + r) * y.PS + j];
                            FW_UNROLL
                            for (int n = 0; n < 4; ++n) acc[r][n] = fmaf(w, x[n], acc[r][n]);
                        }
                    }
                    hist_x_table_ck(acc, smb, y.B1, y.ro_hq, 4 * rg, y.tab[0][0], y.tab[0][1], y.ro_tab, cq, kh ? ch : 0, kh ? Rpad : ch);
                    dst = OQ + ((size_t)kh * BQ + 4 * rg) * HD + 4 * cq;
                } else if (t < uq + uk) {   // gk[j] = sum_i gs_ij q_i + Sk[j] . T_k
                    const int u = t - uq;
                    const int kh = u & 1, cq = (u >> 1) & 3, rg = u >> 3;
                    const int ih = round4(nq4 / 2), ch = round4(Rpad / 2);
                    for (int i = kh ? ih : 0; i < (kh ? nq4 : ih); ++i) {
                        float x[4];
                        FW_UNROLL
                        for (int n = 0; n < 4; ++n) x[n] = tform(smb, y.qT[0], y.qT[1], y.ro_tq, 4 * cq + n, i);
                        FW_UNROLL
                        for (int r = 0; r < 4; ++r) {
                            const float w = GS[i * y.PS + 4 * rg + r];
                            FW_UNROLL
                            for (int n = 0; n < 4; ++n) acc[r][n] = fmaf(w, x[n], acc[r][n]);
                        }
                    }
                    hist_x_table_ck(acc, smb, y.B3, y.ro_hk, 4 * rg, y.tab[1][0], y.tab[1][1], y.ro_tab, cq, kh ? ch : 0, kh ? Rpad : ch);
                    dst = OK + ((size_t)kh * BK + 4 * rg) * HD + 4 * cq;
                } else {                    // gv[j] = sum_i p_ij g_i
                    const int u = t - uq - uk;
                    const int kh = u & 1, cq = (u >> 1) & 3, rg = u >> 3;
                    const int ih = round4(nq4 / 2);
                    for (int i = kh ? ih : 0; i < (kh ? nq4 : ih); ++i) {
                        float x[4];
                        FW_UNROLL
                        for (int n = 0; n < 4; ++n) x[n] = tform(smb, y.gT[0], y.gT[1], y.ro_tq, 4 * cq + n, i);
                        FW_UNROLL
                        for (int r = 0; r < 4; ++r) {
                            const float w = Pm[i * y.PS + 4 * rg + r];
                            FW_UNROLL
                            for (int n = 0; n < 4; ++n) acc[r][n] = fmaf(w, x[n], acc[r][n]);
                        }
                    }
                    dst = OV + ((size_t)kh * BK + 4 * rg) * HD + 4 * cq;
                }
                FW_UNROLL
                for (int r = 0; r < 4; ++r) *reinterpret_cast<float4 *>(dst + r * HD) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
            }
        TC_PHASE_END
        TC_STAMP(14);

        TC_PHASE_BEGIN   // ---- 12: split the histograms in place (hi stays, lo in registers); write the gradient rows
            const int tid = FW_TID;
            TC_PER_THREAD_USE(float, lo_sq);
            TC_PER_THREAD_USE(float, lo_ph);
            TC_PER_THREAD_USE(float, lo_sk);
            FW_UNROLL
            for (int c = 0; c < HQ_CH; ++c) {
                const uint32_t e = (uint32_t)tid + (uint32_t)c * NTC;
                if (e < hq_bytes / 16) {
                    float4 *a = reinterpret_cast<float4 *>(smb + y.B1) + e, *b = reinterpret_cast<float4 *>(smb + y.B2) + e;
                    const float4 x = *a, z = *b;
                    const float4 xh = make_float4(tc::tf32_hi(x.x), tc::tf32_hi(x.y), tc::tf32_hi(x.z), tc::tf32_hi(x.w));
                    const float4 zh = make_float4(tc::tf32_hi(z.x), tc::tf32_hi(z.y), tc::tf32_hi(z.z), tc::tf32_hi(z.w));
                    *a = xh; *b = zh;
                    lo_sq[4 * c] = x.x - xh.x; lo_sq[4 * c + 1] = x.y - xh.y; lo_sq[4 * c + 2] = x.z - xh.z; lo_sq[4 * c + 3] = x.w - xh.w;
                    lo_ph[4 * c] = z.x - zh.x; lo_ph[4 * c + 1] = z.y - zh.y; lo_ph[4 * c + 2] = z.z - zh.z; lo_ph[4 * c + 3] = z.w - zh.w;
                }
            }
            FW_UNROLL
            for (int c = 0; c < HK_CH; ++c) {
                const uint32_t e = (uint32_t)tid + (uint32_t)c * NTC;
                if (e < hk_bytes / 16) {
                    float4 *a = reinterpret_cast<float4 *>(smb + y.B3) + e;
                    const float4 x = *a;
                    const float4 xh = make_float4(tc::tf32_hi(x.x), tc::tf32_hi(x.y), tc::tf32_hi(x.z), tc::tf32_hi(x.w));
                    *a = xh;
                    lo_sk[4 * c] = x.x - xh.x; lo_sk[4 * c + 1] = x.y - xh.y; lo_sk[4 * c + 2] = x.z - xh.z; lo_sk[4 * c + 3] = x.w - xh.w;
                }
            }
            tc_publish_smem();
            for (int e = tid; e < it.nq * 4; e += NTC) {
                const int i = e >> 2, c4 = e & 3;
                const float4 a = *reinterpret_cast<const float4 *>(OQ + (size_t)i * HD + 4 * c4);
                const float4 b = *reinterpret_cast<const float4 *>(OQ + ((size_t)BQ + i) * HD + 4 * c4);
                float4 o = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
                float *dst = P.gq + ((size_t)rowinfo[i].x * h + head) * HD + 4 * c4;
                if (!(it.flags & F_FIRST)) {
                    const float4 x = *reinterpret_cast<const float4 *>(dst);
                    o = make_float4(o.x + x.x, o.y + x.y, o.z + x.z, o.w + x.w);
                }
                *reinterpret_cast<float4 *>(dst) = o;
            }
            for (int e = tid; e < it.nk * 4; e += NTC) {
                const int j = e >> 2, c4 = e & 3;
                const size_t off = ((size_t)keyid[j].x * h + head) * HD + 4 * c4;
                FW_UNROLL
                for (int which = 0; which < 2; ++which) {
                    const float *src = which ? OV : OK;
                    float *dst = (which ? P.gv : P.gk) + off;
                    const float4 a = *reinterpret_cast<const float4 *>(src + (size_t)j * HD + 4 * c4);
                    const float4 b = *reinterpret_cast<const float4 *>(src + ((size_t)BK + j) * HD + 4 * c4);
                    const float4 o = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
                    if (it.flags & F_KEY_ATOMIC) {
                        atomic_add_f(dst, o.x); atomic_add_f(dst + 1, o.y); atomic_add_f(dst + 2, o.z); atomic_add_f(dst + 3, o.w);
                    } else {
                        *reinterpret_cast<float4 *>(dst) = o;
                    }
                }
            }
        TC_PHASE_END
        TC_STAMP(15);

        // ---- 13 / 14: table gradients on the tensor cores: gT_q += Sq^T Q, gT_v += Ph^T G, gT_k += Sk^T K.
        // A = histogram [bin][row] (M = 128 bins per MMA, K = rows), B = T-form rows [16][row] (N = 16).
        for (int round = 0; round < 2; ++round) {
            TC_PHASE_BEGIN
                const int tid = FW_TID;
                if (round == 1) {   // the lo parts take over the histogram buffers
                    TC_PER_THREAD_USE(float, lo_sq);
                    TC_PER_THREAD_USE(float, lo_ph);
                    TC_PER_THREAD_USE(float, lo_sk);
                    FW_UNROLL
                    for (int c = 0; c < HQ_CH; ++c) {
                        const uint32_t e = (uint32_t)tid + (uint32_t)c * NTC;
                        if (e < hq_bytes / 16) {
                            reinterpret_cast<float4 *>(smb + y.B1)[e] = make_float4(lo_sq[4 * c], lo_sq[4 * c + 1], lo_sq[4 * c + 2], lo_sq[4 * c + 3]);
                            reinterpret_cast<float4 *>(smb + y.B2)[e] = make_float4(lo_ph[4 * c], lo_ph[4 * c + 1], lo_ph[4 * c + 2], lo_ph[4 * c + 3]);
                        }
                    }
                    FW_UNROLL
                    for (int c = 0; c < HK_CH; ++c) {
                        const uint32_t e = (uint32_t)tid + (uint32_t)c * NTC;
                        if (e < hk_bytes / 16)
                            reinterpret_cast<float4 *>(smb + y.B3)[e] = make_float4(lo_sk[4 * c], lo_sk[4 * c + 1], lo_sk[4 * c + 2], lo_sk[4 * c + 3]);
                    }
                    tc_publish_smem();
                }
            TC_PHASE_END
            TC_STAMP(16);
            TC_PHASE_BEGIN
                if (FW_TID == 0) {
                    tc_begin_issue();
                    // round 0: H_hi * (X_hi + X_lo); round 1: H_lo * X_hi   (the buffer holds H_hi resp. H_lo as the "hi" operand)
                    const int terms = round == 0 ? 3 : 1;
                    const bool acc_q = q_acc_live || round == 1, acc_k = k_acc_live || round == 1;
                    for (int blk = 0; blk * 128 < y.HR; ++blk) {
                        const int M = (y.HR - blk * 128) >= 128 ? 128 : 64;
                        const int col = C_GT + blk * 48;
                        const uint32_t aq = (uint32_t)(blk * 16) * y.ro_hq, ak = (uint32_t)(blk * 16) * y.ro_hk;
                        if ((P.dbg & 1) && nk8 > 0) tc_gemm(ctx, col + 16, M, 16, nk8, y.B3 + ak, y.B3 + ak, y.ro_hk, y.kT[0], y.kT[1], y.ro_tk, terms, acc_k);
                        if (nq8 > 0) {
                            tc_gemm(ctx, col, M, 16, nq8, y.B1 + aq, y.B1 + aq, y.ro_hq, y.qT[0], y.qT[1], y.ro_tq, terms, acc_q);
                            tc_gemm(ctx, col + 32, M, 16, nq8, y.B2 + aq, y.B2 + aq, y.ro_hq, y.gT[0], y.gT[1], y.ro_tq, terms, acc_q);
                        }
                        if (!(P.dbg & 1) && nk8 > 0) tc_gemm(ctx, col + 16, M, 16, nk8, y.B3 + ak, y.B3 + ak, y.ro_hk, y.kT[0], y.kT[1], y.ro_tk, terms, acc_k);
                    }
                    tc_commit(ctx);
                }
            TC_PHASE_END
            TC_STAMP(17);
            tc_wait(ctx);
            TC_STAMP(18);
        }
        q_acc_live = q_acc_live || nq8 > 0;
        k_acc_live = k_acc_live || nk8 > 0;
    }

    TC_PHASE_BEGIN   // flush the table gradients of this CTA: gT[l][head][c][a] += accumulator
        const int tid = FW_TID;
        {
            const int w = tid >> 5, l = tid & 31;
            const int lane = 32 * (w & 3) + l, which = w >> 2;   // warps 0-3: T_q, 4-7: T_k, 8-11: T_v
            for (int blk = 0; blk * 128 < y.HR; ++blk) {
                const int M = (y.HR - blk * 128) >= 128 ? 128 : 64;
                float v[16];
                tc_load16(ctx, tid, C_GT + blk * 48 + (which < 3 ? which : 0) * 16, v);
                int col = -1;
                if (M == 128) col = blk * 128 + lane;
                else if (l < 16) col = blk * 128 + 16 * (w & 3) + l;
                if (which < 3 && (which == 1 ? k_acc_live : q_acc_live) && col >= 0 && col < Rpad) {
                    const int a = col / RB, bin = lo + (col - a * RB);
                    if (a < 3 && bin >= 0 && bin < P.L) {
                        float *gt = which == 0 ? P.gtq : (which == 1 ? P.gtk : P.gtv);
                        FW_UNROLL
                        for (int c = 0; c < 16; ++c)
                            if (v[c] != 0.f) atomic_add_f(gt + (((size_t)bin * h + head) * HD + c) * 3 + a, v[c]);
                    }
                }
            }
        }
        tc_tmem_reads_done();
    TC_PHASE_END
    TC_STAMP(19);
}

}  // namespace fwtc
}  // namespace stb200
