// Window-centric fused attention: the body one CTA executes, written as a sequence of barrier-separated phases.
//
// What it replaces (reference, paths under /root/reference): the whole pair path of WindowAttention.forward,
// model/stratified_transformer.py:183-210 (attention_step1_v2 + dot_prod_with_idx_v3 + add + scatter_softmax +
// attention_step2_with_rel_pos_value_v2) and its autograd backward, for pair lists produced by
// get_indice_pairs (:10-42): dense pairs inside each small window, sparse pairs query -> FPS-sampled keys of its 2x window.
//
// Formulation.  A work item is a block (<= BQ query rows) x (<= BK key rows) of the pair matrix:
//   dense pass : queries = keys = the points of one small window (several small windows packed per item)
//   sparse pass: queries = a chunk of the points of one large window, keys = its sampled points, pairs masked by
//                the window_coord test (flag bit of the rel word)
// and per item, for ONE head (tables of that head stay in shared memory for the CTA's lifetime):
//   products  QT[i][a,l] = q_i . T_q[l,:,a]   KT[j][a,l] = k_j . T_k[l,:,a]   (GT[i][a,l] = g_i . T_v[l,:,a] backward)
//             as register-tiled fp32 GEMMs over the rows of the item, only for the bin range the pass can reach
//   logits    s_ij = q_i.k_j + sum_a QT[i][a,r_a] + KT[j][a,r_a]      (6 scalar look-ups instead of 6 table rows)
//   softmax   per pass partial (max, sum, unnormalised out); the second pass merges and normalises, LSE is kept
//   aggregate out_i = sum_j p_ij v_j + sum_{a,l} Ph[i][a,l] T_v[l,:,a],  Ph = per-row histogram of p over the bins
//   backward  p = exp(s - LSE), gp = g_i.v_j + GT look-ups, gs = p (gp - g_i.out_i); histograms Sq, Sk, Ph of gs / p
//             turn every table term into a small GEMM (gq += Sq T_q, gk += Sk T_k, gT_q += Sq^T Q, ...)
// No [M,h] tensor exists; per pair the kernels read one 32-bit word (three bins + the mask flag).
//
// The same source is compiled for the host by tests/emu (FW_HOST_EMU): a phase becomes a loop over thread ids, so the
// index arithmetic, the buffer aliasing and the math are checked on the CPU against the oracle.  No thread
// communicates inside a phase (no shuffles); everything crosses phases through shared memory.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__) && !defined(FW_HOST_EMU)
#define FW_FN __device__ __forceinline__
#define FW_HD __host__ __device__ inline
#define FW_PHASE_BEGIN {
#define FW_PHASE_END } __syncthreads();
#define FW_TID ((int)threadIdx.x)
#define FW_PER_THREAD(type, name, n) type name[n]
#define FW_PER_THREAD_USE(type, name)
#define FW_UNROLL _Pragma("unroll")
#else
#include <vector_types.h>
#define FW_FN inline
#define FW_HD inline
#define FW_PHASE_BEGIN for (int fw_tid_ = 0; fw_tid_ < NT; ++fw_tid_) {
#define FW_PHASE_END }
#define FW_TID fw_tid_
#define FW_PER_THREAD(type, name, n) static thread_local type name##_all[NT][n]
#define FW_PER_THREAD_USE(type, name) type *name = name##_all[fw_tid_]
#define FW_UNROLL
#endif

namespace stb200 {
namespace fw {

constexpr int NT = 256;   // threads per CTA
constexpr int HD = 16;    // head dim of the fused path (all shipped configs; other head dims use the per-op kernels)
constexpr int KS = 4;     // K-slices of the [rows x 16] output GEMMs

enum : int { F_PACKED = 1, F_FIRST = 2, F_FINAL = 4, F_KEY_ATOMIC = 8 };
constexpr unsigned REL_INVALID = 0x80000000u;

struct Item {          // 32 bytes
    int q_pos, nq;     // query rows = q_order[q_pos .. q_pos+nq)
    int k_pos, nk;     // key rows   = k_order[k_pos .. k_pos+nk)
    int rel_off;       // rel words of row r, key j: rel[rel_off + r*rel_pitch + j]   (not PACKED)
    int rel_pitch;
    int flags;
    int pad;
};

struct PassParams {
    const Item *items;
    int n_items;
    const int *q_order, *k_order;   // sorted position -> point id
    const unsigned *rel;
    const int *pos_win;             // PACKED items: small-window rank of a sorted position
    const int *wstart;              //               window boundaries in sorted positions [n_win+1]
    const int *tile_base;           //               first rel word of a window's [n x n] tile [n_win]
    int bin_lo, RB, Rpad, L, h;
    const float *q, *k, *v;         // [N,h,16]
    const float *tq, *tk, *tv;      // [L,h,16,3]
    float *out, *m, *l;             // forward: out [N,h,16] (partial, then final), m/l [N,h] (final pass leaves LSE in m)
    const float *g, *lse;           // backward: grad_out [N,h,16], LSE [N,h] (= m of the forward), out = forward output
    float *gq, *gk, *gv;            // [N,h,16]
    float *gtq, *gtk, *gtv;         // [L,h,16,3], accumulated into
    int dbg;                        // development switches (STB200_FUSED_DBG), 0 in production
    long long *prof;                // optional [64] per-phase cycle totals of CTA 0 (stb200_fused_phase_profile), else NULL
};

struct Layout {   // offsets in floats into the CTA's shared memory
    int Rpad, RP, PT, PTK, PS, PH, PHK, PBUF_Q, PBUF_K;
    int tq, tk, tv, qT, kT, gT, vT, vR, QT, KT, GT, S, GS, U, rowinfo, keyid, red, mrow, lrow, drow, total;
};

FW_HD int fw_max(int a, int b) { return a > b ? a : b; }
FW_HD int fw_min(int a, int b) { return a < b ? a : b; }
FW_HD int round4(int x) { return (x + 3) & ~3; }

FW_HD Layout make_layout(int BQ, int BK, int Rpad, bool bwd) {
    Layout y;
    y.Rpad = Rpad;
    y.RP = Rpad + 4;                 // product row pitch: multiple of 4 (vector stores), = 4 mod 8
    y.PT = BQ + 4;                   // pitch of the transposed row arrays [16][rows]
    y.PTK = BK + 4;
    y.PS = BK + 1;                   // tile pitch (odd: rows of a tile land in different banks)
    y.PH = BQ + 4;                   // histogram pitch [Rpad][rows]
    y.PHK = BK + 4;
    y.PBUF_Q = fw_max(BQ * y.RP, Rpad * y.PH);   // a product buffer is later reused as the histogram of the same rows
    y.PBUF_K = fw_max(BK * y.RP, Rpad * y.PHK);
    int o = 0;
    auto take = [&](int n) { int r = o; o += (n + 3) & ~3; return r; };
    y.tq = take(HD * Rpad); y.tk = take(HD * Rpad); y.tv = take(HD * Rpad);
    y.qT = take(HD * y.PT); y.kT = take(HD * y.PTK);
    y.gT = bwd ? take(HD * y.PT) : 0;
    y.vT = bwd ? take(HD * y.PTK) : 0;
    y.vR = bwd ? 0 : take(BK * HD);
    y.QT = take(y.PBUF_Q); y.KT = take(y.PBUF_K);
    y.GT = bwd ? take(y.PBUF_Q) : 0;
    y.S = take(BQ * y.PS);
    y.GS = bwd ? take(BQ * y.PS) : 0;
    // union: rel tile (until the histograms are built) / K-slice partial sums of the output GEMMs
    const int opart = bwd ? 2 * (BQ + 2 * BK) * HD : KS * BQ * HD;
    y.U = take(fw_max(BQ * y.PS, opart));
    y.rowinfo = take(BQ * 4); y.keyid = take(BK * 4);
    y.red = take(fw_max(BQ, BK) * 4); y.mrow = take(BQ); y.lrow = take(BQ); y.drow = take(BQ);
    y.total = o;
    return y;
}

FW_FN float4 ld4(const float *p) {
#if defined(__CUDACC__) && !defined(FW_HOST_EMU)
    return __ldg(reinterpret_cast<const float4 *>(p));
#else
    return *reinterpret_cast<const float4 *>(p);
#endif
}
FW_FN void atomic_add_f(float *p, float v) {
#if defined(__CUDACC__) && !defined(FW_HOST_EMU)
    atomicAdd(p, v);
#else
    *p += v;
#endif
}
FW_FN int rel_col(unsigned w, int a, int lo, int RB) {   // column of axis a's bin inside a product row
    const int b = (int)((w >> (8 * a)) & 0xffu) - lo;
    return a * RB + fw_min(fw_max(b, 0), RB - 1);
}

// tables of one head, transposed and restricted to the staged bin range: dst[c][a*RB + (l-lo)] = table[l][head][c][a]
FW_FN void stage_table(float *dst, const float *table, int head, int h, int L, int lo, int RB, int Rpad, int tid) {
    for (int e = tid; e < HD * Rpad; e += NT) {
        const int c = e / Rpad, col = e - c * Rpad;
        const int a = col / RB, l = lo + (col - a * RB);
        float x = 0.f;
        if (a < 3 && l >= 0 && l < L) x = table[(((size_t)l * h + head) * HD + c) * 3 + a];
        dst[e] = x;
    }
}

// C[i][col] = sum_c AT[c][i] * BT[c][col], i < rows (multiple of 4 after padding), col < Rpad (multiple of 8); K = 16.
// One 4x8 output tile per call.
FW_FN void product_tile(const float *AT, int lda, const float *BT, int Rpad, float *C, int ldc, int rg, int cg) {
    float acc[4][8];
    FW_UNROLL
    for (int r = 0; r < 4; ++r)
        FW_UNROLL
        for (int n = 0; n < 8; ++n) acc[r][n] = 0.f;
    FW_UNROLL
    for (int c = 0; c < HD; ++c) {
        const float4 a = *reinterpret_cast<const float4 *>(AT + c * lda + 4 * rg);
        const float4 b0 = *reinterpret_cast<const float4 *>(BT + c * Rpad + 8 * cg);
        const float4 b1 = *reinterpret_cast<const float4 *>(BT + c * Rpad + 8 * cg + 4);
        const float av[4] = {a.x, a.y, a.z, a.w};
        const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        FW_UNROLL
        for (int r = 0; r < 4; ++r)
            FW_UNROLL
            for (int n = 0; n < 8; ++n) acc[r][n] = fmaf(av[r], bv[n], acc[r][n]);
    }
    FW_UNROLL
    for (int r = 0; r < 4; ++r) {
        float *dst = C + (size_t)(4 * rg + r) * ldc + 8 * cg;
        *reinterpret_cast<float4 *>(dst) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
        *reinterpret_cast<float4 *>(dst + 4) = make_float4(acc[r][4], acc[r][5], acc[r][6], acc[r][7]);
    }
}

// C[i][j] = sum_c AT[c][i] * BT[c][j]: one 4x4 tile, scalar stores (tile pitch is odd)
FW_FN void dot_tile(const float *AT, int lda, const float *BT, int ldb, float *C, int ldc, int rg, int cg) {
    float acc[4][4];
    FW_UNROLL
    for (int r = 0; r < 4; ++r)
        FW_UNROLL
        for (int n = 0; n < 4; ++n) acc[r][n] = 0.f;
    FW_UNROLL
    for (int c = 0; c < HD; ++c) {
        const float4 a = *reinterpret_cast<const float4 *>(AT + c * lda + 4 * rg);
        const float4 b = *reinterpret_cast<const float4 *>(BT + c * ldb + 4 * cg);
        const float av[4] = {a.x, a.y, a.z, a.w};
        const float bv[4] = {b.x, b.y, b.z, b.w};
        FW_UNROLL
        for (int r = 0; r < 4; ++r)
            FW_UNROLL
            for (int n = 0; n < 4; ++n) acc[r][n] = fmaf(av[r], bv[n], acc[r][n]);
    }
    FW_UNROLL
    for (int r = 0; r < 4; ++r)
        FW_UNROLL
        for (int n = 0; n < 4; ++n) C[(4 * rg + r) * ldc + 4 * cg + n] = acc[r][n];
}

// acc[r][n] += sum_{k in [k0,k1)} W[row0+r][k] * XT[4*cq+n][k]   (W: tile, row-major, odd pitch; XT: [16][rows] transposed rows)
FW_FN void tile_rows_x_colsT(float (&acc)[4][4], const float *W, int ldw, int row0, const float *XT, int ldx, int cq, int k0, int k1) {
    for (int k = k0; k < k1; ++k) {
        float w[4], x[4];
        FW_UNROLL
        for (int r = 0; r < 4; ++r) w[r] = W[(row0 + r) * ldw + k];
        FW_UNROLL
        for (int n = 0; n < 4; ++n) x[n] = XT[(4 * cq + n) * ldx + k];
        FW_UNROLL
        for (int r = 0; r < 4; ++r)
            FW_UNROLL
            for (int n = 0; n < 4; ++n) acc[r][n] = fmaf(w[r], x[n], acc[r][n]);
    }
}
// acc[r][n] += sum_{k in [k0,k1)} W[k][col0+r] * XT[4*cq+n][k]     (transposed use of the tile: its columns are the output rows)
FW_FN void tile_cols_x_colsT(float (&acc)[4][4], const float *W, int ldw, int col0, const float *XT, int ldx, int cq, int k0, int k1) {
    for (int k = k0; k < k1; ++k) {
        float w[4], x[4];
        FW_UNROLL
        for (int r = 0; r < 4; ++r) w[r] = W[k * ldw + col0 + r];
        FW_UNROLL
        for (int n = 0; n < 4; ++n) x[n] = XT[(4 * cq + n) * ldx + k];
        FW_UNROLL
        for (int r = 0; r < 4; ++r)
            FW_UNROLL
            for (int n = 0; n < 4; ++n) acc[r][n] = fmaf(w[r], x[n], acc[r][n]);
    }
}
// acc[r][n] += sum_{c' in [c0,c1)} HT[c'][row0+r] * TT[4*cq+n][c']   (histogram [Rpad][rows] times transposed table [16][Rpad]); c0, c1 multiples of 4
FW_FN void hist_x_table(float (&acc)[4][4], const float *HT, int ldh, int row0, const float *TT, int Rpad, int cq, int c0, int c1) {
    for (int cc = c0; cc < c1; cc += 4) {
        float4 hrow[4], tcol[4];
        FW_UNROLL
        for (int e = 0; e < 4; ++e) hrow[e] = *reinterpret_cast<const float4 *>(HT + (size_t)(cc + e) * ldh + row0);   // 4 rows at c'=cc+e
        FW_UNROLL
        for (int n = 0; n < 4; ++n) tcol[n] = *reinterpret_cast<const float4 *>(TT + (4 * cq + n) * Rpad + cc);       // 4 c' at channel n
        const float hv[4][4] = {{hrow[0].x, hrow[0].y, hrow[0].z, hrow[0].w}, {hrow[1].x, hrow[1].y, hrow[1].z, hrow[1].w},
                                {hrow[2].x, hrow[2].y, hrow[2].z, hrow[2].w}, {hrow[3].x, hrow[3].y, hrow[3].z, hrow[3].w}};
        const float tv[4][4] = {{tcol[0].x, tcol[0].y, tcol[0].z, tcol[0].w}, {tcol[1].x, tcol[1].y, tcol[1].z, tcol[1].w},
                                {tcol[2].x, tcol[2].y, tcol[2].z, tcol[2].w}, {tcol[3].x, tcol[3].y, tcol[3].z, tcol[3].w}};
        FW_UNROLL
        for (int e = 0; e < 4; ++e)
            FW_UNROLL
            for (int r = 0; r < 4; ++r)
                FW_UNROLL
                for (int n = 0; n < 4; ++n) acc[r][n] = fmaf(hv[e][r], tv[n][e], acc[r][n]);
    }
}
// table gradient tile: acc[t][n] += sum_{i in [i0,i1)} HT[4*mg+t][i] * XT[4*ng+n][i]; i0, i1 multiples of 4
FW_FN void hist_x_rows(float *acc /*[16]*/, const float *HT, int ldh, int mg, const float *XT, int ldx, int ng, int i0, int i1) {
    for (int i = i0; i < i1; i += 4) {
        float4 hrow[4], xrow[4];
        FW_UNROLL
        for (int t = 0; t < 4; ++t) hrow[t] = *reinterpret_cast<const float4 *>(HT + (size_t)(4 * mg + t) * ldh + i);
        FW_UNROLL
        for (int n = 0; n < 4; ++n) xrow[n] = *reinterpret_cast<const float4 *>(XT + (4 * ng + n) * ldx + i);
        FW_UNROLL
        for (int t = 0; t < 4; ++t)
            FW_UNROLL
            for (int n = 0; n < 4; ++n) {
                float s = acc[t * 4 + n];
                s = fmaf(hrow[t].x, xrow[n].x, s);
                s = fmaf(hrow[t].y, xrow[n].y, s);
                s = fmaf(hrow[t].z, xrow[n].z, s);
                s = fmaf(hrow[t].w, xrow[n].w, s);
                acc[t * 4 + n] = s;
            }
    }
}

// Row / key descriptors of an item (phase 0 of both directions).
//   rowinfo[r] = {point id (-1: padding), first key, one past last key, first rel word of the row's valid keys}
//   keyid[j]   = {point id (-1: padding), first query row, one past last query row, unused}
template <int BQ, int BK>
FW_FN void describe_item(const PassParams &P, const Item &it, int4 *rowinfo, int4 *keyid, int tid) {
    for (int r = tid; r < BQ; r += NT) {
        int4 ri = make_int4(-1, 0, 0, 0);
        if (r < it.nq) {
            const int pos = it.q_pos + r;
            ri.x = P.q_order[pos];
            if (it.flags & F_PACKED) {
                const int win = P.pos_win[pos];
                const int ws = P.wstart[win], we = P.wstart[win + 1];
                ri.y = ws - it.k_pos;
                ri.z = we - it.k_pos;
                ri.w = P.tile_base[win] + (pos - ws) * (we - ws);
            } else {
                ri.y = 0;
                ri.z = it.nk;
                ri.w = it.rel_off + r * it.rel_pitch;
            }
        }
        rowinfo[r] = ri;
    }
    for (int j = tid; j < BK; j += NT) {
        int4 ki = make_int4(-1, 0, 0, 0);
        if (j < it.nk) {
            const int pos = it.k_pos + j;
            ki.x = P.k_order[pos];
            if (it.flags & F_PACKED) {
                const int win = P.pos_win[pos];
                ki.y = P.wstart[win] - it.q_pos;
                ki.z = P.wstart[win + 1] - it.q_pos;
            } else {
                ki.y = 0;
                ki.z = it.nq;
            }
        }
        keyid[j] = ki;
    }
}

// global row [16] of point `pid` -> column `r` of a transposed array [16][ld]; quarter c4 (4 channels) per call
FW_FN float4 load_row_quarter(const float *base, int pid, int h, int head, int c4) {
    return pid >= 0 ? ld4(base + ((size_t)pid * h + head) * HD + 4 * c4) : make_float4(0.f, 0.f, 0.f, 0.f);
}
FW_FN void put_transposed(float *T, int ld, int r, int c4, float4 x) {
    T[(4 * c4 + 0) * ld + r] = x.x;
    T[(4 * c4 + 1) * ld + r] = x.y;
    T[(4 * c4 + 2) * ld + r] = x.z;
    T[(4 * c4 + 3) * ld + r] = x.w;
}

// ============================================================================================ forward
template <int BQ, int BK>
FW_FN void forward_cta(const PassParams &P, int head, int cta, int n_cta, float *sm) {
    const Layout y = make_layout(BQ, BK, P.Rpad, false);
    float *tqT = sm + y.tq, *tkT = sm + y.tk, *tvT = sm + y.tv;
    float *qT = sm + y.qT, *kT = sm + y.kT, *vR = sm + y.vR;
    float *QT = sm + y.QT, *KT = sm + y.KT, *S = sm + y.S;
    float *PhT = QT;   // histogram of p over the bins, [Rpad][PH], reuses the query product buffer
    unsigned *REL = reinterpret_cast<unsigned *>(sm + y.U);
    float *Opart = sm + y.U;
    int4 *rowinfo = reinterpret_cast<int4 *>(sm + y.rowinfo);
    int4 *keyid = reinterpret_cast<int4 *>(sm + y.keyid);
    float *red = sm + y.red, *mrow = sm + y.mrow, *lrow = sm + y.lrow;
    const int h = P.h, lo = P.bin_lo, RB = P.RB, Rpad = P.Rpad;

    FW_PHASE_BEGIN
        const int tid = FW_TID;
        stage_table(tqT, P.tq, head, h, P.L, lo, RB, Rpad, tid);
        stage_table(tkT, P.tk, head, h, P.L, lo, RB, Rpad, tid);
        stage_table(tvT, P.tv, head, h, P.L, lo, RB, Rpad, tid);
    FW_PHASE_END

    for (int ii = cta; ii < P.n_items; ii += n_cta) {
        const Item it = P.items[ii];
        const int nq4 = round4(it.nq), nk4 = round4(it.nk);

        FW_PHASE_BEGIN   // ---- 0: row / key descriptors
            describe_item<BQ, BK>(P, it, rowinfo, keyid, FW_TID);
        FW_PHASE_END

        FW_PHASE_BEGIN   // ---- 1: stage q (transposed), k (transposed), v (row major), the rel tile
            const int tid = FW_TID;
            for (int e = tid; e < BQ * 4; e += NT) {
                const int r = e >> 2, c4 = e & 3;
                put_transposed(qT, y.PT, r, c4, load_row_quarter(P.q, rowinfo[r].x, h, head, c4));
                const int4 ri = rowinfo[r];
                if (ri.x >= 0)
                    for (int t = c4; t < ri.z - ri.y; t += 4) REL[r * y.PS + ri.y + t] = P.rel[ri.w + t];
            }
            for (int e = tid; e < BK * 4; e += NT) {
                const int j = e >> 2, c4 = e & 3;
                const int pid = keyid[j].x;
                put_transposed(kT, y.PTK, j, c4, load_row_quarter(P.k, pid, h, head, c4));
                *reinterpret_cast<float4 *>(vR + j * HD + 4 * c4) = load_row_quarter(P.v, pid, h, head, c4);
            }
        FW_PHASE_END

        FW_PHASE_BEGIN   // ---- 2: products QT, KT and the q.k tile
            const int tid = FW_TID;
            const int ncg = Rpad / 8, nrq = nq4 / 4, nrk = nk4 / 4;
            const int t_q = nrq * ncg, t_k = nrk * ncg, t_s = nrq * nrk;
            for (int t = tid; t < t_q + t_k + t_s; t += NT) {
                if (t < t_q) product_tile(qT, y.PT, tqT, Rpad, QT, y.RP, t % nrq, t / nrq);
                else if (t < t_q + t_k) product_tile(kT, y.PTK, tkT, Rpad, KT, y.RP, (t - t_q) % nrk, (t - t_q) / nrk);
                else dot_tile(qT, y.PT, kT, y.PTK, S, y.PS, (t - t_q - t_k) % nrq, (t - t_q - t_k) / nrq);
            }
        FW_PHASE_END

        FW_PHASE_BEGIN   // ---- 3: logits of the valid pairs, -inf elsewhere; partial row maxima (4 stripes per row)
            const int tid = FW_TID;
            for (int e = tid; e < nq4 * 4; e += NT) {
                const int i = e >> 2, st = e & 3;
                const int4 ri = rowinfo[i];
                float mx = -INFINITY;
                for (int j = st; j < nk4; j += 4) {
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
                red[i * 4 + st] = mx;
            }
        FW_PHASE_END

        FW_PHASE_BEGIN   // ---- 4: row maxima; clear the histogram (the query products are dead now)
            const int tid = FW_TID;
            for (int i = tid; i < nq4; i += NT) mrow[i] = fmaxf(fmaxf(red[i * 4], red[i * 4 + 1]), fmaxf(red[i * 4 + 2], red[i * 4 + 3]));
            for (int e = tid; e < Rpad * y.PH / 4; e += NT) reinterpret_cast<float4 *>(PhT)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
        FW_PHASE_END

        FW_PHASE_BEGIN   // ---- 5: p = exp(s - max) in place, partial row sums
            const int tid = FW_TID;
            for (int e = tid; e < nq4 * 4; e += NT) {
                const int i = e >> 2, st = e & 3;
                const float mx = mrow[i];
                float sum = 0.f;
                for (int j = st; j < nk4; j += 4) {
                    const float s = S[i * y.PS + j];
                    const float p = (s == -INFINITY) ? 0.f : expf(s - mx);
                    S[i * y.PS + j] = p;
                    sum += p;
                }
                red[i * 4 + st] = sum;
            }
        FW_PHASE_END

        FW_PHASE_BEGIN   // ---- 6: histogram Ph[a,l][i] += p_ij (one thread per (row, axis)); row sums
            const int tid = FW_TID;
            for (int e = tid; e < nq4 * 3; e += NT) {
                const int i = e / 3, a = e - 3 * i;
                const int4 ri = rowinfo[i];
                if (ri.x >= 0) {
                    for (int j = ri.y; j < ri.z; ++j) {
                        const float p = S[i * y.PS + j];
                        if (p != 0.f) PhT[(size_t)rel_col(REL[i * y.PS + j], a, lo, RB) * y.PH + i] += p;
                    }
                }
                if (a == 0) lrow[i] = (red[i * 4] + red[i * 4 + 1]) + (red[i * 4 + 2] + red[i * 4 + 3]);
            }
        FW_PHASE_END

        FW_PHASE_BEGIN   // ---- 7: out tile = P V + Ph T_v, K split in KS slices (the rel tile is dead: its space holds the partials)
            const int tid = FW_TID;
            const int nrq = nq4 / 4;
            for (int t = tid; t < nrq * 4 * KS; t += NT) {
                const int ks = t % KS, cq = (t / KS) & 3, rg = t / (KS * 4);
                float acc[4][4];
                FW_UNROLL
                for (int r = 0; r < 4; ++r)
                    FW_UNROLL
                    for (int n = 0; n < 4; ++n) acc[r][n] = 0.f;
                const int jper = round4((nk4 + KS - 1) / KS);
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
                const int cper = round4((Rpad + KS - 1) / KS);
                const int c0 = fw_min(ks * cper, Rpad), c1 = fw_min(c0 + cper, Rpad);
                hist_x_table(acc, PhT, y.PH, 4 * rg, tvT, Rpad, cq, c0, c1);
                FW_UNROLL
                for (int r = 0; r < 4; ++r)
                    *reinterpret_cast<float4 *>(Opart + ((size_t)ks * BQ + 4 * rg + r) * HD + 4 * cq) =
                        make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
            }
            if (!(it.flags & F_FIRST))   // the partial (max, sum) of the earlier pass: read here, overwritten in phase 8
                for (int i = tid; i < it.nq; i += NT) {
                    const size_t rowh = (size_t)rowinfo[i].x * h + head;
                    red[i * 4] = P.m[rowh];
                    red[i * 4 + 1] = P.l[rowh];
                }
        FW_PHASE_END

        FW_PHASE_BEGIN   // ---- 8: merge with the partial of an earlier pass, normalise on the final pass, store
            const int tid = FW_TID;
            for (int e = tid; e < it.nq * 4; e += NT) {
                const int i = e >> 2, c4 = e & 3;
                const int pid = rowinfo[i].x;
                float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                FW_UNROLL
                for (int ks = 0; ks < KS; ++ks) {
                    const float4 x = *reinterpret_cast<const float4 *>(Opart + ((size_t)ks * BQ + i) * HD + 4 * c4);
                    o.x += x.x; o.y += x.y; o.z += x.z; o.w += x.w;
                }
                float mx = mrow[i], l = lrow[i];
                const size_t rowh = (size_t)pid * h + head;
                float *dst = P.out + rowh * HD + 4 * c4;
                if (!(it.flags & F_FIRST)) {
                    const float m0 = red[i * 4], l0 = red[i * 4 + 1];
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
                    if (it.flags & F_FINAL) {
                        P.m[rowh] = mx + logf(l);   // log-sum-exp of the row, what the backward pass needs
                        P.l[rowh] = l;
                    } else {
                        P.m[rowh] = mx;
                        P.l[rowh] = l;
                    }
                }
            }
        FW_PHASE_END
    }
}

// ============================================================================================ backward
template <int BQ, int BK>
FW_FN void backward_cta(const PassParams &P, int head, int cta, int n_cta, float *sm) {
    const Layout y = make_layout(BQ, BK, P.Rpad, true);
    float *tqT = sm + y.tq, *tkT = sm + y.tk, *tvT = sm + y.tv;
    float *qT = sm + y.qT, *kT = sm + y.kT, *gT = sm + y.gT, *vT = sm + y.vT;
    float *QT = sm + y.QT, *KT = sm + y.KT, *GT = sm + y.GT;
    float *Pm = sm + y.S, *GS = sm + y.GS;          // tiles: q.k -> p, g.v -> gs
    float *SqT = QT, *SkT = KT, *PhT = GT;           // histograms reuse the product buffers
    unsigned *REL = reinterpret_cast<unsigned *>(sm + y.U);
    float *OQ = sm + y.U, *OK = OQ + 2 * BQ * HD, *OV = OK + 2 * BK * HD;   // two K-halves each
    int4 *rowinfo = reinterpret_cast<int4 *>(sm + y.rowinfo);
    int4 *keyid = reinterpret_cast<int4 *>(sm + y.keyid);
    float *red = sm + y.red, *lse = sm + y.mrow, *drow = sm + y.drow;
    const int h = P.h, lo = P.bin_lo, RB = P.RB, Rpad = P.Rpad;
    // table-gradient accumulators: every thread owns one 4x4 tile of [Rpad x 16] per table (and one half of the rows
    // when 2 * tiles <= NT), for all items of this CTA
    const int tg_tiles = (Rpad / 4) * 4;
    const int tg_split = (2 * tg_tiles <= NT) ? 2 : 1;
    FW_PER_THREAD(float, gacc, 48);

    FW_PHASE_BEGIN
        const int tid = FW_TID;
        FW_PER_THREAD_USE(float, gacc);
        stage_table(tqT, P.tq, head, h, P.L, lo, RB, Rpad, tid);
        stage_table(tkT, P.tk, head, h, P.L, lo, RB, Rpad, tid);
        stage_table(tvT, P.tv, head, h, P.L, lo, RB, Rpad, tid);
        FW_UNROLL
        for (int e = 0; e < 48; ++e) gacc[e] = 0.f;
    FW_PHASE_END

    for (int ii = cta; ii < P.n_items; ii += n_cta) {
        const Item it = P.items[ii];
        const int nq4 = round4(it.nq), nk4 = round4(it.nk);

        FW_PHASE_BEGIN   // ---- 0
            describe_item<BQ, BK>(P, it, rowinfo, keyid, FW_TID);
        FW_PHASE_END

        FW_PHASE_BEGIN   // ---- 1: stage q, g (+ partial g.out), k, v transposed; rel tile; LSE
            const int tid = FW_TID;
            for (int e = tid; e < BQ * 4; e += NT) {
                const int r = e >> 2, c4 = e & 3;
                const int4 ri = rowinfo[r];
                put_transposed(qT, y.PT, r, c4, load_row_quarter(P.q, ri.x, h, head, c4));
                const float4 gg = load_row_quarter(P.g, ri.x, h, head, c4);
                const float4 oo = load_row_quarter(P.out, ri.x, h, head, c4);
                put_transposed(gT, y.PT, r, c4, gg);
                red[r * 4 + c4] = (gg.x * oo.x + gg.y * oo.y) + (gg.z * oo.z + gg.w * oo.w);
                if (c4 == 0) lse[r] = ri.x >= 0 ? P.lse[(size_t)ri.x * h + head] : 0.f;
                if (ri.x >= 0)
                    for (int t = c4; t < ri.z - ri.y; t += 4) REL[r * y.PS + ri.y + t] = P.rel[ri.w + t];
            }
            for (int e = tid; e < BK * 4; e += NT) {
                const int j = e >> 2, c4 = e & 3;
                const int pid = keyid[j].x;
                put_transposed(kT, y.PTK, j, c4, load_row_quarter(P.k, pid, h, head, c4));
                put_transposed(vT, y.PTK, j, c4, load_row_quarter(P.v, pid, h, head, c4));
            }
        FW_PHASE_END

        FW_PHASE_BEGIN   // ---- 2: products QT, KT, GT; tiles q.k and g.v; D_i = g_i . out_i
            const int tid = FW_TID;
            const int ncg = Rpad / 8, nrq = nq4 / 4, nrk = nk4 / 4;
            const int t_q = nrq * ncg, t_k = nrk * ncg, t_s = nrq * nrk;
            for (int t = tid; t < 2 * t_q + t_k + 2 * t_s; t += NT) {
                int u = t;
                if (u < t_q) { product_tile(qT, y.PT, tqT, Rpad, QT, y.RP, u % nrq, u / nrq); continue; }
                u -= t_q;
                if (u < t_q) { product_tile(gT, y.PT, tvT, Rpad, GT, y.RP, u % nrq, u / nrq); continue; }
                u -= t_q;
                if (u < t_k) { product_tile(kT, y.PTK, tkT, Rpad, KT, y.RP, u % nrk, u / nrk); continue; }
                u -= t_k;
                if (u < t_s) { dot_tile(qT, y.PT, kT, y.PTK, Pm, y.PS, u % nrq, u / nrq); continue; }
                u -= t_s;
                dot_tile(gT, y.PT, vT, y.PTK, GS, y.PS, u % nrq, u / nrq);
            }
            for (int i = tid; i < BQ; i += NT) drow[i] = (red[i * 4] + red[i * 4 + 1]) + (red[i * 4 + 2] + red[i * 4 + 3]);
        FW_PHASE_END

        FW_PHASE_BEGIN   // ---- 3: p and gs of every pair of the block (zero outside the valid pairs)
            const int tid = FW_TID;
            for (int e = tid; e < nq4 * 4; e += NT) {
                const int i = e >> 2, st = e & 3;
                const int4 ri = rowinfo[i];
                const float ls = lse[i], dd = drow[i];
                for (int j = st; j < nk4; j += 4) {
                    float p = 0.f, gs = 0.f;
                    if (ri.x >= 0 && j >= ri.y && j < ri.z) {
                        const unsigned w = REL[i * y.PS + j];
                        if (!(w & REL_INVALID)) {
                            const int c0 = rel_col(w, 0, lo, RB), c1 = rel_col(w, 1, lo, RB), c2 = rel_col(w, 2, lo, RB);
                            const float *qt = QT + (size_t)i * y.RP, *kt = KT + (size_t)j * y.RP, *gt = GT + (size_t)i * y.RP;
                            const float s = Pm[i * y.PS + j] + ((qt[c0] + qt[c1]) + qt[c2]) + ((kt[c0] + kt[c1]) + kt[c2]);
                            p = expf(s - ls);
                            const float gp = GS[i * y.PS + j] + ((gt[c0] + gt[c1]) + gt[c2]);
                            gs = p * (gp - dd);
                        }
                    }
                    Pm[i * y.PS + j] = p;
                    GS[i * y.PS + j] = gs;
                }
            }
        FW_PHASE_END

        FW_PHASE_BEGIN   // ---- 4: clear the three histograms (the products are dead)
            const int tid = FW_TID;
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int e = tid; e < Rpad * y.PH / 4; e += NT) {
                reinterpret_cast<float4 *>(SqT)[e] = z;
                reinterpret_cast<float4 *>(PhT)[e] = z;
            }
            for (int e = tid; e < Rpad * y.PHK / 4; e += NT) reinterpret_cast<float4 *>(SkT)[e] = z;
        FW_PHASE_END

        FW_PHASE_BEGIN   // ---- 5: histograms Sq, Ph (thread per (query row, axis)) and Sk (thread per (key row, axis))
            const int tid = FW_TID;
            const int wq = nq4 * 3, wk = nk4 * 3;
            for (int e = tid; e < wq + wk; e += NT) {
                if (e < wq) {
                    const int i = e / 3, a = e - 3 * i;
                    const int4 ri = rowinfo[i];
                    if (ri.x < 0) continue;
                    for (int j = ri.y; j < ri.z; ++j) {
                        const unsigned w = REL[i * y.PS + j];
                        if (w & REL_INVALID) continue;
                        const size_t col = (size_t)rel_col(w, a, lo, RB) * y.PH + i;
                        SqT[col] += GS[i * y.PS + j];
                        PhT[col] += Pm[i * y.PS + j];
                    }
                } else {
                    const int j = (e - wq) / 3, a = (e - wq) - 3 * j;
                    const int4 ki = keyid[j];
                    if (ki.x < 0) continue;
                    for (int i = ki.y; i < ki.z; ++i) {
                        const unsigned w = REL[i * y.PS + j];
                        if (w & REL_INVALID) continue;
                        SkT[(size_t)rel_col(w, a, lo, RB) * y.PHK + j] += GS[i * y.PS + j];
                    }
                }
            }
        FW_PHASE_END

        FW_PHASE_BEGIN   // ---- 6: gq, gk, gv tiles (two K-halves each) and the table-gradient tiles
            const int tid = FW_TID;
            FW_PER_THREAD_USE(float, gacc);
            const int nrq = nq4 / 4, nrk = nk4 / 4;
            const int uq = nrq * 4 * 2, uk = nrk * 4 * 2;
            for (int t = tid; t < uq + 2 * uk; t += NT) {
                float acc[4][4];
                FW_UNROLL
                for (int r = 0; r < 4; ++r)
                    FW_UNROLL
                    for (int n = 0; n < 4; ++n) acc[r][n] = 0.f;
                float *dst;
                if (t < uq) {            // gq[i] = sum_j gs_ij k_j + Sq[i] . T_q
                    const int kh = t & 1, cq = (t >> 1) & 3, rg = t >> 3;
                    const int jh = round4(nk4 / 2), ch = round4(Rpad / 2);
                    tile_rows_x_colsT(acc, GS, y.PS, 4 * rg, kT, y.PTK, cq, kh ? jh : 0, kh ? nk4 : jh);
                    hist_x_table(acc, SqT, y.PH, 4 * rg, tqT, Rpad, cq, kh ? ch : 0, kh ? Rpad : ch);
                    dst = OQ + ((size_t)kh * BQ + 4 * rg) * HD + 4 * cq;
                } else if (t < uq + uk) {   // gk[j] = sum_i gs_ij q_i + Sk[j] . T_k
                    const int u = t - uq;
                    const int kh = u & 1, cq = (u >> 1) & 3, rg = u >> 3;
                    const int ih = round4(nq4 / 2), ch = round4(Rpad / 2);
                    tile_cols_x_colsT(acc, GS, y.PS, 4 * rg, qT, y.PT, cq, kh ? ih : 0, kh ? nq4 : ih);
                    hist_x_table(acc, SkT, y.PHK, 4 * rg, tkT, Rpad, cq, kh ? ch : 0, kh ? Rpad : ch);
                    dst = OK + ((size_t)kh * BK + 4 * rg) * HD + 4 * cq;
                } else {                    // gv[j] = sum_i p_ij g_i
                    const int u = t - uq - uk;
                    const int kh = u & 1, cq = (u >> 1) & 3, rg = u >> 3;
                    const int ih = round4(nq4 / 2);
                    tile_cols_x_colsT(acc, Pm, y.PS, 4 * rg, gT, y.PT, cq, kh ? ih : 0, kh ? nq4 : ih);
                    dst = OV + ((size_t)kh * BK + 4 * rg) * HD + 4 * cq;
                }
                FW_UNROLL
                for (int r = 0; r < 4; ++r) *reinterpret_cast<float4 *>(dst + r * HD) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
            }
            if (tid < tg_tiles * tg_split) {   // gT_q += Sq^T Q, gT_k += Sk^T K, gT_v += Ph^T G
                const int tile = tid % tg_tiles, half = tid / tg_tiles;
                const int mg = tile >> 2, ng = tile & 3;
                const int qh = tg_split == 2 ? round4(nq4 / 2) : nq4, kh = tg_split == 2 ? round4(nk4 / 2) : nk4;
                const int qi0 = half ? qh : 0, qi1 = (tg_split == 2 && !half) ? qh : nq4;
                const int ki0 = half ? kh : 0, ki1 = (tg_split == 2 && !half) ? kh : nk4;
                hist_x_rows(gacc, SqT, y.PH, mg, qT, y.PT, ng, qi0, qi1);
                hist_x_rows(gacc + 16, SkT, y.PHK, mg, kT, y.PTK, ng, ki0, ki1);
                hist_x_rows(gacc + 32, PhT, y.PH, mg, gT, y.PT, ng, qi0, qi1);
            }
        FW_PHASE_END

        FW_PHASE_BEGIN   // ---- 7: write the gradient rows
            const int tid = FW_TID;
            for (int e = tid; e < it.nq * 4; e += NT) {
                const int i = e >> 2, c4 = e & 3;
                const float4 a = *reinterpret_cast<const float4 *>(OQ + (size_t)i * HD + 4 * c4);
                const float4 b = *reinterpret_cast<const float4 *>(OQ + ((size_t)BQ + i) * HD + 4 * c4);
                float4 o = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
                float *dst = P.gq + ((size_t)rowinfo[i].x * h + head) * HD + 4 * c4;
                if (!(it.flags & F_FIRST)) {   // rows of an item are owned by it inside one launch: plain read-modify-write
                    const float4 x = *reinterpret_cast<const float4 *>(dst);
                    o = make_float4(o.x + x.x, o.y + x.y, o.z + x.z, o.w + x.w);
                }
                *reinterpret_cast<float4 *>(dst) = o;
            }
            for (int e = tid; e < it.nk * 4; e += NT) {
                const int j = e >> 2, c4 = e & 3;
                const size_t off = ((size_t)keyid[j].x * h + head) * HD + 4 * c4;
                FW_UNROLL
                for (int which = 0; which < 2; ++which) {
                    const float *src = which ? OV : OK;
                    float *dst = (which ? P.gv : P.gk) + off;
                    const float4 a = *reinterpret_cast<const float4 *>(src + (size_t)j * HD + 4 * c4);
                    const float4 b = *reinterpret_cast<const float4 *>(src + ((size_t)BK + j) * HD + 4 * c4);
                    const float4 o = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
                    if (it.flags & F_KEY_ATOMIC) {   // key rows shared with other items of this launch / an earlier pass
                        atomic_add_f(dst, o.x); atomic_add_f(dst + 1, o.y); atomic_add_f(dst + 2, o.z); atomic_add_f(dst + 3, o.w);
                    } else {
                        *reinterpret_cast<float4 *>(dst) = o;
                    }
                }
            }
        FW_PHASE_END
    }

    FW_PHASE_BEGIN   // flush the table gradients of this CTA: gT[l][head][c][a] += acc
        const int tid = FW_TID;
        FW_PER_THREAD_USE(float, gacc);
        if (tid < tg_tiles * tg_split) {
            const int tile = tid % tg_tiles;
            const int mg = tile >> 2, ng = tile & 3;
            FW_UNROLL
            for (int which = 0; which < 3; ++which) {
                float *gt = which == 0 ? P.gtq : (which == 1 ? P.gtk : P.gtv);
                FW_UNROLL
                for (int t = 0; t < 4; ++t) {
                    const int col = 4 * mg + t;
                    const int a = col / RB, l = lo + (col - a * RB);
                    if (a >= 3 || l < 0 || l >= P.L) continue;
                    FW_UNROLL
                    for (int n = 0; n < 4; ++n) {
                        const float x = gacc[which * 16 + t * 4 + n];
                        if (x != 0.f) atomic_add_f(gt + (((size_t)l * h + head) * HD + 4 * ng + n) * 3 + a, x);
                    }
                }
            }
        }
    FW_PHASE_END
}

}  // namespace fw
}  // namespace stb200
