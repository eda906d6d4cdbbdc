// Variable-length window attention over a CSR pair list: the v2/v3 pointops2 operators, fwd + bwd.
//
// What is computed is specified by the reference kernels (paths under /root/reference/lib/pointops2/src):
//   attention_v2/attention_cuda_kernel_v2.cu:7-91        step1 fwd/bwd
//   rpe_v2/relative_pos_encoding_cuda_kernel_v2.cu:247-340   dot_prod_with_idx v3 fwd/bwd
//   rpe_v2/relative_pos_encoding_cuda_kernel_v2.cu:397-484   step2 with rel-pos value v2 fwd/bwd
// How it is computed is new.  The reference launches one CTA per (query, head) with one thread per
// pair, re-reads index1 inside the channel loop, sums through shared/global float atomics and hits
// the rel-pos tables with 96 scalar global loads per pair-head.  Here:
//   * three generic warp-per-segment kernels cover all ten operators:
//       seg_dot     per pair-head scalar   = <x_row, y_row + Ex> + <y_row, Ey>
//       seg_reduce  per row vector         = sum over the row's pairs  w * (y_row + E)
//       table_grad  per table              = sum_n  X[n] (x) hist_n   (a small dense GEMM, see below)
//   * a head row (16/32 floats) is split over 4/8 lanes as float4, so a warp works on 8/4 pairs at a time
//     and every gathered 64/128 B row chunk is one fully used set of sectors;
//   * rel-pos tables are staged once per CTA in shared memory, re-laid out [axis][l][head][c] so the three
//     lookups per channel quad are three LDS.128 instead of twelve stride-3 scalar loads;
//   * grad_k / grad_v gather over the transposed CSR (pairs grouped by key) instead of scattering with
//     atomics, grad_q / grad_attn are segment-local register reductions;
//   * table gradients use  gT[l,h,c,a] = sum_n X[n,h,c] * W_a[n,l,h],  W_a[n,l,h] = sum_{m in seg(n), r[m,a]=l} w[m,h]:
//     a per-row histogram (3 scalar adds per pair-head instead of 3*d atomics) followed by a register-tiled
//     fp32 outer-product accumulation; one flush of red.global.add per CTA at the end.
#include <algorithm>
#include <type_traits>
#include <cstdlib>

#include <cub/cub.cuh>

#include "common.cuh"

namespace stb200 {

constexpr int kThreads = 256;
constexpr int kMaxSegThreads = 512;   // seg_dot / seg_reduce may run with 256 or 512 threads per CTA (same staged tables, more warps)
constexpr int kRowsPerChunk = 64;  // rows (queries or keys) a CTA takes per grid-stride step (upper bound, see rows_per_chunk)

struct SegParams {
    int N, h, L;
    const float *X;        // rows indexed by the segment owner n            [N, h, D]
    const float *Y;        // rows indexed by gather_idx                      [*, h, D]
    const float *w;        // per pair-head weights                           [M, h]
    const int *offsets;    // CSR offsets of the segment owner                [N+1]
    const int *gather_idx; // row id in Y per segment slot                    [M]
    const int *pair_id;    // pair id per segment slot (transposed CSR) or null
    const float *Tx, *Ty;  // rel-pos tables [L, h, D, 3]
    const int *rel_idx;    // [M, 3]
    const int *row_order;   // optional: process rows in this order (window-sorted => neighbouring warps share gathered rows in L1)
    const unsigned *packed; // optional: the three bins of each segment slot packed 10 bits each (replaces rel_idx)
    float *out;
    int accumulate;
    int w_by_slot;          // PERM kernels: w is already in segment-slot (transposed) order, index it by slot instead of pair id
    int rows_per_chunk;     // rows a CTA takes per grid-stride step: 64, fewer for small N so that every CTA gets >= 8 steps
};

// Copy one head group of a [L,h,D,3] table into shared memory as [axis][l][hh][copy][c].
// A 16-float row is 64 B = half of the 32 banks, and the lanes of one LDS.128 quarter-warp (two lane groups) read
// two different, data-dependent rows: with a single copy they collide whenever both rows fall into the same half
// (always, for an even head group in seg_reduce; half of the time otherwise).  For D = 16 the table is therefore
// stored twice, copy c in bank half c, and lane group s reads copy s & 1: every look-up is conflict-free.
// A 32-float row already spans all banks and needs one copy.  seg_dot keeps one copy as well: its lane groups walk
// (pair, head) items, so neighbouring groups mostly hit different halves already (1-14 % conflicts measured), and with
// two staged tables the second copy costs more in lost L1 capacity for the row gathers than the conflicts do.
template <int D>
constexpr int kReduceTableCopies = D == 16 ? 2 : 1;

template <int D, int HG, int NC>
__device__ __forceinline__ void stage_table(float *dst, const float *__restrict__ src, int L, int h, int h0) {
    const int total = 3 * L * HG * D;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int c = i % D;
        const int hh = (i / D) % HG;
        const int l = (i / (D * HG)) % L;
        const int a = i / (D * HG * L);
        const float v = __ldg(src + ((size_t)(l * h + h0 + hh) * D + c) * 3 + a);
#pragma unroll
        for (int cp = 0; cp < NC; ++cp) dst[((i / D) * NC + cp) * D + c] = v;
    }
}

// seg_dot with a three-head group: four 64 B slots per (axis, bin) = [h0, h1, h2, h2] (256 B, slot parity = bank half).
// The two lane groups of an LDS.128 quarter-warp hold consecutive items (pair, head): (h0,h1) and (h1,h2) of one pair
// sit in different halves by construction; the pair-crossing case (h2 of one pair, h0 of the next) takes the second
// copy of h2 in the other half - so no look-up ever has a bank conflict (ncu before: 15 % of the shared wavefronts).
template <int D>
__device__ __forceinline__ void stage_table_3in4(float *dst, const float *__restrict__ src, int L, int h, int h0) {
    const int total = 3 * L * 4 * D;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int c = i % D;
        const int slot = (i / D) % 4;
        const int l = (i / (D * 4)) % L;
        const int a = i / (D * 4 * L);
        dst[i] = __ldg(src + ((size_t)(l * h + h0 + min(slot, 2)) * D + c) * 3 + a);
    }
}

__device__ __forceinline__ int clampi(int v, int hi) { return min(max(v, 0), hi); }

constexpr int kUnroll = 4;  // independent row gathers kept in flight per lane

// the three rel-pos bins of one pair, clamped to [0, L) and packed 10 bits each (L <= 1024)
__device__ __forceinline__ unsigned pack_bins(const int *__restrict__ r, int L) {
    const unsigned r0 = clampi(ld_stream(r + 0), L - 1), r1 = clampi(ld_stream(r + 1), L - 1), r2 = clampi(ld_stream(r + 2), L - 1);
    return r0 | (r1 << 10) | (r2 << 20);
}

// E[c..c+3] = (T[0][r0] + T[1][r1]) + T[2][r2] for one head chunk (left-to-right adds like the reference)
// `gc` = copy * G + g: the lane's float4 column inside the (duplicated) row, see stage_table
template <int D, int HG, int NC>
__device__ __forceinline__ float4 table_sum4(const float *ts, int L, int r0, int r1, int r2, int hh, int gc) {
    const float4 *t4 = reinterpret_cast<const float4 *>(ts);
    constexpr int G = D / 4 * NC;
    const float4 a = t4[((0 * L + r0) * HG + hh) * G + gc];
    const float4 b = t4[((1 * L + r1) * HG + hh) * G + gc];
    const float4 c = t4[((2 * L + r2) * HG + hh) * G + gc];
    return f4_add(f4_add(a, b), c);
}

// ------------------------------------------------------------------------------------------------
// seg_dot: out[m, h] = XY ? <x,y> : 0  +  EX ? <x,Ex(m)> : 0  +  EY ? <y,Ey(m)> : 0
//   step1 fwd            XY           x=q[n]        y=k[i1]
//   dot_prod_with_idx v3 EX|EY        x=q[n]        y=k[i1]      Tx=table_q Ty=table_k
//   step2-rpv bwd gattn  XY|EX        x=grad_out[n] y=v[i1]      Tx=table_v
// A warp owns a query; its len*HG (pair, head) items are spread over 32/G lane groups of G=D/4 lanes.
// Table variants run 512 threads per CTA; capping them at 64 registers keeps two CTAs (32 warps) resident per SM,
// which hides the shared-memory look-up latency better than the 80-register / 16-warp build (measured: -20 %).
template <int D, int HG, bool XY, bool EX, bool EY, bool PAD = false>
__global__ void __launch_bounds__(kMaxSegThreads, (EX || EY) ? 2 : 1) seg_dot_kernel(const SegParams p) {
    extern __shared__ float4 smem4[];
    float *smem = reinterpret_cast<float *>(smem4);
    constexpr int G = D / 4, NS = kWarp / G;
    constexpr int RS = PAD ? 4 : HG;   // 64 B slots per (axis, bin): PAD = three heads in four slots, see stage_table_3in4
    static_assert(!PAD || (HG == 3 && D == 16), "the padded layout is for three 16-float heads");
    const int L = p.L, h = p.h, C = p.h * D;
    const int h0 = blockIdx.y * HG;
    const int tsz = 3 * L * RS * D;
    float *tx = smem;
    float *ty = tx + (EX ? tsz : 0);
    float *xs = ty + (EY ? tsz : 0);
    if (PAD) {
        if (EX) stage_table_3in4<D>(tx, p.Tx, L, h, h0);
        if (EY) stage_table_3in4<D>(ty, p.Ty, L, h, h0);
    } else {
        if (EX) stage_table<D, HG, 1>(tx, p.Tx, L, h, h0);
        if (EY) stage_table<D, HG, 1>(ty, p.Ty, L, h, h0);
    }
    if (EX || EY) __syncthreads();

    const int warp = threadIdx.x / kWarp, lane = threadIdx.x % kWarp, nwarps = blockDim.x / kWarp;
    const int grp = lane / G, g = lane % G;
    float4 *xw = reinterpret_cast<float4 *>(xs + warp * RS * D);

    for (int base_n = blockIdx.x * p.rows_per_chunk; base_n < p.N; base_n += gridDim.x * p.rows_per_chunk) {
        const int end_n = min(p.N, base_n + p.rows_per_chunk);
        for (int nn = base_n + warp; nn < end_n; nn += nwarps) {
            const int n = p.row_order ? __ldg(p.row_order + nn) : nn;
            const int start = __ldg(p.offsets + n), len = __ldg(p.offsets + n + 1) - start;
            if (len <= 0) continue;
            __syncwarp();
            for (int i = lane; i < RS * G; i += kWarp) xw[i] = ld_row4(p.X + (size_t)n * C + (h0 + min(i / G, HG - 1)) * D + 4 * (i % G));
            __syncwarp();
            // 32 pairs at a time: one coalesced load of the key ids (and the packed rel-pos bins) per chunk,
            // handed to the lane groups with shuffles; kUnroll independent row gathers in flight per lane.
            for (int c0 = 0; c0 < len; c0 += kWarp) {
                const int cnt = min(kWarp, len - c0);
                const int mt = start + c0 + min(lane, cnt - 1);
                const int j_l = ld_stream(p.gather_idx + mt);
                unsigned pk_l = 0;
                if (EX || EY) pk_l = p.packed ? __ldg(p.packed + mt) : pack_bins(p.rel_idx + 3 * (size_t)mt, L);
                const int items = cnt * HG;
                for (int e0 = 0; e0 < items; e0 += NS * kUnroll) {
                    float4 y4[kUnroll];
                    int pl[kUnroll], hh[kUnroll];
                    bool act[kUnroll];
#pragma unroll
                    for (int u = 0; u < kUnroll; ++u) {
                        const int e = e0 + u * NS + grp;
                        act[u] = e < items;
                        const int ee = act[u] ? e : items - 1;
                        pl[u] = ee / HG;
                        hh[u] = ee - pl[u] * HG;
                        const int j = __shfl_sync(0xffffffffu, j_l, pl[u]);
                        y4[u] = ld_row4(p.Y + (size_t)j * C + (h0 + hh[u]) * D + 4 * g);
                    }
#pragma unroll
                    for (int u = 0; u < kUnroll; ++u) {
                        if (e0 + u * NS >= items) break;   // warp-uniform
                        // PAD: head 2 has a copy in either bank half; an even lane group takes the upper one (its
                        // quarter-warp partner then holds head 0 of the next pair), an odd one the lower (partner: head 1)
                        const int sl = PAD && hh[u] == 2 ? 3 - (grp & 1) : hh[u];
                        const float4 x4 = xw[sl * G + g];
                        float acc = 0.f;
                        if (XY) acc = f4_dot(x4, y4[u], acc);
                        if (EX || EY) {
                            const unsigned pk = __shfl_sync(0xffffffffu, pk_l, pl[u]);
                            const int r0 = pk & 0x3ff, r1 = (pk >> 10) & 0x3ff, r2 = pk >> 20;
                            if (EX) acc = f4_dot(x4, table_sum4<D, RS, 1>(tx, L, r0, r1, r2, sl, g), acc);
                            if (EY) acc = f4_dot(y4[u], table_sum4<D, RS, 1>(ty, L, r0, r1, r2, sl, g), acc);
                        }
                        acc = group_sum<G>(acc);
                        if (act[u] && g == 0) p.out[(size_t)(start + c0 + pl[u]) * h + h0 + hh[u]] = acc;
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// seg_reduce: out[n, h, :] (+)= sum_{t in seg(n)} w[m(t), h] * ( HAS_Y ? Y[gather_idx[t], h, :] : 0  +  HAS_T ? E(m(t), h, :) : 0 )
//   step2-rpv fwd        HAS_Y|HAS_T   w=attn  Y=v  by index1            T=table_v
//   step1 bwd grad_q     HAS_Y         w=g     Y=k  by index1
//   rpe  bwd grad_q      HAS_T         w=g                                T=table_q
//   step1 bwd grad_k     HAS_Y  PERM   w=g     Y=q  by t_index0   (rows = keys, transposed CSR)
//   rpe  bwd grad_k      HAS_T  PERM   w=g                                T=table_k
//   step2 bwd grad_v     HAS_Y  PERM   w=attn  Y=grad_out by t_index0
// A warp owns a row; 32/G pair slots of G lanes each accumulate float4 per head, then xor-shuffle across slots.
template <int D, int HG, bool HAS_Y, bool HAS_T, bool PERM>
__global__ void __launch_bounds__(kMaxSegThreads) seg_reduce_kernel(const SegParams p) {
    extern __shared__ float4 smem4[];
    float *ts = reinterpret_cast<float *>(smem4);
    constexpr int G = D / 4, NS = kWarp / G;
    const int L = p.L, h = p.h, C = p.h * D;
    const int h0 = blockIdx.y * HG;
    if (HAS_T) {
        stage_table<D, HG, kReduceTableCopies<D>>(ts, p.Tx, L, h, h0);
        __syncthreads();
    }
    const int warp = threadIdx.x / kWarp, lane = threadIdx.x % kWarp, nwarps = blockDim.x / kWarp;
    const int slot = lane / G, g = lane % G;
    const int gc = (kReduceTableCopies<D> == 2 ? (slot & 1) * G : 0) + g;

    for (int base_n = blockIdx.x * p.rows_per_chunk; base_n < p.N; base_n += gridDim.x * p.rows_per_chunk) {
        const int end_n = min(p.N, base_n + p.rows_per_chunk);
        for (int nn = base_n + warp; nn < end_n; nn += nwarps) {
            const int n = p.row_order ? __ldg(p.row_order + nn) : nn;
            const int start = __ldg(p.offsets + n), end = __ldg(p.offsets + n + 1);
            float4 acc[HG];
#pragma unroll
            for (int hh = 0; hh < HG; ++hh) acc[hh] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int c0 = start; c0 < end; c0 += kWarp) {
                const int cnt = min(kWarp, end - c0);
                const int tl = c0 + min(lane, cnt - 1);
                const bool need_m = PERM && !(p.w_by_slot && (!HAS_T || p.packed));
                const int m_l = need_m ? ld_stream(p.pair_id + tl) : tl;
                const int j_l = HAS_Y ? ld_stream(p.gather_idx + tl) : 0;
                unsigned pk_l = 0;
                if (HAS_T) pk_l = p.packed ? __ldg(p.packed + tl) : pack_bins(p.rel_idx + 3 * (size_t)m_l, L);
                // A step takes UU pair slots per lane group (UU * HG row gathers in flight).  Full steps use two slots;
                // a tail of at most NS pairs takes a one-slot step, so a row wastes fewer than NS pair slots.
                auto step = [&](auto uu_c, int s0) {
                    constexpr int UU = decltype(uu_c)::value;
                    float4 val[UU][HG];
                    float wv[UU][HG];
#pragma unroll
                    for (int u = 0; u < UU; ++u) {
                        const int pl = s0 + u * NS + slot;
                        const bool act = pl < cnt;
                        const int pc = act ? pl : cnt - 1;
                        const int m = need_m ? __shfl_sync(0xffffffffu, m_l, pc) : c0 + pc;
                        const int mw = p.w_by_slot ? c0 + pc : m;
                        const int j = HAS_Y ? __shfl_sync(0xffffffffu, j_l, pc) : 0;
                        const unsigned pk = HAS_T ? __shfl_sync(0xffffffffu, pk_l, pc) : 0u;
                        const int r0 = pk & 0x3ff, r1 = (pk >> 10) & 0x3ff, r2 = pk >> 20;
#pragma unroll
                        for (int hh = 0; hh < HG; ++hh) {
                            wv[u][hh] = act ? ld_stream(p.w + (size_t)mw * h + h0 + hh) : 0.f;
                            val[u][hh] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (HAS_Y) val[u][hh] = ld_row4(p.Y + (size_t)j * C + (h0 + hh) * D + 4 * g);
                            if (HAS_T) val[u][hh] = f4_add(table_sum4<D, HG, kReduceTableCopies<D>>(ts, L, r0, r1, r2, hh, gc), val[u][hh]);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < UU; ++u)
#pragma unroll
                        for (int hh = 0; hh < HG; ++hh) acc[hh] = f4_fma(wv[u][hh], val[u][hh], acc[hh]);
                };
                int s0 = 0;
                for (; s0 + NS < cnt; s0 += 2 * NS) step(std::integral_constant<int, 2>{}, s0);
                if (s0 < cnt) step(std::integral_constant<int, 1>{}, s0);
            }
#pragma unroll
            for (int hh = 0; hh < HG; ++hh) {
#pragma unroll
                for (int o = G; o < kWarp; o <<= 1) {
                    acc[hh].x += __shfl_xor_sync(0xffffffffu, acc[hh].x, o);
                    acc[hh].y += __shfl_xor_sync(0xffffffffu, acc[hh].y, o);
                    acc[hh].z += __shfl_xor_sync(0xffffffffu, acc[hh].z, o);
                    acc[hh].w += __shfl_xor_sync(0xffffffffu, acc[hh].w, o);
                }
            }
            if (slot == 0) {
#pragma unroll
                for (int hh = 0; hh < HG; ++hh) {
                    float4 *dst = reinterpret_cast<float4 *>(p.out + (size_t)n * C + (h0 + hh) * D + 4 * g);
                    *dst = p.accumulate ? f4_add(*dst, acc[hh]) : acc[hh];
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// table_grad: gT[l, h, c, a] += sum_n X[n, h, c] * W_a[n, l],  W_a[n, l] = sum_{t in seg(n), rel_idx[m(t), a] = l} w[m(t), h]
//
// The reference issues 3*d float atomics per pair-head onto the 2*9216 table addresses.  Here the sum is split
// into a per-row histogram (3 scalar adds per pair-head, no atomics) and a dense product with the row matrix:
//   grid = (persistent tiles of TQ rows, head groups).  Per tile:
//     A0  stage pairs [c0, c0+32) of EVERY row of the tile (the rows come from a length-sorted order, so the slices
//         are equally full): packed bins + the weights of all heads of the group, one round trip to memory per
//         chunk with every thread issuing independent loads;
//     A1  one thread per (head, axis, row) adds its row's weights into its private column of
//         W[head][(axis, l)][row] in shared memory (no atomics);
//     B   per head: C[(axis,l), c] += W^T X on the tensor cores: mma.sync m16n8k8 TF32 with the 3-term split
//         (hi*hi + hi*lo + lo*hi), i.e. fp32-level accuracy; accumulators stay in registers across all tiles;
//   one red.global.add per accumulator element at the end of the CTA.
constexpr int kTQ = 32;        // rows per tile
constexpr int kTQP = kTQ + 4;  // pitch of W rows: conflict-free A-fragment loads
constexpr int kPR = 32;         // pairs of EACH row staged per chunk (a chunk = the same slice of all 32 rows)
constexpr int kPRP = kPR + 1;   // pitch of a row's slice: lanes = rows read conflict-free
constexpr int kPC = kTQ * kPRP; // staged elements per chunk
// one thread per (head, axis, row) in the histogram phase
__host__ __device__ constexpr int kTGThreads(int hgc) { return hgc * 3 * kTQ < 128 ? 128 : hgc * 3 * kTQ; }

// 3xTF32 operand split.  The tensor core reads only the top 19 bits of a tf32 operand, i.e. it truncates: so the
// high part is the raw fp32 word, and the low part is x - trunc(x) (exact in fp32), again passed raw.  Two
// instructions per element (cvt.rna.tf32 is a ~5-instruction emulation on sm_100a and dominated the first version).
__device__ __forceinline__ void split_tf32(float x, unsigned &hi, unsigned &lo) {
    hi = __float_as_uint(x);
    lo = __float_as_uint(x - __uint_as_float(hi & 0xffffe000u));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int D, int HGC, bool PERM, bool MULTI>
__global__ void __launch_bounds__(kTGThreads(HGC), (D == 16 ? 2 : 1)) table_grad_kernel(const SegParams p, int row_pass_base, int Rpad) {
    extern __shared__ float4 smem4[];
    constexpr int NT = D / 8;                               // n-tiles (8 channels each)
    constexpr int NW = kTGThreads(HGC) / kWarp;             // warps
    constexpr int UPW = (HGC * 16 + NW - 1) / NW;           // (head, m-tile) units per warp (<= 256 table rows per pass)
    constexpr int XP = D + 8;                               // pitch of the X tile: conflict-free B-fragment loads
    const int L = p.L, h = p.h, R = 3 * L;
    const int Rp = min(256, R - row_pass_base);             // table rows handled in this pass
    const int n_mt = (Rp + 15) / 16;
    const int h0 = blockIdx.y * HGC;
    float *W = reinterpret_cast<float *>(smem4);            // [HGC][Rpad][kTQP]
    float *Xs = W + HGC * Rpad * kTQP;                      // [HGC][kTQ][XP]
    float *sw = Xs + HGC * kTQ * XP;                        // [HGC][kTQ][kPRP] weights of the staged slice
    unsigned *pk = reinterpret_cast<unsigned *>(sw + HGC * kPC);   // [kTQ][kPRP] r0 | r1<<8 | r2<<16
    int *soff = reinterpret_cast<int *>(pk + kPC);          // [kTQ + 1] pair counts of the tile's rows, [kTQ] = their maximum
    int *gst = soff + kTQ + 8;                              // [kTQ] where each row's pairs start in the CSR
    int *rown = gst + kTQ;                                  // [kTQ] row ids (-1 beyond N)
    static_assert(kTQ == kWarp, "the tile prologue scans the row lengths in one warp");
    const int tid = threadIdx.x, lane = tid % kWarp, warp = tid / kWarp, nthr = blockDim.x;
    const int gid = lane >> 2, tig = lane & 3;

    float acc[UPW][NT][4];
#pragma unroll
    for (int a = 0; a < UPW; ++a)
#pragma unroll
        for (int b = 0; b < NT; ++b)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.f;

    // Row descriptors are fetched two tiles ahead (row id) / one tile ahead (segment start and length) by warp 0, so the
    // row_order -> offsets -> pairs chain of dependent global loads is off the critical path of a tile.
    constexpr int SE = (kTQ * kPR + kTGThreads(HGC) - 1) / kTGThreads(HGC);              // staged pairs per thread and chunk
    constexpr int XE = (HGC * kTQ * (D / 4) + kTGThreads(HGC) - 1) / kTGThreads(HGC);    // X row quarters per thread
    const int tile_stride = gridDim.x * kTQ;
    int n_cur = -1, gs_cur = 0, len_cur = 0, n_next = -1;   // warp 0, lane = row of the tile
    if (warp == 0) {
        const int r = blockIdx.x * kTQ + lane, r2 = r + tile_stride;
        if (r < p.N) {
            n_cur = p.row_order ? __ldg(p.row_order + r) : r;
            gs_cur = __ldg(p.offsets + n_cur);
            len_cur = __ldg(p.offsets + n_cur + 1) - gs_cur;
        }
        if (r2 < p.N) n_next = p.row_order ? __ldg(p.row_order + r2) : r2;
    }

    for (int base_n = blockIdx.x * kTQ; base_n < p.N; base_n += tile_stride) {
        __syncthreads();   // previous tile fully consumed
        // The tile's rows are row_order[base_n .. base_n+31] (or consecutive rows); their pair segments need not be
        // adjacent in the CSR: gst[] = where each row's segment starts, soff[] = its length, soff[kTQ] = the longest.
        // With rows sorted by length (len_order) every lane of the histogram phase runs the same trip count.
        if (warp == 0) {
            soff[lane] = len_cur;
            const int mx = __reduce_max_sync(0xffffffffu, len_cur);
            if (lane == 0) soff[kTQ] = mx;
            gst[lane] = gs_cur;
            rown[lane] = n_cur;
            // descriptors of the next tile (consumed one iteration later) and row ids of the one after
            n_cur = n_next;
            gs_cur = 0;
            len_cur = 0;
            if (n_cur >= 0) {
                gs_cur = __ldg(p.offsets + n_cur);
                len_cur = __ldg(p.offsets + n_cur + 1) - gs_cur;
            }
            const int r2 = base_n + 2 * tile_stride + lane;
            n_next = r2 < p.N ? (p.row_order ? __ldg(p.row_order + r2) : r2) : -1;
        }
        for (int i = tid; i < HGC * Rpad * kTQP / 4; i += nthr) smem4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncthreads();
        const int maxlen = soff[kTQ];
        // X rows of the tile: loads issued here, stored together with the first pair slice below
        float4 xv[XE];
#pragma unroll
        for (int e = 0; e < XE; ++e) {
            const int i = tid + e * nthr;
            xv[e] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < HGC * kTQ * (D / 4)) {
                const int hh = i / (kTQ * (D / 4)), t = (i / (D / 4)) % kTQ, c4 = i % (D / 4);
                const int n = rown[t];
                if (n >= 0) xv[e] = ld_row4(p.X + ((size_t)n * h + h0 + hh) * D + 4 * c4);
            }
        }
        // ---- histograms of all heads of the group; per chunk the pairs [c0, c0 + kPR) of every row are staged as
        // slice[row][kPRP] (128 B runs of one row -> coalesced loads; pitch 33 -> conflict-free reads by lanes = rows).
        // All global loads of a chunk are issued before the first shared-memory store (one memory round trip).
        for (int c0 = 0; c0 < maxlen || c0 == 0; c0 += kPR) {
            if (c0 > 0) __syncthreads();   // previous chunk consumed
            unsigned sq[SE];
            float swv[SE][HGC];
            int so[SE];
#pragma unroll
            for (int e = 0; e < SE; ++e) {
                const int i = tid + e * nthr;
                const int t = i / kPR, j = i - t * kPR;
                so[e] = -1;
                if (i < kTQ * kPR && c0 + j < soff[t]) {
                    const int gpos = gst[t] + c0 + j;
                    const int m = PERM && !(p.w_by_slot && p.packed) ? __ldg(p.pair_id + gpos) : gpos;
                    const int mw = p.w_by_slot ? gpos : m;
                    so[e] = t * kPRP + j;
                    if (p.packed) {
                        sq[e] = __ldg(p.packed + gpos);
                    } else {
                        const unsigned r0 = clampi(__ldg(p.rel_idx + 3 * (size_t)m + 0), L - 1);
                        const unsigned r1 = clampi(__ldg(p.rel_idx + 3 * (size_t)m + 1), L - 1);
                        const unsigned r2 = clampi(__ldg(p.rel_idx + 3 * (size_t)m + 2), L - 1);
                        sq[e] = r0 | (r1 << 10) | (r2 << 20);
                    }
#pragma unroll
                    for (int hh = 0; hh < HGC; ++hh) swv[e][hh] = __ldg(p.w + (size_t)mw * h + h0 + hh);
                }
            }
            if (c0 == 0) {
#pragma unroll
                for (int e = 0; e < XE; ++e) {
                    const int i = tid + e * nthr;
                    if (i < HGC * kTQ * (D / 4)) {
                        const int hh = i / (kTQ * (D / 4)), t = (i / (D / 4)) % kTQ, c4 = i % (D / 4);
                        *reinterpret_cast<float4 *>(Xs + (hh * kTQ + t) * XP + 4 * c4) = xv[e];
                    }
                }
            }
#pragma unroll
            for (int e = 0; e < SE; ++e) {
                if (so[e] < 0) continue;
                const unsigned q = sq[e];   // 10-bit fields -> 8-bit fields
                pk[so[e]] = (q & 0xffu) | (((q >> 10) & 0xffu) << 8) | (((q >> 20) & 0xffu) << 16);
#pragma unroll
                for (int hh = 0; hh < HGC; ++hh) sw[hh * kPC + so[e]] = swv[e][hh];
            }
            __syncthreads();
            for (int item = tid; item < HGC * 3 * kTQ; item += nthr) {   // one thread per (head, axis, row)
                const int t = item % kTQ, a = (item / kTQ) % 3, hh = item / (3 * kTQ);
                const int cnt = min(kPR, soff[t] - c0);
                float *col = W + hh * Rpad * kTQP + t + (a * L - row_pass_base) * kTQP;
                const float *wsrc = sw + hh * kPC + t * kPRP;
                const unsigned *psrc = pk + t * kPRP;
                const int sh = 8 * a;
                for (int i = 0; i < cnt; ++i) {
                    const int bin = (int)((psrc[i] >> sh) & 0xffu);
                    if (MULTI) {   // several passes over the table rows: skip bins outside this pass
                        const int row = a * L - row_pass_base + bin;
                        if (row < 0 || row >= Rp) continue;
                    }
                    col[bin * kTQP] += wsrc[i];
                }
            }
        }
        __syncthreads();
        // ---- C += W^T X  (rows of W^T = table rows, k = tile rows, n = channels).  A warp owns (head, m-tile) units:
        // the W fragment of a k-step is loaded and split once and used for every channel tile.
#pragma unroll
        for (int j = 0; j < UPW; ++j) {
            const int u = warp + j * NW;
            if (u < HGC * n_mt) {   // warp-uniform
                const int hh = u / n_mt, mt = u - hh * n_mt;
                const float *wa = W + (hh * Rpad + mt * 16 + gid) * kTQP + tig;
                const float *xb = Xs + (hh * kTQ + tig) * XP + gid;
#pragma unroll
                for (int k0 = 0; k0 < kTQ; k0 += 8) {
                    unsigned ah[4], al[4];
                    split_tf32(wa[k0], ah[0], al[0]);
                    split_tf32(wa[8 * kTQP + k0], ah[1], al[1]);
                    split_tf32(wa[k0 + 4], ah[2], al[2]);
                    split_tf32(wa[8 * kTQP + k0 + 4], ah[3], al[3]);
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) {
                        unsigned bh[2], bl[2];
                        split_tf32(xb[k0 * XP + nt * 8], bh[0], bl[0]);
                        split_tf32(xb[(k0 + 4) * XP + nt * 8], bh[1], bl[1]);
                        mma_tf32(acc[j][nt], al, bh);
                        mma_tf32(acc[j][nt], ah, bl);
                        mma_tf32(acc[j][nt], ah, bh);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < UPW; ++j) {
        const int u = warp + j * NW;
        if (u >= HGC * n_mt) continue;
        const int hh = u / n_mt, mt = u - hh * n_mt;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int row = row_pass_base + mt * 16 + gid + (c >= 2 ? 8 : 0);
                if (row >= R) continue;
                const int a = row / L, l = row - a * L;
                const int ch = nt * 8 + 2 * tig + (c & 1);
                atomicAdd(p.out + ((size_t)(l * h + h0 + hh) * D + ch) * 3 + a, acc[j][nt][c]);
            }
    }
}

// ------------------------------------------------------------------------------------------------
// Segment softmax (fwd: p = softmax_seg(a + b); bwd: gs = p * (gp - <p, gp>_seg)), one warp per query.
constexpr int kSoftCache = 4;  // values per lane kept in registers (segments up to 128 pairs in one pass)

__global__ void __launch_bounds__(kThreads) segment_softmax_fwd_kernel(int N, int h, const float *__restrict__ a,
                                                                       const float *__restrict__ b,
                                                                       const int *__restrict__ offsets,
                                                                       float *__restrict__ p) {
    const int lane = threadIdx.x % kWarp;
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) / kWarp, nw = gridDim.x * blockDim.x / kWarp;
    for (int n = wid; n < N; n += nw) {
        const int start = __ldg(offsets + n), len = __ldg(offsets + n + 1) - start;
        if (len <= 0) continue;
        for (int hh = 0; hh < h; ++hh) {
            const size_t base = (size_t)start * h + hh;
            float v[kSoftCache];
            float mx = -INFINITY;
#pragma unroll
            for (int u = 0; u < kSoftCache; ++u) {
                const int i = lane + u * kWarp;
                v[u] = -INFINITY;
                if (i < len) {
                    v[u] = a[base + (size_t)i * h];
                    if (b) v[u] += b[base + (size_t)i * h];
                }
                mx = fmaxf(mx, v[u]);
            }
            for (int i = lane + kSoftCache * kWarp; i < len; i += kWarp) {
                float s = a[base + (size_t)i * h];
                if (b) s += b[base + (size_t)i * h];
                mx = fmaxf(mx, s);
            }
            mx = warp_max(mx);
            float sum = 0.f;
#pragma unroll
            for (int u = 0; u < kSoftCache; ++u) {
                v[u] = (lane + u * kWarp < len) ? expf(v[u] - mx) : 0.f;
                sum += v[u];
            }
            for (int i = lane + kSoftCache * kWarp; i < len; i += kWarp) {
                float s = a[base + (size_t)i * h];
                if (b) s += b[base + (size_t)i * h];
                sum += expf(s - mx);
            }
            sum = group_sum<kWarp>(sum);
#pragma unroll
            for (int u = 0; u < kSoftCache; ++u) {
                const int i = lane + u * kWarp;
                if (i < len) p[base + (size_t)i * h] = v[u] / sum;
            }
            for (int i = lane + kSoftCache * kWarp; i < len; i += kWarp) {
                float s = a[base + (size_t)i * h];
                if (b) s += b[base + (size_t)i * h];
                p[base + (size_t)i * h] = expf(s - mx) / sum;
            }
        }
    }
}

__global__ void __launch_bounds__(kThreads) segment_softmax_bwd_kernel(int N, int h, const float *__restrict__ p,
                                                                       const float *__restrict__ gp,
                                                                       const int *__restrict__ offsets,
                                                                       float *__restrict__ gs) {
    const int lane = threadIdx.x % kWarp;
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) / kWarp, nw = gridDim.x * blockDim.x / kWarp;
    for (int n = wid; n < N; n += nw) {
        const int start = __ldg(offsets + n), len = __ldg(offsets + n + 1) - start;
        if (len <= 0) continue;
        for (int hh = 0; hh < h; ++hh) {
            const size_t base = (size_t)start * h + hh;
            float dot = 0.f;
            for (int i = lane; i < len; i += kWarp) dot = fmaf(p[base + (size_t)i * h], gp[base + (size_t)i * h], dot);
            dot = group_sum<kWarp>(dot);
            for (int i = lane; i < len; i += kWarp) {
                const size_t o = base + (size_t)i * h;
                gs[o] = p[o] * (gp[o] - dot);
            }
        }
    }
}

// Head-major variants for h <= 32: lane = (pair slot, head) with HP = next power of two >= h heads per slot, so
// the 32 lanes read 32/HP consecutive pairs x h heads = one contiguous span of the [M, h] array per step instead of
// a stride-h walk per head.  Values of the first kSoftCacheHP steps stay in registers.
constexpr int kSoftCacheHP = 8;

// kSoftRows rows per warp and step: their loads are issued together, which doubles the bytes in flight per warp (the
// kernels are latency-bound streams: a row is only ~100 floats and sits behind a dependent offsets load).
template <int HP, int R, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) segment_softmax_fwd_hp_kernel(int N, int h, const float *__restrict__ a,
                                                                          const float *__restrict__ b,
                                                                          const int *__restrict__ offsets,
                                                                          float *__restrict__ p,
                                                                          const int *__restrict__ rows = nullptr) {
    constexpr int SL = kWarp / HP;
    const int lane = threadIdx.x % kWarp, hd = lane % HP, slot = lane / HP;
    const bool on = hd < h;
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) / kWarp, nw = gridDim.x * blockDim.x / kWarp;
    for (int nn = wid; nn < N; nn += nw * R) {
        int len[R];
        size_t base[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int row = nn + r * nw;
            int start = 0;
            len[r] = 0;
            if (row < N) {
                const int n = rows ? __ldg(rows + row) : row;
                start = __ldg(offsets + n);
                len[r] = __ldg(offsets + n + 1) - start;
            }
            base[r] = (size_t)start * h + hd;
        }
        float v[R][kSoftCacheHP], mx[R], sum[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            mx[r] = -INFINITY;
#pragma unroll
            for (int u = 0; u < kSoftCacheHP; ++u) {
                const int i = slot + u * SL;
                v[r][u] = -INFINITY;
                if (on && i < len[r]) {
                    v[r][u] = a[base[r] + (size_t)i * h];
                    if (b) v[r][u] += b[base[r] + (size_t)i * h];
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
#pragma unroll
            for (int u = 0; u < kSoftCacheHP; ++u) mx[r] = fmaxf(mx[r], v[r][u]);
            for (int i = slot + kSoftCacheHP * SL; on && i < len[r]; i += SL) {
                float s = a[base[r] + (size_t)i * h];
                if (b) s += b[base[r] + (size_t)i * h];
                mx[r] = fmaxf(mx[r], s);
            }
#pragma unroll
            for (int o = HP; o < kWarp; o <<= 1) mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], o));
            sum[r] = 0.f;
#pragma unroll
            for (int u = 0; u < kSoftCacheHP; ++u) {
                v[r][u] = (on && slot + u * SL < len[r]) ? expf(v[r][u] - mx[r]) : 0.f;
                sum[r] += v[r][u];
            }
            for (int i = slot + kSoftCacheHP * SL; on && i < len[r]; i += SL) {
                float s = a[base[r] + (size_t)i * h];
                if (b) s += b[base[r] + (size_t)i * h];
                sum[r] += expf(s - mx[r]);
            }
#pragma unroll
            for (int o = HP; o < kWarp; o <<= 1) sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], o);
#pragma unroll
            for (int u = 0; u < kSoftCacheHP; ++u) {
                const int i = slot + u * SL;
                if (on && i < len[r]) p[base[r] + (size_t)i * h] = v[r][u] / sum[r];
            }
            for (int i = slot + kSoftCacheHP * SL; on && i < len[r]; i += SL) {
                float s = a[base[r] + (size_t)i * h];
                if (b) s += b[base[r] + (size_t)i * h];
                p[base[r] + (size_t)i * h] = expf(s - mx[r]) / sum[r];
            }
        }
    }
}

template <int HP, int R, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) segment_softmax_bwd_hp_kernel(int N, int h, const float *__restrict__ p,
                                                                          const float *__restrict__ gp,
                                                                          const int *__restrict__ offsets,
                                                                          float *__restrict__ gs) {
    constexpr int SL = kWarp / HP;
    const int lane = threadIdx.x % kWarp, hd = lane % HP, slot = lane / HP;
    const bool on = hd < h;
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) / kWarp, nw = gridDim.x * blockDim.x / kWarp;
    for (int nn = wid; nn < N; nn += nw * R) {
        int len[R];
        size_t base[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int n = nn + r * nw;
            int start = 0;
            len[r] = 0;
            if (n < N) {
                start = __ldg(offsets + n);
                len[r] = __ldg(offsets + n + 1) - start;
            }
            base[r] = (size_t)start * h + hd;
        }
        float pv[R][kSoftCacheHP], gv[R][kSoftCacheHP];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int u = 0; u < kSoftCacheHP; ++u) {
                const int i = slot + u * SL;
                pv[r][u] = gv[r][u] = 0.f;
                if (on && i < len[r]) {
                    pv[r][u] = p[base[r] + (size_t)i * h];
                    gv[r][u] = gp[base[r] + (size_t)i * h];
                }
            }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float dot = 0.f;
#pragma unroll
            for (int u = 0; u < kSoftCacheHP; ++u) dot = fmaf(pv[r][u], gv[r][u], dot);
            for (int i = slot + kSoftCacheHP * SL; on && i < len[r]; i += SL)
                dot = fmaf(p[base[r] + (size_t)i * h], gp[base[r] + (size_t)i * h], dot);
#pragma unroll
            for (int o = HP; o < kWarp; o <<= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
#pragma unroll
            for (int u = 0; u < kSoftCacheHP; ++u) {
                const int i = slot + u * SL;
                if (on && i < len[r]) gs[base[r] + (size_t)i * h] = pv[r][u] * (gv[r][u] - dot);
            }
            for (int i = slot + kSoftCacheHP * SL; on && i < len[r]; i += SL) {
                const size_t o = base[r] + (size_t)i * h;
                gs[o] = p[o] * (gp[o] - dot);
            }
        }
    }
}

// Span variants (default for h <= 32): the rows of a CSR chunk are ONE contiguous span of the [M, h] array, so a CTA copies the
// whole span of `rpc` consecutive rows into shared memory with 16-byte loads issued back to back (a few KB in flight per
// CTA instead of one ~100-float row per warp behind a dependent offsets load), works on the rows from shared memory with
// the head-major lane mapping above (same summation order, bit-identical results), and streams the span out again with
// 16-byte stores.  Chunks whose span exceeds the buffer (very long rows) take the per-row path from global memory.
#ifndef SPAN_EXP
#define SPAN_EXP expf
#endif
constexpr int kSpanFloats = 8192;   // per staged array: 32 KB (forward) / 2 x 32 KB (backward) per CTA

// copy [lo, hi) of src into sm (sm[0] <-> element lo - shift, shift = lo % 4), 16-byte accesses on the aligned interior
__device__ __forceinline__ void span_load(float *sm, const float *__restrict__ src, const float *__restrict__ add, size_t lo, size_t hi) {
    const size_t lo_al = lo & ~(size_t)3;
    const int nvec = (int)((hi - lo_al + 3) / 4);
    for (int vi = threadIdx.x; vi < nvec; vi += blockDim.x) {
        const size_t g = lo_al + 4 * (size_t)vi;
        if (g >= lo && g + 4 <= hi) {
            float4 x = __ldg(reinterpret_cast<const float4 *>(src + g));
            if (add) {
                const float4 y = __ldg(reinterpret_cast<const float4 *>(add + g));
                x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w;
            }
            *reinterpret_cast<float4 *>(sm + 4 * vi) = x;
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (g + e >= lo && g + e < hi) sm[4 * vi + e] = __ldg(src + g + e) + (add ? __ldg(add + g + e) : 0.f);
        }
    }
}
__device__ __forceinline__ void span_store(float *__restrict__ dst, const float *sm, size_t lo, size_t hi) {
    const size_t lo_al = lo & ~(size_t)3;
    const int nvec = (int)((hi - lo_al + 3) / 4);
    for (int vi = threadIdx.x; vi < nvec; vi += blockDim.x) {
        const size_t g = lo_al + 4 * (size_t)vi;
        if (g >= lo && g + 4 <= hi) {
            *reinterpret_cast<float4 *>(dst + g) = *reinterpret_cast<const float4 *>(sm + 4 * vi);
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (g + e >= lo && g + e < hi) dst[g + e] = sm[4 * vi + e];
        }
    }
}

// BWD = false: out = softmax_seg(x [+ y]);  BWD = true: x = p, y = grad_p, out = p * (grad_p - <p, grad_p>_seg)
template <int HP, bool BWD>
__global__ void __launch_bounds__(kThreads) segment_softmax_span_kernel(int N, int h, const float *__restrict__ x, const float *__restrict__ y,
                                                                        const int *__restrict__ offsets, float *__restrict__ out, int rpc) {
    extern __shared__ __align__(16) float span_smem[];
    float *sa = span_smem;                                    // [kSpanFloats + 4]
    float *sb = span_smem + (kSpanFloats + 4);                // backward only
    int *soff = reinterpret_cast<int *>(span_smem + (BWD ? 2 : 1) * (kSpanFloats + 4));   // [rpc + 1]
    constexpr int SL = kWarp / HP;
    const int lane = threadIdx.x % kWarp, warp = threadIdx.x / kWarp, nwarps = blockDim.x / kWarp;
    const int hd = lane % HP, slot = lane / HP;
    const bool on = hd < h;
    const int n_chunks = (N + rpc - 1) / rpc;
    for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const int r0 = c * rpc, nr = min(rpc, N - r0);
        for (int i = threadIdx.x; i <= nr; i += blockDim.x) soff[i] = __ldg(offsets + r0 + i);
        __syncthreads();
        const int s0 = soff[0], s1 = soff[nr];
        const size_t lo = (size_t)s0 * h, hi = (size_t)s1 * h;
        const int shift = (int)(lo & 3);
        const bool staged = hi - lo + 3 <= (size_t)kSpanFloats;
        if (staged) {
            span_load(sa, x, BWD ? nullptr : y, lo, hi);
            if (BWD) span_load(sb, y, nullptr, lo, hi);
            __syncthreads();
        }
        for (int rr = warp; rr < nr; rr += nwarps) {
            const int start = soff[rr], len = soff[rr + 1] - start;
            if (len <= 0) continue;
            if (staged) {
                float *ra = sa + (size_t)(start - s0) * h + shift + hd;
                if (!BWD) {
                    float mx = -INFINITY;
                    for (int i = slot; on && i < len; i += SL) mx = fmaxf(mx, ra[i * h]);
#pragma unroll
                    for (int o = HP; o < kWarp; o <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                    float sum = 0.f;
#pragma unroll 2
                    for (int i = slot; on && i < len; i += SL) {
                        const float e = SPAN_EXP(ra[i * h] - mx);
                        ra[i * h] = e;
                        sum += e;
                    }
#pragma unroll
                    for (int o = HP; o < kWarp; o <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
                    const float inv = 1.0f / sum;   // one division per (row, head); the products differ from e / sum by <= 1 ulp
                    for (int i = slot; on && i < len; i += SL) ra[i * h] *= inv;
                } else {
                    const float *rb = sb + (size_t)(start - s0) * h + shift + hd;
                    float dot = 0.f;
                    for (int i = slot; on && i < len; i += SL) dot = fmaf(ra[i * h], rb[i * h], dot);
#pragma unroll
                    for (int o = HP; o < kWarp; o <<= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
                    for (int i = slot; on && i < len; i += SL) ra[i * h] = ra[i * h] * (rb[i * h] - dot);
                }
            } else {   // oversized chunk: the row straight from global memory
                const size_t base = (size_t)start * h + hd;
                if (!BWD) {
                    float mx = -INFINITY;
                    for (int i = slot; on && i < len; i += SL) mx = fmaxf(mx, x[base + (size_t)i * h] + (y ? y[base + (size_t)i * h] : 0.f));
#pragma unroll
                    for (int o = HP; o < kWarp; o <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                    float sum = 0.f;
                    for (int i = slot; on && i < len; i += SL) sum += expf(x[base + (size_t)i * h] + (y ? y[base + (size_t)i * h] : 0.f) - mx);
#pragma unroll
                    for (int o = HP; o < kWarp; o <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
                    for (int i = slot; on && i < len; i += SL)
                        out[base + (size_t)i * h] = expf(x[base + (size_t)i * h] + (y ? y[base + (size_t)i * h] : 0.f) - mx) / sum;
                } else {
                    float dot = 0.f;
                    for (int i = slot; on && i < len; i += SL) dot = fmaf(x[base + (size_t)i * h], y[base + (size_t)i * h], dot);
#pragma unroll
                    for (int o = HP; o < kWarp; o <<= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
                    for (int i = slot; on && i < len; i += SL) {
                        const size_t o2 = base + (size_t)i * h;
                        out[o2] = x[o2] * (y[o2] - dot);
                    }
                }
            }
        }
        __syncthreads();
        if (staged) {
            span_store(out, sa, lo, hi);
            __syncthreads();
        }
    }
}

// dst[t, :] = src[perm[t], :]  (rows of h floats): per-pair weights into transposed-CSR order
__global__ void permute_rows_kernel(int M, int h, const float *__restrict__ src, const int *__restrict__ perm,
                                    float *__restrict__ dst) {
    const long long total = (long long)M * h;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(i / h), c = (int)(i - (long long)t * h);
        dst[i] = ld_stream(src + (size_t)__ldg(perm + t) * h + c);
    }
}

// ------------------------------------------------------------------------------------------------
// host-side launch helpers
// threads per CTA of the segment kernels: tuning knobs STB200_SEG_THREADS (both) / STB200_SEGRED_THREADS (seg_reduce
// only), multiples of 32 up to 512; default by table use
static int seg_threads(int ntables, bool reduce = false) {
    static const int env = getenv("STB200_SEG_THREADS") ? atoi(getenv("STB200_SEG_THREADS")) : 0;
    static const int env_red = getenv("STB200_SEGRED_THREADS") ? atoi(getenv("STB200_SEGRED_THREADS")) : 0;
    const int e = reduce && env_red ? env_red : env;
    if (e >= 64 && e <= kMaxSegThreads && e % kWarp == 0) return e;
    // seg_dot with tables: shared-memory / latency bound, 2 x 512 threads per SM share the staged tables best;
    // seg_reduce (about 100 registers per thread) does slightly better with two 256-thread CTAs, which also still fit
    // beside a resident FPS CTA of the geometry stream
    return ntables > 0 && !reduce ? 512 : kThreads;
}

// softmax launch variants: rows per warp step (R) x resident CTAs per SM (MINB); tuning knob STB200_SOFTMAX_VARIANT
static int softmax_variant() {
    static const int v = getenv("STB200_SOFTMAX_VARIANT") ? atoi(getenv("STB200_SOFTMAX_VARIANT")) : 0;
    return v;
}
template <int HP>
static void launch_softmax_fwd_hp1(int blocks, cudaStream_t s, int N, int h, const float *a, const float *b, const int *off,
                                   float *p, const int *rows) {
    switch (softmax_variant()) {
        case 1: segment_softmax_fwd_hp_kernel<HP, 2, 4><<<blocks, kThreads, 0, s>>>(N, h, a, b, off, p, rows); break;
        case 2: segment_softmax_fwd_hp_kernel<HP, 2, 6><<<blocks, kThreads, 0, s>>>(N, h, a, b, off, p, rows); break;
        case 3: segment_softmax_fwd_hp_kernel<HP, 1, 8><<<blocks, kThreads, 0, s>>>(N, h, a, b, off, p, rows); break;
        default: segment_softmax_fwd_hp_kernel<HP, 1, 1><<<blocks, kThreads, 0, s>>>(N, h, a, b, off, p, rows); break;   // 4 (and the rows variant)
    }
}
static void launch_softmax_fwd_hp(int blocks, cudaStream_t s, int N, int h, const float *a, const float *b, const int *off,
                                  float *p, const int *rows = nullptr) {
    if (h <= 1) launch_softmax_fwd_hp1<1>(blocks, s, N, h, a, b, off, p, rows);
    else if (h <= 2) launch_softmax_fwd_hp1<2>(blocks, s, N, h, a, b, off, p, rows);
    else if (h <= 4) launch_softmax_fwd_hp1<4>(blocks, s, N, h, a, b, off, p, rows);
    else if (h <= 8) launch_softmax_fwd_hp1<8>(blocks, s, N, h, a, b, off, p, rows);
    else if (h <= 16) launch_softmax_fwd_hp1<16>(blocks, s, N, h, a, b, off, p, rows);
    else launch_softmax_fwd_hp1<32>(blocks, s, N, h, a, b, off, p, rows);
}
template <int HP>
static void launch_softmax_bwd_hp1(int blocks, cudaStream_t s, int N, int h, const float *p, const float *gp, const int *off,
                                   float *gs) {
    switch (softmax_variant()) {
        case 1: segment_softmax_bwd_hp_kernel<HP, 2, 4><<<blocks, kThreads, 0, s>>>(N, h, p, gp, off, gs); break;
        case 2: segment_softmax_bwd_hp_kernel<HP, 2, 6><<<blocks, kThreads, 0, s>>>(N, h, p, gp, off, gs); break;
        case 3: segment_softmax_bwd_hp_kernel<HP, 1, 8><<<blocks, kThreads, 0, s>>>(N, h, p, gp, off, gs); break;
        default: segment_softmax_bwd_hp_kernel<HP, 1, 1><<<blocks, kThreads, 0, s>>>(N, h, p, gp, off, gs); break;
    }
}
static void launch_softmax_bwd_hp(int blocks, cudaStream_t s, int N, int h, const float *p, const float *gp, const int *off,
                                  float *gs) {
    if (h <= 1) launch_softmax_bwd_hp1<1>(blocks, s, N, h, p, gp, off, gs);
    else if (h <= 2) launch_softmax_bwd_hp1<2>(blocks, s, N, h, p, gp, off, gs);
    else if (h <= 4) launch_softmax_bwd_hp1<4>(blocks, s, N, h, p, gp, off, gs);
    else if (h <= 8) launch_softmax_bwd_hp1<8>(blocks, s, N, h, p, gp, off, gs);
    else if (h <= 16) launch_softmax_bwd_hp1<16>(blocks, s, N, h, p, gp, off, gs);
    else launch_softmax_bwd_hp1<32>(blocks, s, N, h, p, gp, off, gs);
}

// CTAs per launch = resident capacity x "waves".  More than one wave lets the hardware block scheduler rebalance when
// another stream (the geometry prefetch) holds part of the machine: with exactly one wave, every CTA that does not fit
// at launch runs after the others finish and doubles the kernel time.  Tuning knob: STB200_SEG_WAVES.
static int seg_waves() {
    static const int env = getenv("STB200_SEG_WAVES") ? atoi(getenv("STB200_SEG_WAVES")) : 1;
    return max(1, env);
}

// Grid and row-chunk size of a segment kernel.  A CTA walks chunks of consecutive rows (window-sorted rows share their
// gathered k/v rows in L1); with few rows (deep layers) 64-row chunks would leave CTAs with 3 vs 4 chunks, i.e. a 25 %
// tail, so the chunk shrinks (down to 16 rows) until every CTA has at least 8 of them.
static int grid_rows(int N, size_t smem, int groups, int *rows_per_chunk) {
    const int ctas_per_sm = (int)max((size_t)1, min((size_t)4, (size_t)(220 * 1024) / max(smem, (size_t)1)));
    const int ctas = max(1, (kNumSMs * ctas_per_sm * seg_waves() + groups - 1) / groups);
    static const int env_rpc = getenv("STB200_SEG_ROWS_PER_CHUNK") ? atoi(getenv("STB200_SEG_ROWS_PER_CHUNK")) : 0;
    int rpc = env_rpc > 0 ? env_rpc : max(16, min(kRowsPerChunk, N / (ctas * 8)));
    *rows_per_chunk = rpc;
    const int chunks = (N + rpc - 1) / rpc;
    return max(1, min(chunks, ctas));
}

template <typename K>
static int prep_smem(K kernel, size_t bytes) {
    if (bytes > 227 * 1024) {
        set_error("shared memory request %zu B exceeds 227 KB (table too large for one head group)", bytes);
        return STB200_ERR_ARG;
    }
    if (bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return STB200_ERR_CUDA;
        }
    }
    return STB200_OK;
}

// span kernels: rows per chunk so that a chunk's span is about half the staging buffer on average
template <bool BWD>
static int launch_softmax_span(cudaStream_t s, int N, long long M, int h, const float *x, const float *y, const int *off, float *out) {
    if (((uintptr_t)x | (uintptr_t)(y ? y : x) | (uintptr_t)out) & 15) return -1;   // unaligned views: per-row kernels
    const double per_row = (double)M * h / max(1, N);
    const int rpc = (int)max(2.0, min(64.0, (kSpanFloats / 2) / max(1.0, per_row)));
    const size_t smem = (size_t)(BWD ? 2 : 1) * (kSpanFloats + 4) * sizeof(float) + (size_t)(rpc + 1) * sizeof(int);
    const int n_chunks = (N + rpc - 1) / rpc;
    const int per_sm = (int)min((size_t)8, (size_t)(200 * 1024) / smem);
    const int blocks = max(1, min(n_chunks, kNumSMs * per_sm));
#define STB200_SPAN(HP)                                                                                                         \
    do {                                                                                                                        \
        static bool attr = false;                                                                                               \
        if (!attr) {                                                                                                            \
            if (int rc = prep_smem(segment_softmax_span_kernel<HP, BWD>, (size_t)2 * (kSpanFloats + 4) * sizeof(float) + 65 * sizeof(int))) return rc; \
            attr = true;                                                                                                        \
        }                                                                                                                       \
        segment_softmax_span_kernel<HP, BWD><<<blocks, kThreads, smem, s>>>(N, h, x, y, off, out, rpc);                         \
    } while (0)
    if (h <= 1) STB200_SPAN(1);
    else if (h <= 2) STB200_SPAN(2);
    else if (h <= 4) STB200_SPAN(4);
    else if (h <= 8) STB200_SPAN(8);
    else if (h <= 16) STB200_SPAN(16);
    else STB200_SPAN(32);
#undef STB200_SPAN
    return STB200_OK;
}


// heads per CTA: as many as divide h, up to 4, while the staged tables stay under ~100 KB
static int pick_hg(int h, int D, int L, int ntables) {
    int cap = 4;
    const int copies = ntables == 1 && D == 16 ? 2 : 1;   // upper bound: kReduceTableCopies
    while (cap > 1 && (size_t)ntables * 3 * L * cap * D * copies * 4 > 100 * 1024) --cap;
    return largest_head_group(h, cap);
}

// algorithmic bytes of one launch: every input read once, every output written once
static double seg_bytes(const SegParams &p, int D, int M, bool rows_x, bool rows_y, bool weights, bool gather, bool perm,
                        int ntables, bool out_pairs) {
    const double C = (double)p.h * D, N = p.N;
    double b = 4.0 * (N + 1);                                // offsets
    if (rows_x) b += 4.0 * N * C;
    if (rows_y) b += 4.0 * N * C;
    if (weights) b += 4.0 * M * p.h;
    if (gather) b += 4.0 * M;
    if (perm) b += 4.0 * M;
    if (ntables) b += (p.packed ? 4.0 : 12.0) * M + ntables * 12.0 * p.L * C;   // packed bins (4 B / pair) or rel_idx (12 B / pair) + tables
    b += out_pairs ? 4.0 * M * p.h : 4.0 * N * C * (p.accumulate ? 2 : 1);
    return b;
}

template <int D, int HG, bool XY, bool EX, bool EY>
static int launch_seg_dot_hg(const SegParams &p, int M, const char *name, cudaStream_t s) {
    const int threads = seg_threads(EX + EY);
    // three-head groups: conflict-free padded table layout when two CTAs with it still fit an SM
    static const bool pad_env = !(getenv("STB200_SEGDOT_NOPAD") && atoi(getenv("STB200_SEGDOT_NOPAD")));
    constexpr bool kCanPad = HG == 3 && D == 16 && (EX || EY);
    const size_t smem_pad = ((size_t)(EX + EY) * 3 * p.L * 4 * D + (threads / kWarp) * 4 * D) * sizeof(float);
    const bool pad = kCanPad && pad_env && 2 * (smem_pad + 1024) <= 227 * 1024;
    const size_t smem = pad ? smem_pad : ((size_t)(EX + EY) * 3 * p.L * HG * D + (threads / kWarp) * HG * D) * sizeof(float);
    void (*kern)(const SegParams) = seg_dot_kernel<D, HG, XY, EX, EY, false>;
    if constexpr (kCanPad) {
        if (pad) kern = seg_dot_kernel<D, HG, XY, EX, EY, true>;
    }
    if (int rc = prep_smem(kern, smem)) return rc;
    SegParams pl = p;
    dim3 grid(grid_rows(p.N, smem, p.h / HG, &pl.rows_per_chunk), p.h / HG);
    {
        KernelScope ks(name, seg_bytes(p, D, M, true, true, false, true, false, EX + EY, true), s);
        kern<<<grid, threads, smem, s>>>(pl);
    }
    return check_launch(name);
}

template <int D, bool XY, bool EX, bool EY>
static int launch_seg_dot_d(const SegParams &p, int M, const char *name, cudaStream_t s) {
    switch (pick_hg(p.h, D, p.L, EX + EY)) {
        case 4: return launch_seg_dot_hg<D, 4, XY, EX, EY>(p, M, name, s);
        case 3: return launch_seg_dot_hg<D, 3, XY, EX, EY>(p, M, name, s);
        case 2: return launch_seg_dot_hg<D, 2, XY, EX, EY>(p, M, name, s);
        default: return launch_seg_dot_hg<D, 1, XY, EX, EY>(p, M, name, s);
    }
}

template <bool XY, bool EX, bool EY>
static int launch_seg_dot(int D, const SegParams &p, int M, const char *name, cudaStream_t s) {
    if (p.N == 0) return STB200_OK;
    if (D == 16) return launch_seg_dot_d<16, XY, EX, EY>(p, M, name, s);
    return launch_seg_dot_d<32, XY, EX, EY>(p, M, name, s);
}

template <int D, int HG, bool HAS_Y, bool HAS_T, bool PERM>
static int launch_seg_reduce_hg(const SegParams &p, int M, const char *name, cudaStream_t s) {
    const size_t smem = (size_t)HAS_T * 3 * p.L * HG * D * kReduceTableCopies<D> * sizeof(float);
    auto kern = seg_reduce_kernel<D, HG, HAS_Y, HAS_T, PERM>;
    if (int rc = prep_smem(kern, smem)) return rc;
    SegParams pl = p;
    dim3 grid(grid_rows(p.N, smem, p.h / HG, &pl.rows_per_chunk), p.h / HG);
    {
        KernelScope ks(name, seg_bytes(p, D, M, false, HAS_Y, true, HAS_Y, PERM, HAS_T, false), s);
        kern<<<grid, seg_threads(HAS_T && !PERM, true), smem, s>>>(pl);
    }
    return check_launch(name);
}

template <int D, bool HAS_Y, bool HAS_T, bool PERM>
static int launch_seg_reduce_d(const SegParams &p, int M, const char *name, cudaStream_t s) {
    switch (pick_hg(p.h, D, p.L, HAS_T)) {
        case 4: return launch_seg_reduce_hg<D, 4, HAS_Y, HAS_T, PERM>(p, M, name, s);
        case 3: return launch_seg_reduce_hg<D, 3, HAS_Y, HAS_T, PERM>(p, M, name, s);
        case 2: return launch_seg_reduce_hg<D, 2, HAS_Y, HAS_T, PERM>(p, M, name, s);
        default: return launch_seg_reduce_hg<D, 1, HAS_Y, HAS_T, PERM>(p, M, name, s);
    }
}

template <bool HAS_Y, bool HAS_T, bool PERM>
static int launch_seg_reduce(int D, const SegParams &p, int M, const char *name, cudaStream_t s) {
    if (p.N == 0) return STB200_OK;
    if (D == 16) return launch_seg_reduce_d<16, HAS_Y, HAS_T, PERM>(p, M, name, s);
    return launch_seg_reduce_d<32, HAS_Y, HAS_T, PERM>(p, M, name, s);
}

template <int D, int HGC, bool PERM>
static int launch_table_grad_hg(const SegParams &p, int M, const char *name, cudaStream_t s) {
    const int R = 3 * p.L;
    const int Rpad = min(256, (R + 15) / 16 * 16);
    const size_t smem = ((size_t)HGC * Rpad * kTQP + HGC * kTQ * (D + 8) + HGC * kPC + kPC + 3 * kTQ + 8) * sizeof(float);
    void (*kern)(const SegParams, int, int) =
        R <= 256 ? table_grad_kernel<D, HGC, PERM, false> : table_grad_kernel<D, HGC, PERM, true>;
    if (int rc = prep_smem(kern, smem)) return rc;
    const int tiles = (p.N + kTQ - 1) / kTQ;
    const int groups = p.h / HGC;
    const int ctas_per_sm = max(1, min(4, (int)((226 * 1024) / (smem + 1024))));
    static const int tg_waves = max(1, getenv("STB200_TG_WAVES") ? atoi(getenv("STB200_TG_WAVES")) : 1);
    const int gx = max(1, min(tiles, (kNumSMs * ctas_per_sm * tg_waves + groups - 1) / groups));
    for (int pass = 0; pass < R; pass += 256) {
        KernelScope ks(name, seg_bytes(p, D, M, true, false, true, false, PERM, 1, false) - 4.0 * p.N * p.h * D, s);
        kern<<<dim3(gx, groups), kTGThreads(HGC), smem, s>>>(p, pass, Rpad);
    }
    return check_launch(name);
}

template <int D, bool PERM>
static int launch_table_grad_d(const SegParams &p, int M, const char *name, cudaStream_t s) {
    STB200_REQUIRE(p.L <= 256, STB200_ERR_ARG, "table length %d > 256 not supported by the table-gradient kernel", p.L);
    constexpr int cap = D == 16 ? 3 : 2;   // heads per CTA: shared-memory histograms and 2*(D/8)*4 accumulators per head
    static const int env_cap = getenv("STB200_TG_HEADS") ? atoi(getenv("STB200_TG_HEADS")) : cap;   // tuning knob
    switch (largest_head_group(p.h, max(1, min(cap, env_cap)))) {
        case 3: return launch_table_grad_hg<D, (cap >= 3 ? 3 : 1), PERM>(p, M, name, s);
        case 2: return launch_table_grad_hg<D, 2, PERM>(p, M, name, s);
        default: return launch_table_grad_hg<D, 1, PERM>(p, M, name, s);
    }
}

template <bool PERM>
static int launch_table_grad(int D, const SegParams &p, int M, const char *name, cudaStream_t s) {
    if (p.N == 0) return STB200_OK;
    if (D == 16) return launch_table_grad_d<16, PERM>(p, M, name, s);
    return launch_table_grad_d<32, PERM>(p, M, name, s);
}

static int check_dims(int N, int M, int h, int D) {
    STB200_REQUIRE(N >= 0 && M >= 0 && h > 0, STB200_ERR_ARG, "bad sizes N=%d M=%d h=%d", N, M, h);
    STB200_REQUIRE(D == 16 || D == 32, STB200_ERR_HEAD_DIM, "d != 16 and d != 32 (got %d)", D);
    return STB200_OK;
}

}  // namespace stb200

using namespace stb200;

extern "C" {

int stb200_attention_step1_forward_v2(int N, int M, int h, int C, unsigned int, const float *q, const float *k,
                                      const int *index0_offsets, const int *index1, float *attn, void *stream) {
    STB200_REQUIRE(h > 0 && C % h == 0, STB200_ERR_ARG, "C=%d not divisible by h=%d", C, h);
    if (int rc = check_dims(N, M, h, C / h)) return rc;
    if (M == 0) return STB200_OK;
    STB200_REQUIRE(q && k && index0_offsets && index1 && attn, STB200_ERR_ARG, "null pointer");
    SegParams p{};
    p.N = N; p.h = h; p.L = 0; p.X = q; p.Y = k; p.offsets = index0_offsets; p.gather_idx = index1; p.out = attn;
    return launch_seg_dot<true, false, false>(C / h, p, M, "seg_dot[step1_fwd]", (cudaStream_t)stream);
}

int stb200_attention_step1_backward_v2(int N, int M, int h, int C, unsigned int, const float *grad_out,
                                       const int *index0_offsets, const int *index1, const float *q, const float *k,
                                       float *grad_q, float *grad_k, const int *t_offsets, const int *t_pair,
                                       const int *t_index0, void *stream) {
    STB200_REQUIRE(h > 0 && C % h == 0, STB200_ERR_ARG, "C=%d not divisible by h=%d", C, h);
    if (int rc = check_dims(N, M, h, C / h)) return rc;
    STB200_REQUIRE(grad_out && index0_offsets && index1 && q && k && grad_q && grad_k, STB200_ERR_ARG, "null pointer");
    STB200_REQUIRE(t_offsets && t_pair && t_index0, STB200_ERR_ARG, "transposed CSR required (stb200_transpose_csr)");
    cudaStream_t s = (cudaStream_t)stream;
    SegParams p{};
    p.N = N; p.h = h; p.w = grad_out; p.Y = k; p.offsets = index0_offsets; p.gather_idx = index1; p.out = grad_q;
    if (int rc = launch_seg_reduce<true, false, false>(C / h, p, M, "seg_reduce[step1_bwd_gq]", s)) return rc;   // grad_q: overwritten (ref :90)
    p.Y = q; p.offsets = t_offsets; p.gather_idx = t_index0; p.pair_id = t_pair; p.out = grad_k; p.accumulate = 1;
    return launch_seg_reduce<true, false, true>(C / h, p, M, "seg_reduce_t[step1_bwd_gk]", s);  // grad_k: accumulated (ref :84)
}

int stb200_dot_prod_with_idx_forward_v3(int N, int M, int h, int hdim, int, int L, const float *q,
                                        const int *index_q_offsets, const float *k, const int *index_k,
                                        const float *table_q, const float *table_k, const int *rel_idx,
                                        float *output, void *stream) {
    if (int rc = check_dims(N, M, h, hdim)) return rc;
    if (M == 0) return STB200_OK;
    STB200_REQUIRE(L > 0 && q && k && index_q_offsets && index_k && table_q && table_k && rel_idx && output,
                   STB200_ERR_ARG, "null pointer or L<=0");
    SegParams p{};
    p.N = N; p.h = h; p.L = L; p.X = q; p.Y = k; p.offsets = index_q_offsets; p.gather_idx = index_k;
    p.Tx = table_q; p.Ty = table_k; p.rel_idx = rel_idx; p.out = output;
    return launch_seg_dot<false, true, true>(hdim, p, M, "seg_dot[rpe_fwd]", (cudaStream_t)stream);
}

int stb200_dot_prod_with_idx_backward_v3(int N, int M, int h, int hdim, int, int L, const float *grad_out,
                                         const float *q, const int *index_q_offsets, const float *k,
                                         const int *index_k, const float *table_q, const float *table_k,
                                         const int *rel_idx, float *grad_q, float *grad_k, float *grad_table_q,
                                         float *grad_table_k, const int *t_offsets, const int *t_pair,
                                         const int *t_index0, void *stream) {
    if (int rc = check_dims(N, M, h, hdim)) return rc;
    STB200_REQUIRE(L > 0 && grad_out && q && k && index_q_offsets && index_k && table_q && table_k && rel_idx &&
                       grad_q && grad_k && grad_table_q && grad_table_k, STB200_ERR_ARG, "null pointer or L<=0");
    STB200_REQUIRE(t_offsets && t_pair && t_index0, STB200_ERR_ARG, "transposed CSR required (stb200_transpose_csr)");
    cudaStream_t s = (cudaStream_t)stream;
    SegParams p{};
    p.N = N; p.h = h; p.L = L; p.w = grad_out; p.rel_idx = rel_idx;
    p.offsets = index_q_offsets; p.Tx = table_q; p.out = grad_q;
    if (int rc = launch_seg_reduce<false, true, false>(hdim, p, M, "seg_reduce[rpe_bwd_gq]", s)) return rc;    // grad_q = sum g*Eq (overwritten, ref :338)
    p.X = q; p.out = grad_table_q;
    if (int rc = launch_table_grad<false>(hdim, p, M, "table_grad[rpe_bwd_gtq]", s)) return rc;  // grad_table_q (+=)
    p.offsets = t_offsets; p.pair_id = t_pair; p.gather_idx = t_index0;
    p.Tx = table_k; p.out = grad_k; p.accumulate = 1;
    if (int rc = launch_seg_reduce<false, true, true>(hdim, p, M, "seg_reduce_t[rpe_bwd_gk]", s)) return rc;     // grad_k += sum g*Ek
    p.X = k; p.out = grad_table_k;
    return launch_table_grad<true>(hdim, p, M, "table_grad_t[rpe_bwd_gtk]", s);    // grad_table_k (+=)
}

int stb200_attention_step2_with_rel_pos_value_forward_v2(int N, int M, int h, int hdim, int, int L, const float *attn,
                                                         const float *v, const int *index0_offsets, const int *index1,
                                                         const float *table, const int *rel_idx, float *output,
                                                         void *stream) {
    if (int rc = check_dims(N, M, h, hdim)) return rc;
    STB200_REQUIRE(L > 0 && index0_offsets && output && (M == 0 || (attn && v && index1 && table && rel_idx)),
                   STB200_ERR_ARG, "null pointer or L<=0");
    SegParams p{};
    p.N = N; p.h = h; p.L = L; p.w = attn; p.Y = v; p.offsets = index0_offsets; p.gather_idx = index1;
    p.Tx = table; p.rel_idx = rel_idx; p.out = output;
    return launch_seg_reduce<true, true, false>(hdim, p, M, "seg_reduce[step2_fwd]", (cudaStream_t)stream);
}

int stb200_attention_step2_with_rel_pos_value_backward_v2(int N, int M, int h, int hdim, int, int L,
                                                          const float *grad_out, const int *index0_offsets,
                                                          const int *index1, const float *attn, const float *v,
                                                          const float *table, const int *rel_idx, float *grad_attn,
                                                          float *grad_v, float *grad_table, const int *t_offsets,
                                                          const int *t_pair, const int *t_index0, void *stream) {
    if (int rc = check_dims(N, M, h, hdim)) return rc;
    STB200_REQUIRE(L > 0 && grad_out && index0_offsets && index1 && attn && v && table && rel_idx && grad_attn &&
                       grad_v && grad_table, STB200_ERR_ARG, "null pointer or L<=0");
    STB200_REQUIRE(t_offsets && t_pair && t_index0, STB200_ERR_ARG, "transposed CSR required (stb200_transpose_csr)");
    cudaStream_t s = (cudaStream_t)stream;
    SegParams p{};
    p.N = N; p.h = h; p.L = L; p.rel_idx = rel_idx;
    p.X = grad_out; p.Y = v; p.offsets = index0_offsets; p.gather_idx = index1; p.Tx = table; p.out = grad_attn;
    if (M > 0)
        if (int rc = launch_seg_dot<true, true, false>(hdim, p, M, "seg_dot[step2_bwd_gattn]", s)) return rc;    // grad_attn = <g, v + Ev>
    p.w = attn; p.out = grad_table;
    if (int rc = launch_table_grad<false>(hdim, p, M, "table_grad[step2_bwd_gtv]", s)) return rc;  // grad_table (+=), X = grad_out
    p.Y = grad_out; p.offsets = t_offsets; p.pair_id = t_pair; p.gather_idx = t_index0; p.out = grad_v;
    p.accumulate = 1;
    return launch_seg_reduce<true, false, true>(hdim, p, M, "seg_reduce_t[step2_bwd_gv]", s);  // grad_v += sum attn*g
}

int stb200_segment_softmax_forward(int N, int M, int h, const float *a, const float *b, const int *index0_offsets,
                                   float *p, void *stream) {
    STB200_REQUIRE(N >= 0 && M >= 0 && h > 0, STB200_ERR_ARG, "bad sizes");
    if (N == 0 || M == 0) return STB200_OK;
    STB200_REQUIRE(a && index0_offsets && p, STB200_ERR_ARG, "null pointer");
    const int blocks = max(1, min((N + kThreads / kWarp - 1) / (kThreads / kWarp), kNumSMs * 8));
    {
        KernelScope ks("segment_softmax_fwd", 4.0 * M * h * (b ? 3 : 2) + 4.0 * (N + 1), (cudaStream_t)stream);
        cudaStream_t s = (cudaStream_t)stream;
        int rc = -1;
        if (h <= 32 && softmax_variant() == 0) rc = launch_softmax_span<false>(s, N, M, h, a, b, index0_offsets, p);
        if (rc > 0) return rc;
        if (rc == 0) {}
        else if (h <= 32) launch_softmax_fwd_hp(blocks, s, N, h, a, b, index0_offsets, p);
        else segment_softmax_fwd_kernel<<<blocks, kThreads, 0, s>>>(N, h, a, b, index0_offsets, p);
    }
    return check_launch("segment_softmax_fwd");
}

int stb200_segment_softmax_forward_rows(int n_rows, const int *rows, int h, const float *a, const float *b,
                                        const int *index0_offsets, float *p, void *stream) {
    STB200_REQUIRE(n_rows >= 0 && h > 0 && h <= 32, STB200_ERR_ARG, "bad sizes (h <= 32)");
    if (n_rows == 0) return STB200_OK;
    STB200_REQUIRE(a && index0_offsets && p, STB200_ERR_ARG, "null pointer");
    const int blocks = max(1, min((n_rows + kThreads / kWarp - 1) / (kThreads / kWarp), kNumSMs * 8));
    cudaStream_t s = (cudaStream_t)stream;
    {
        KernelScope ks("segment_softmax_fwd", 0.0, s);
        if (h <= 32) launch_softmax_fwd_hp(blocks, s, n_rows, h, a, b, index0_offsets, p, rows);
    }
    return check_launch("segment_softmax_fwd_rows");
}

int stb200_segment_softmax_backward(int N, int M, int h, const float *p, const float *grad_p,
                                    const int *index0_offsets, float *grad_s, void *stream) {
    STB200_REQUIRE(N >= 0 && M >= 0 && h > 0, STB200_ERR_ARG, "bad sizes");
    if (N == 0 || M == 0) return STB200_OK;
    STB200_REQUIRE(p && grad_p && index0_offsets && grad_s, STB200_ERR_ARG, "null pointer");
    const int blocks = max(1, min((N + kThreads / kWarp - 1) / (kThreads / kWarp), kNumSMs * 8));
    {
        KernelScope ks("segment_softmax_bwd", 4.0 * M * h * 3 + 4.0 * (N + 1), (cudaStream_t)stream);
        cudaStream_t s = (cudaStream_t)stream;
        int rc = -1;
        // the per-row kernel streams at 3.3 TB/s on layer 0; the span version measured 2.0 TB/s (two staged arrays, three
        // CTAs per SM): used only on request (STB200_SOFTMAX_VARIANT=5)
        if (h <= 32 && softmax_variant() == 5) rc = launch_softmax_span<true>(s, N, M, h, p, grad_p, index0_offsets, grad_s);
        if (rc > 0) return rc;
        if (rc == 0) {}
        else if (h <= 32) launch_softmax_bwd_hp(blocks, s, N, h, p, grad_p, index0_offsets, grad_s);
        else segment_softmax_bwd_kernel<<<blocks, kThreads, 0, s>>>(N, h, p, grad_p, index0_offsets, grad_s);
    }
    return check_launch("segment_softmax_bwd");
}

// ---- fused entry points (new; see include/stb200.h): logits = q.k + rel-pos bias in one pass over the pairs,
// and the same aggregation / gradient kernels driven by an stb200_index (optionally with pre-packed bins).
static int check_index(const stb200_index *ix, bool need_t) {
    STB200_REQUIRE(ix && ix->N >= 0 && ix->M >= 0 && ix->index0_offsets, STB200_ERR_ARG, "bad stb200_index");
    STB200_REQUIRE(ix->M == 0 || (ix->index1 && (ix->rel_idx || ix->rel_packed)), STB200_ERR_ARG, "stb200_index: index1 / rel_idx missing");
    if (need_t) {
        STB200_REQUIRE(ix->t_offsets && ix->t_pair && ix->t_index0, STB200_ERR_ARG, "transposed CSR required (stb200_transpose_csr)");
        STB200_REQUIRE(ix->rel_idx || (ix->rel_packed && ix->t_rel_packed), STB200_ERR_ARG, "stb200_index: rel_idx or both packed arrays required");
    }
    return STB200_OK;
}

int stb200_window_logits_forward(const stb200_index *ix, int h, int hdim, int L, const float *q, const float *k,
                                 const float *table_q, const float *table_k, float *logits, void *stream) {
    if (int rc = check_index(ix, false)) return rc;
    if (int rc = check_dims(ix->N, ix->M, h, hdim)) return rc;
    if (ix->M == 0) return STB200_OK;
    STB200_REQUIRE(L > 0 && L <= 1024 && q && k && table_q && table_k && logits, STB200_ERR_ARG, "null pointer or bad L");
    SegParams p{};
    p.N = ix->N; p.h = h; p.L = L; p.X = q; p.Y = k; p.offsets = ix->index0_offsets; p.gather_idx = ix->index1;
    p.Tx = table_q; p.Ty = table_k; p.rel_idx = ix->rel_idx; p.packed = ix->rel_packed; p.row_order = ix->row_order; p.out = logits;
    return launch_seg_dot<true, true, true>(hdim, p, ix->M, "seg_dot[logits_fwd]", (cudaStream_t)stream);
}

int stb200_window_logits_backward(const stb200_index *ix, int h, int hdim, int L, const float *grad_logits,
                                  const float *q, const float *k, const float *table_q, const float *table_k,
                                  float *grad_q, float *grad_k, float *grad_table_q, float *grad_table_k, void *stream) {
    return stb200_window_logits_backward_ws(ix, h, hdim, L, grad_logits, q, k, table_q, table_k, grad_q, grad_k, grad_table_q,
                                            grad_table_k, nullptr, 0, stream);
}

size_t stb200_window_logits_backward_workspace_bytes(int M, int h) { return (size_t)max(M, 0) * max(h, 0) * sizeof(float) + 256; }

int stb200_window_logits_backward_ws(const stb200_index *ix, int h, int hdim, int L, const float *grad_logits,
                                     const float *q, const float *k, const float *table_q, const float *table_k,
                                     float *grad_q, float *grad_k, float *grad_table_q, float *grad_table_k,
                                     void *workspace, size_t workspace_bytes, void *stream) {
    if (int rc = check_index(ix, true)) return rc;
    if (int rc = check_dims(ix->N, ix->M, h, hdim)) return rc;
    STB200_REQUIRE(L > 0 && L <= 256 && grad_logits && q && k && table_q && table_k && grad_q && grad_k && grad_table_q &&
                       grad_table_k, STB200_ERR_ARG, "null pointer or bad L");
    cudaStream_t s = (cudaStream_t)stream;
    const int M = ix->M;
    SegParams p{};
    p.N = ix->N; p.h = h; p.L = L; p.w = grad_logits; p.rel_idx = ix->rel_idx; p.row_order = ix->row_order;
    // grad_q = sum g * (k[i1] + Eq)                                              (overwritten)
    p.packed = ix->rel_packed; p.offsets = ix->index0_offsets; p.gather_idx = ix->index1; p.Y = k; p.Tx = table_q; p.out = grad_q;
    if (int rc = launch_seg_reduce<true, true, false>(hdim, p, M, "seg_reduce[logits_bwd_gq]", s)) return rc;
    p.X = q; p.out = grad_table_q; p.row_order = ix->len_order;   // tiles of equally long rows (balance), may be null
    if (int rc = launch_table_grad<false>(hdim, p, M, "table_grad[logits_bwd_gtq]", s)) return rc;
    // grad_k = sum over incoming pairs g * (q[i0] + Ek)                           (overwritten)
    p.row_order = ix->row_order;
    p.packed = ix->t_rel_packed; p.offsets = ix->t_offsets; p.pair_id = ix->t_pair; p.gather_idx = ix->t_index0;
    p.Y = q; p.Tx = table_k; p.out = grad_k; p.accumulate = 0;   // fused API: grad_k is overwritten
    // Both key-side kernels read grad_logits through t_pair (12-96 B useful out of every 32 B sector fetched, and a
    // dependent load).  With a workspace the rows are brought into transposed order once and both kernels stream them.
    if (workspace && workspace_bytes >= stb200_window_logits_backward_workspace_bytes(M, h) && M > 0) {
        float *wt = reinterpret_cast<float *>(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
        {
            KernelScope ks("permute_rows[logits_bwd]", 8.0 * M * h + 4.0 * M, s);
            const long long total = (long long)M * h;
            const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)kNumSMs * 16);
            permute_rows_kernel<<<blocks, 256, 0, s>>>(M, h, grad_logits, ix->t_pair, wt);
        }
        if (int rc = check_launch("permute_rows")) return rc;
        p.w = wt;
        p.w_by_slot = 1;
    }
    if (int rc = launch_seg_reduce<true, true, true>(hdim, p, M, "seg_reduce_t[logits_bwd_gk]", s)) return rc;
    p.X = k; p.out = grad_table_k; p.row_order = ix->t_len_order;
    return launch_table_grad<true>(hdim, p, M, "table_grad_t[logits_bwd_gtk]", s);
}

int stb200_window_aggregate_forward(const stb200_index *ix, int h, int hdim, int L, const float *attn, const float *v,
                                    const float *table_v, float *output, void *stream) {
    if (int rc = check_index(ix, false)) return rc;
    if (int rc = check_dims(ix->N, ix->M, h, hdim)) return rc;
    STB200_REQUIRE(L > 0 && L <= 1024 && output && (ix->M == 0 || (attn && v && table_v)), STB200_ERR_ARG, "null pointer or bad L");
    SegParams p{};
    p.N = ix->N; p.h = h; p.L = L; p.w = attn; p.Y = v; p.offsets = ix->index0_offsets; p.gather_idx = ix->index1;
    p.Tx = table_v; p.rel_idx = ix->rel_idx; p.packed = ix->rel_packed; p.row_order = ix->row_order; p.out = output;
    return launch_seg_reduce<true, true, false>(hdim, p, ix->M, "seg_reduce[aggregate_fwd]", (cudaStream_t)stream);
}

int stb200_window_aggregate_backward(const stb200_index *ix, int h, int hdim, int L, const float *grad_out,
                                     const float *attn, const float *v, const float *table_v, float *grad_attn,
                                     float *grad_v, float *grad_table_v, void *stream) {
    if (int rc = check_index(ix, true)) return rc;
    if (int rc = check_dims(ix->N, ix->M, h, hdim)) return rc;
    STB200_REQUIRE(L > 0 && L <= 256 && grad_out && attn && v && table_v && grad_attn && grad_v && grad_table_v, STB200_ERR_ARG,
                   "null pointer or bad L");
    cudaStream_t s = (cudaStream_t)stream;
    const int M = ix->M;
    SegParams p{};
    p.N = ix->N; p.h = h; p.L = L; p.rel_idx = ix->rel_idx; p.packed = ix->rel_packed; p.row_order = ix->row_order;
    p.X = grad_out; p.Y = v; p.offsets = ix->index0_offsets; p.gather_idx = ix->index1; p.Tx = table_v; p.out = grad_attn;
    if (M > 0)
        if (int rc = launch_seg_dot<true, true, false>(hdim, p, M, "seg_dot[aggregate_bwd_gattn]", s)) return rc;
    p.w = attn; p.out = grad_table_v; p.row_order = ix->len_order;
    if (int rc = launch_table_grad<false>(hdim, p, M, "table_grad[aggregate_bwd_gtv]", s)) return rc;
    p.row_order = ix->row_order;
    p.packed = nullptr; p.Y = grad_out; p.offsets = ix->t_offsets; p.pair_id = ix->t_pair; p.gather_idx = ix->t_index0;
    p.out = grad_v; p.accumulate = 0;   // fused API: grad_v is overwritten
    return launch_seg_reduce<true, false, true>(hdim, p, M, "seg_reduce_t[aggregate_bwd_gv]", s);
}

}  // extern "C"
