// Pair-index construction for stratified window attention, entirely on the device.
//
// Specification = the reference's Python (paths under /root/reference):
//   model/stratified_transformer.py:44-65    grid_sample  (voxel_grid -> unique -> p2v_map)
//   model/stratified_transformer.py:10-42    get_indice_pairs (dense pairs per small window; sparse pairs =
//                                            FPS-sampled keys of the 2x window whose window_coord differs)
//   model/stratified_transformer.py:311-317  sort by query, bincount, cumsum -> CSR
//   model/stratified_transformer.py:186-188  relative position index
//   torch_geometric voxel_grid (third party) restated in SURVEY Appendix B.2
// Output order inside a query segment is the canonical one (dense keys ascending by point id, then sparse keys
// ascending by point id); the reference's own order is unspecified because it uses unstable sorts.
//
// The reference materialises [n,k,k] boolean masks and an [n,k,k,3] coordinate compare per block and sorts M
// 64-bit pair ids.  Here: two N-element radix sorts group the points by small / large window, sampled points
// are compacted per large window, then one warp per query counts and (after a scan) emits its keys and the
// rel-pos indices in a single pass.  Nothing M-sized is ever sorted.
//
// fp32 arithmetic that decides integers is reproduced operation by operation with rounding-explicit
// intrinsics (no contraction): voxel ids use truncating division, window_coord and the rel-pos index use
// torch's fmod-based floor division (c10/util/generic_math.h div_floor_floating).
#include <cub/cub.cuh>

#include "common.cuh"

namespace stb200 {

// ---- exact fp32 helpers ------------------------------------------------------------------------------
__device__ __forceinline__ float floor_div_f32(float a, float b) {  // torch `a // b`
    if (b == 0.f) return __fdiv_rn(a, b);
    const float mod = fmodf(a, b);
    float div = __fdiv_rn(__fsub_rn(a, mod), b);
    if (mod != 0.f && ((b < 0.f) != (mod < 0.f))) div = __fsub_rn(div, 1.f);
    if (div != 0.f) {
        float fl = floorf(div);
        if (__fsub_rn(div, fl) > 0.5f) fl = __fadd_rn(fl, 1.f);
        return fl;
    }
    return copysignf(0.f, __fdiv_rn(a, b));
}

__device__ __forceinline__ float remainder_f32(float a, float b) {  // torch `a % b`
    float mod = fmodf(a, b);
    if (mod != 0.f && ((b < 0.f) != (mod < 0.f))) mod = __fadd_rn(mod, b);
    return mod;
}

// `torch.round(rel * 100000) / 100000` (model/stratified_transformer.py:187): on a CUDA tensor torch evaluates a true
// division by a Python scalar as a multiplication by the fp32 reciprocal (measured on the B200 box, tools/dbg_torch_cuda_div.py:
// 30 % of the quotients differ from the IEEE division by one ulp, which moves 1 pair in ~20 000 into the neighbouring bin);
// on a CPU tensor it divides.  The reference only ever runs on the GPU, so the CUDA form is the default; the CPU form
// (stb200_set_torch_semantics(0)) exists to reproduce fixtures generated with CPU torch.  Nothing else on the path differs
// between the two devices (floor division and remainder by a scalar, tensor // tensor: 0 mismatches in 2e7 samples).
__constant__ int c_rel_cuda_division = 1;

__device__ __forceinline__ int rel_index_stratified(float xa, float xb, float two_w, float quant) {
    float r = __fsub_rn(xa, xb);
    r = rintf(__fmul_rn(r, 100000.f));
    r = c_rel_cuda_division ? __fmul_rn(r, 1.0f / 100000.0f) : __fdiv_rn(r, 100000.f);
    const float t = __fsub_rn(__fadd_rn(r, two_w), 0.0001f);
    return (int)floor_div_f32(t, quant);
}

__device__ __forceinline__ unsigned f2ord(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

// ---- builder state (lives in the caller's workspace) ---------------------------------------------------
struct GridParams {       // written by params_kernel
    float mn[3], mx[3];   // global min / max of xyz over the whole batch
    float start[3];       // voxel_grid start (== mn for both parities)
    float shift_s, shift_l, size_s, size_l;
    long long stride_s[4], stride_l[4];
    int err;              // 1: more than 2^32 voxels
    int M, n_max;
};

struct BuilderState {
    GridParams *gp;
    unsigned *mm;                 // ordered-uint min[3], max[3]
    unsigned *key_s, *key_l, *skey_s, *skey_l;
    int *iota, *order_s, *order_l;
    int *flag_s, *flag_l, *rank_s, *rank_l;   // rank = inclusive scan of flags
    int *wstart_s, *wstart_l;     // [N+1] window start positions in order_*
    int *win_s, *win_l;           // window rank of each point
    int *sflag, *spos;            // [N+1] sampled flag over order_l positions and its exclusive scan
    int *samp;                    // sampled points grouped by large window
    unsigned char *ds_mask;       // [N]
    int4 *wc;                     // window_coord per point (x,y,z,unused)
    int *counts;                  // [N+1]
    void *cub_tmp;
    size_t cub_bytes;
};

static size_t al(size_t x) { return (x + 255) & ~(size_t)255; }

static size_t cub_temp_bytes(int N) {
    size_t a = 0, b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (const unsigned *)nullptr, (unsigned *)nullptr, (const int *)nullptr,
                                    (int *)nullptr, N, 0, 32);
    cub::DeviceScan::InclusiveSum(nullptr, b, (const int *)nullptr, (int *)nullptr, N + 1);
    size_t c = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, c, (const int *)nullptr, (int *)nullptr, N + 1);
    if (c > b) b = c;
    return a > b ? a : b;
}

static size_t carve(BuilderState &st, char *base, int N) {
    size_t o = 0;
    auto take = [&](size_t bytes) { char *p = base ? base + o : nullptr; o += al(bytes); return p; };
    const size_t ni = (size_t)(N + 1) * sizeof(int);
    st.gp = (GridParams *)take(sizeof(GridParams));
    st.mm = (unsigned *)take(6 * sizeof(unsigned));
    st.key_s = (unsigned *)take(ni); st.key_l = (unsigned *)take(ni);
    st.skey_s = (unsigned *)take(ni); st.skey_l = (unsigned *)take(ni);
    st.iota = (int *)take(ni); st.order_s = (int *)take(ni); st.order_l = (int *)take(ni);
    st.flag_s = (int *)take(ni); st.flag_l = (int *)take(ni); st.rank_s = (int *)take(ni); st.rank_l = (int *)take(ni);
    st.wstart_s = (int *)take(ni); st.wstart_l = (int *)take(ni);
    st.win_s = (int *)take(ni); st.win_l = (int *)take(ni);
    st.sflag = (int *)take(ni); st.spos = (int *)take(ni); st.samp = (int *)take(ni);
    st.ds_mask = (unsigned char *)take((size_t)N + 1);
    st.wc = (int4 *)take((size_t)(N + 1) * sizeof(int4));
    st.counts = (int *)take(ni);
    st.cub_bytes = cub_temp_bytes(N);
    st.cub_tmp = take(st.cub_bytes);
    return o + 256;
}

// ---- kernels -------------------------------------------------------------------------------------------
__global__ void init_state_kernel(unsigned *mm, GridParams *gp, unsigned char *ds_mask, int N) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 3) mm[i] = 0xffffffffu;
    else if (i < 6) mm[i] = 0u;
    if (i == 0) { gp->err = 0; gp->M = 0; gp->n_max = 0; }
    for (int j = i; j < N; j += gridDim.x * blockDim.x) ds_mask[j] = 0;
}

__global__ void minmax_kernel(int N, const float *__restrict__ xyz, unsigned *mm) {
    unsigned lo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, hi[3] = {0u, 0u, 0u};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const unsigned u = f2ord(__ldg(xyz + (size_t)i * 3 + a));
            lo[a] = min(lo[a], u);
            hi[a] = max(hi[a], u);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        lo[a] = __reduce_min_sync(0xffffffffu, lo[a]);
        hi[a] = __reduce_max_sync(0xffffffffu, hi[a]);
    }
    if (threadIdx.x % kWarp == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            atomicMin(mm + a, lo[a]);
            atomicMax(mm + 3 + a, hi[a]);
        }
    }
}

__global__ void mark_sampled_kernel(int m, const int *__restrict__ ds_idx, unsigned char *ds_mask, int N) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
        const int p = __ldg(ds_idx + i);
        if (p >= 0 && p < N) ds_mask[p] = 1;
    }
}

// voxel_grid strides: k_{d+1} = k_d * (trunc((end_d - start_d) / size_d) + 1); 4th coordinate = batch id, size 1
__global__ void params_kernel(const unsigned *mm, GridParams *gp, float w, int parity, int b) {
    if (threadIdx.x || blockIdx.x) return;
    for (int a = 0; a < 3; ++a) {
        gp->mn[a] = ord2f(mm[a]);
        gp->mx[a] = ord2f(mm[3 + a]);
        gp->start[a] = gp->mn[a];
    }
    const float w2 = __fmul_rn(2.f, w);
    gp->size_s = w;
    gp->size_l = w2;
    gp->shift_s = parity ? __fmul_rn(0.5f, w) : 0.f;
    gp->shift_l = parity ? __fmul_rn(0.5f, w2) : 0.f;
    for (int which = 0; which < 2; ++which) {
        const float size = which ? w2 : w, shift = which ? gp->shift_l : gp->shift_s;
        long long *stride = which ? gp->stride_l : gp->stride_s;
        long long k = 1;
        for (int a = 0; a < 3; ++a) {
            stride[a] = k;
            const float end = parity ? __fadd_rn(gp->mx[a], shift) : gp->mx[a];
            k *= (long long)__fdiv_rn(__fsub_rn(end, gp->start[a]), size) + 1;
        }
        stride[3] = k;
        k *= (long long)(b - 1) + 1;
        if (k >= (1LL << 32) || k <= 0) gp->err = 1;
    }
}

__global__ void keys_kernel(int N, int b, const float *__restrict__ xyz, const int *__restrict__ offset,
                            const GridParams *__restrict__ gp, int parity, unsigned *key_s, unsigned *key_l, int *iota,
                            int4 *wc) {
    const GridParams g = *gp;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        int lo = 0, hi = b - 1;  // batch id = first scene whose cumulative offset exceeds i
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(offset + mid) > i) hi = mid; else lo = mid + 1;
        }
        long long ks = (long long)lo * g.stride_s[3], kl = (long long)lo * g.stride_l[3];
        int4 c = make_int4(0, 0, 0, 0);
        int *cp = &c.x;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float x = __ldg(xyz + (size_t)i * 3 + a);
            const float ps = parity ? __fadd_rn(x, g.shift_s) : x, pl = parity ? __fadd_rn(x, g.shift_l) : x;
            ks += (long long)__fdiv_rn(__fsub_rn(ps, g.start[a]), g.size_s) * g.stride_s[a];
            kl += (long long)__fdiv_rn(__fsub_rn(pl, g.start[a]), g.size_l) * g.stride_l[a];
            // window_coord (get_indice_pairs): (xyz [+ w/2] - xyz_min) // w with torch's floor division
            cp[a] = (int)floor_div_f32(__fsub_rn(ps, g.mn[a]), g.size_s);
        }
        key_s[i] = (unsigned)ks;
        key_l[i] = (unsigned)kl;
        iota[i] = i;
        wc[i] = c;
    }
}

__global__ void flags_kernel(int N, const unsigned *__restrict__ sk_s, const unsigned *__restrict__ sk_l,
                             const int *__restrict__ order_l, const unsigned char *__restrict__ ds_mask, int *flag_s,
                             int *flag_l, int *sflag) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= N; i += gridDim.x * blockDim.x) {
        if (i == N) {
            flag_s[i] = flag_l[i] = sflag[i] = 0;
        } else {
            flag_s[i] = i == 0 || sk_s[i] != sk_s[i - 1];
            flag_l[i] = i == 0 || sk_l[i] != sk_l[i - 1];
            sflag[i] = ds_mask[order_l[i]];
        }
    }
}

__global__ void scatter_windows_kernel(int N, const int *__restrict__ order_s, const int *__restrict__ order_l,
                                       const int *__restrict__ flag_s, const int *__restrict__ flag_l,
                                       const int *__restrict__ rank_s, const int *__restrict__ rank_l,
                                       const int *__restrict__ sflag, const int *__restrict__ spos, int *wstart_s,
                                       int *wstart_l, int *win_s, int *win_l, int *samp) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        const int rs = rank_s[i] - 1, rl = rank_l[i] - 1;
        win_s[order_s[i]] = rs;
        win_l[order_l[i]] = rl;
        if (flag_s[i]) wstart_s[rs] = i;
        if (flag_l[i]) wstart_l[rl] = i;
        if (i == N - 1) {
            wstart_s[rs + 1] = N;
            wstart_l[rl + 1] = N;
        }
        if (sflag[i]) samp[spos[i]] = order_l[i];
    }
}

__device__ __forceinline__ bool wc_differs(int4 a, int4 b) { return a.x != b.x || a.y != b.y || a.z != b.z; }

// one warp per query: number of keys = |small window| + #{sampled b in large window : wc(a) != wc(b)}
__global__ void count_pairs_kernel(int N, const int *__restrict__ win_s, const int *__restrict__ win_l,
                                   const int *__restrict__ wstart_s, const int *__restrict__ wstart_l,
                                   const int *__restrict__ spos, const int *__restrict__ samp,
                                   const int4 *__restrict__ wc, int *counts, GridParams *gp, int has_sparse) {
    const int lane = threadIdx.x % kWarp;
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) / kWarp, nw = gridDim.x * blockDim.x / kWarp;
    int local_max = 0;
    for (int a = wid; a < N; a += nw) {
        const int ws = win_s[a];
        int cnt = wstart_s[ws + 1] - wstart_s[ws];
        if (has_sparse) {
            const int wl = win_l[a];
            const int s0 = spos[wstart_l[wl]], s1 = spos[wstart_l[wl + 1]];
            const int4 ca = wc[a];
            int sparse = 0;
            for (int s = s0 + lane; s < s1; s += kWarp) sparse += wc_differs(ca, wc[samp[s]]);
            cnt += __reduce_add_sync(0xffffffffu, sparse);
        }
        if (lane == 0) counts[a] = cnt;
        local_max = max(local_max, cnt);
    }
    if (lane == 0) {
        if (wid == 0) counts[N] = 0;
        atomicMax(&gp->n_max, local_max);
    }
}

__global__ void finish_count_kernel(int N, const int *__restrict__ offsets, const int *__restrict__ rank_s, GridParams *gp, int *totals) {
    if (threadIdx.x || blockIdx.x) return;
    gp->M = offsets[N];
    totals[0] = offsets[N];
    totals[1] = gp->n_max;
    totals[2] = gp->err;
    totals[3] = rank_s[N - 1];   // number of small windows
}

// one warp per query: emit keys (dense ascending id, then sparse ascending id) + rel-pos index (+ index_0)
__global__ void fill_pairs_kernel(int N, const float *__restrict__ xyz, const int *__restrict__ offsets,
                                  const int *__restrict__ win_s, const int *__restrict__ win_l,
                                  const int *__restrict__ wstart_s, const int *__restrict__ wstart_l,
                                  const int *__restrict__ order_s, const int *__restrict__ spos,
                                  const int *__restrict__ samp, const int4 *__restrict__ wc, int has_sparse,
                                  float two_w, float quant, int *__restrict__ index_1, int *__restrict__ rel_idx,
                                  int *__restrict__ index_0) {
    const int lane = threadIdx.x % kWarp;
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) / kWarp, nw = gridDim.x * blockDim.x / kWarp;
    for (int a = wid; a < N; a += nw) {
        const float xa = __ldg(xyz + (size_t)a * 3), ya = __ldg(xyz + (size_t)a * 3 + 1), za = __ldg(xyz + (size_t)a * 3 + 2);
        int out = offsets[a];
        auto emit = [&](int pos, int b) {
            index_1[pos] = b;
            if (index_0) index_0[pos] = a;
            if (rel_idx) {
                const float xb = __ldg(xyz + (size_t)b * 3), yb = __ldg(xyz + (size_t)b * 3 + 1), zb = __ldg(xyz + (size_t)b * 3 + 2);
                rel_idx[(size_t)pos * 3 + 0] = rel_index_stratified(xa, xb, two_w, quant);
                rel_idx[(size_t)pos * 3 + 1] = rel_index_stratified(ya, yb, two_w, quant);
                rel_idx[(size_t)pos * 3 + 2] = rel_index_stratified(za, zb, two_w, quant);
            }
        };
        const int ws = win_s[a];
        const int d0 = wstart_s[ws], d1 = wstart_s[ws + 1];
        for (int i = d0 + lane; i < d1; i += kWarp) emit(out + (i - d0), order_s[i]);
        out += d1 - d0;
        if (has_sparse) {
            const int wl = win_l[a];
            const int s0 = spos[wstart_l[wl]], s1 = spos[wstart_l[wl + 1]];
            const int4 ca = wc[a];
            for (int sb = s0; sb < s1; sb += kWarp) {
                const int s = sb + lane;
                int b = 0;
                bool keep = false;
                if (s < s1) {
                    b = samp[s];
                    keep = wc_differs(ca, wc[b]);
                }
                const unsigned bal = __ballot_sync(0xffffffffu, keep);
                if (keep) emit(out + __popc(bal & ((1u << lane) - 1u)), b);
                out += __popc(bal);
            }
        }
    }
}

// stand-alone rel-pos index for an existing CSR pair list (drop-in for stratified_transformer.py:186-188)
__global__ void rel_index_csr_kernel(int N, const float *__restrict__ xyz, const int *__restrict__ offsets,
                                     const int *__restrict__ index_1, float two_w, float quant, int *__restrict__ rel_idx) {
    const int lane = threadIdx.x % kWarp;
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) / kWarp, nw = gridDim.x * blockDim.x / kWarp;
    for (int a = wid; a < N; a += nw) {
        const float xa = __ldg(xyz + (size_t)a * 3), ya = __ldg(xyz + (size_t)a * 3 + 1), za = __ldg(xyz + (size_t)a * 3 + 2);
        const int s = offsets[a], e = offsets[a + 1];
        for (int m = s + lane; m < e; m += kWarp) {
            const int b = __ldg(index_1 + m);
            rel_idx[(size_t)m * 3 + 0] = rel_index_stratified(xa, __ldg(xyz + (size_t)b * 3), two_w, quant);
            rel_idx[(size_t)m * 3 + 1] = rel_index_stratified(ya, __ldg(xyz + (size_t)b * 3 + 1), two_w, quant);
            rel_idx[(size_t)m * 3 + 2] = rel_index_stratified(za, __ldg(xyz + (size_t)b * 3 + 2), two_w, quant);
        }
    }
}

// Swin variant (swin3d_transformer.py:151-154,129-130): per-point quantised coordinate, then a difference
__global__ void swin_quant_kernel(int N, const float *__restrict__ xyz, const unsigned *__restrict__ mm, float shift,
                                  float w, float quant, float *__restrict__ xq) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N * 3; i += gridDim.x * blockDim.x) {
        const float mn = ord2f(mm[i % 3]);
        const float t = __fadd_rn(__fsub_rn(__ldg(xyz + i), mn), shift);
        xq[i] = floor_div_f32(remainder_f32(t, w), quant);
    }
}
__global__ void swin_rel_kernel(int N, const float *__restrict__ xq, const int *__restrict__ offsets,
                                const int *__restrict__ index_1, float bias, int *__restrict__ rel_idx) {
    const int lane = threadIdx.x % kWarp;
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) / kWarp, nw = gridDim.x * blockDim.x / kWarp;
    for (int a = wid; a < N; a += nw) {
        const int s = offsets[a], e = offsets[a + 1];
        for (int m = s + lane; m < e; m += kWarp) {
            const int b = __ldg(index_1 + m);
#pragma unroll
            for (int ax = 0; ax < 3; ++ax)
                rel_idx[(size_t)m * 3 + ax] = (int)__fadd_rn(__fsub_rn(xq[(size_t)a * 3 + ax], xq[(size_t)b * 3 + ax]), bias);
        }
    }
}


// ---- work plan of the fused window-attention kernels (fused_window.cu; spec: oracle/fused_plan_oracle.py) -----------
// Dense tiles: per small window the [n x n] matrix of packed rel-pos bins; sparse tiles: per large window the
// [n_V x n_s] matrix (queries = its points, keys = its sampled points) with bit 31 set where get_indice_pairs drops the
// pair (equal window_coord, model/stratified_transformer.py:28-35).  Items = blocks of <= BQ x BK of those matrices.
constexpr int kPlanSeg = 128;     // small windows per greedy-packing segment
constexpr int kPlanMaxOrd = 8;    // key chunks per window the launch sequence supports
enum { PT_MDENSE = 0, PT_STOT = 1, PT_NWIN_L = 2, PT_MAXWIN = 3, PT_MAXNS = 4, PT_ERR = 5, PT_BINMIN = 6, PT_BINMAX = 7,
       PT_DENSE_ITEMS = 8, PT_SPARSE_ITEMS = 16, PT_CURSOR_D = 24, PT_CURSOR_S = 32, PT_INTS = 40 };

struct PlanItem { int q_pos, nq, k_pos, nk, rel_off, rel_pitch, flags, pad; };   // == stb200::fw::Item
enum { PF_PACKED = 1, PF_FIRST = 2, PF_FINAL = 4, PF_KEY_ATOMIC = 8 };

struct PlanScratch {
    long long *sq, *sv;      // [N+1] tile sizes (n^2 per small window, n_V*n_s per large window) -> exclusive scans
    long long *tb, *sb;      // tile bases
    float *xq;               // [N,3] Swin per-point quantised coordinates
    void *cub_tmp;           // scan temp storage (64-bit sums: more than the builder's int scans need for small N)
    size_t cub_bytes;
};
static size_t plan_carve(PlanScratch &ps, char *base, int N) {
    size_t o = 0;
    auto take = [&](size_t bytes) { char *p = base ? base + o : nullptr; o += al(bytes); return p; };
    const size_t nl = (size_t)(N + 1) * sizeof(long long);
    ps.sq = (long long *)take(nl); ps.sv = (long long *)take(nl); ps.tb = (long long *)take(nl); ps.sb = (long long *)take(nl);
    ps.xq = (float *)take((size_t)N * 3 * sizeof(float));
    ps.cub_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, ps.cub_bytes, (const long long *)nullptr, (long long *)nullptr, N + 1);
    ps.cub_tmp = take(ps.cub_bytes);
    return o + 256;
}

// walks the small windows of one segment the way oracle/fused_plan_oracle.py does; EMIT = false only counts
template <bool EMIT>
__device__ void plan_dense_segment(int seg, int n_win, const int *__restrict__ wstart, const long long *__restrict__ tb, int BQ,
                                   int BK, int final_flag, int *totals, const int *__restrict__ ord_off, PlanItem *items) {
    const int w0 = seg * kPlanSeg, w1 = min(w0 + kPlanSeg, n_win);
    int cur_start = -1, cur_rows = 0, n_packed = 0, max_win = 0;
    auto flush = [&]() {
        if (cur_rows) {
            if (EMIT) {
                const int slot = ord_off[0] + atomicAdd(&totals[PT_CURSOR_D], 1);
                items[slot] = PlanItem{cur_start, cur_rows, cur_start, cur_rows, 0, 0, PF_PACKED | PF_FIRST | final_flag, 0};
            }
            ++n_packed;
        }
        cur_start = -1;
        cur_rows = 0;
    };
    for (int w = w0; w < w1; ++w) {
        const int ws = wstart[w], n = wstart[w + 1] - ws;
        max_win = max(max_win, n);
        if (n > BK) {
            flush();
            const int nc = (n + BK - 1) / BK;
            if (nc > kPlanMaxOrd) { if (!EMIT) atomicOr(&totals[PT_ERR], 1); continue; }
            for (int kc = 0; kc < nc; ++kc) {
                if (!EMIT) { atomicAdd(&totals[PT_DENSE_ITEMS + kc], nc); continue; }
                for (int qc = 0; qc < nc; ++qc) {
                    const int qa = qc * BQ, ka = kc * BK;
                    const int slot = ord_off[kc] + atomicAdd(&totals[PT_CURSOR_D + kc], 1);
                    items[slot] = PlanItem{ws + qa, min(BQ, n - qa), ws + ka, min(BK, n - ka), (int)(tb[w] + (long long)qa * n + ka), n,
                                           (kc == 0 ? PF_FIRST : 0) | PF_KEY_ATOMIC | (kc == nc - 1 ? final_flag : 0), 0};
                }
            }
            continue;
        }
        if (cur_rows + n > BQ) flush();
        if (cur_rows == 0) cur_start = ws;
        cur_rows += n;
    }
    flush();
    if (!EMIT) {
        if (n_packed) atomicAdd(&totals[PT_DENSE_ITEMS], n_packed);
        atomicMax(&totals[PT_MAXWIN], max_win);
    }
}

__global__ void plan_count_kernel(int N, const int *__restrict__ rank_s, const int *__restrict__ rank_l,
                                  const int *__restrict__ wstart_s, const int *__restrict__ wstart_l, const int *__restrict__ spos,
                                  int has_sparse, int BQ, int BK, int BQS, int BKS, long long *sq, long long *sv, int *totals) {
    const int n_win = rank_s[N - 1], n_win_l = rank_l[N - 1];
    const int t = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    for (int w = t; w <= N; w += nt) {
        long long a = 0, b = 0;
        if (w < n_win) { const long long n = wstart_s[w + 1] - wstart_s[w]; a = n * n; }
        if (has_sparse && w < n_win_l) {
            const long long nv = wstart_l[w + 1] - wstart_l[w], ns = spos[wstart_l[w + 1]] - spos[wstart_l[w]];
            b = nv * ns;
            const int nkc = max(1, (int)((ns + BKS - 1) / BKS)), nqc = (int)((nv + BQS - 1) / BQS);
            if (nkc > kPlanMaxOrd) atomicOr(&totals[PT_ERR], 2);
            else for (int kc = 0; kc < nkc; ++kc) atomicAdd(&totals[PT_SPARSE_ITEMS + kc], nqc);
            atomicMax(&totals[PT_MAXNS], (int)ns);
        }
        sq[w] = a;
        sv[w] = b;
    }
    const int n_seg = (n_win + kPlanSeg - 1) / kPlanSeg;
    for (int seg = t; seg < n_seg; seg += nt) plan_dense_segment<false>(seg, n_win, wstart_s, nullptr, BQ, BK, 0, totals, nullptr, nullptr);
    if (t == 0) totals[PT_NWIN_L] = has_sparse ? n_win_l : 0;
}

__global__ void plan_totals_kernel(int N, const long long *__restrict__ tb, const long long *__restrict__ sb, int *totals) {
    if (threadIdx.x || blockIdx.x) return;
    if (tb[N] >= (1LL << 31) || sb[N] >= (1LL << 31)) totals[PT_ERR] |= 4;
    totals[PT_MDENSE] = (int)tb[N];
    totals[PT_STOT] = (int)sb[N];
}

__global__ void plan_items_kernel(int N, const int *__restrict__ rank_s, const int *__restrict__ rank_l,
                                  const int *__restrict__ wstart_s, const int *__restrict__ wstart_l, const int *__restrict__ spos,
                                  int has_sparse, int BQ, int BK, int BQS, int BKS, const long long *__restrict__ tb,
                                  const long long *__restrict__ sb, int *totals, PlanItem *dense_items, PlanItem *sparse_items) {
    __shared__ int ord_d[kPlanMaxOrd], ord_s[kPlanMaxOrd];
    if (threadIdx.x == 0) {
        int a = 0, b = 0;
        for (int o = 0; o < kPlanMaxOrd; ++o) {
            ord_d[o] = a; a += totals[PT_DENSE_ITEMS + o];
            ord_s[o] = b; b += totals[PT_SPARSE_ITEMS + o];
        }
    }
    __syncthreads();
    const int n_win = rank_s[N - 1], n_win_l = rank_l[N - 1];
    const int t = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    const int n_seg = (n_win + kPlanSeg - 1) / kPlanSeg;
    for (int seg = t; seg < n_seg; seg += nt)
        plan_dense_segment<true>(seg, n_win, wstart_s, tb, BQ, BK, has_sparse ? 0 : PF_FINAL, totals, ord_d, dense_items);
    if (has_sparse)
        for (int v = t; v < n_win_l; v += nt) {
            const int vs = wstart_l[v], nv = wstart_l[v + 1] - vs;
            const int ss = spos[vs], ns = spos[wstart_l[v + 1]] - ss;
            const int nkc = max(1, (ns + BKS - 1) / BKS), nqc = (nv + BQS - 1) / BQS;
            if (nkc > kPlanMaxOrd) continue;
            for (int kc = 0; kc < nkc; ++kc)
                for (int qc = 0; qc < nqc; ++qc) {
                    const int qa = qc * BQS, ka = kc * BKS;
                    const int slot = ord_s[kc] + atomicAdd(&totals[PT_CURSOR_S + kc], 1);
                    sparse_items[slot] = PlanItem{vs + qa, min(BQS, nv - qa), ss + ka, max(0, min(BKS, ns - ka)),
                                                  (int)(sb[v] + (long long)qa * ns + ka), ns, PF_KEY_ATOMIC | (kc == nkc - 1 ? PF_FINAL : 0), 0};
                }
        }
}

__device__ __forceinline__ unsigned pack_bins(int r0, int r1, int r2, int *bmin, int *bmax) {
    *bmin = min(*bmin, min(r0, min(r1, r2)));
    *bmax = max(*bmax, max(r0, max(r1, r2)));
    return (unsigned)(r0 & 0xff) | ((unsigned)(r1 & 0xff) << 8) | ((unsigned)(r2 & 0xff) << 16);
}

// one warp per sorted position: the row of its window's dense tile (and of its large window's sparse tile)
__global__ void plan_tiles_kernel(int N, const float *__restrict__ xyz, const int *__restrict__ order_s, const int *__restrict__ rank_s,
                                  const int *__restrict__ wstart_s, const long long *__restrict__ tb, const int *__restrict__ order_l,
                                  const int *__restrict__ rank_l, const int *__restrict__ wstart_l, const int *__restrict__ spos,
                                  const int *__restrict__ samp, const int4 *__restrict__ wc, const long long *__restrict__ sb,
                                  int has_sparse, int swin, const float *__restrict__ xq, float swin_bias, float two_w, float quant,
                                  unsigned *__restrict__ drel, unsigned *__restrict__ srel, int *__restrict__ pos_win, int *totals) {
    const int lane = threadIdx.x % kWarp;
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) / kWarp, nw = gridDim.x * blockDim.x / kWarp;
    int bmin = 1 << 30, bmax = -(1 << 30);
    for (int p = wid; p < N; p += nw) {
        {   // dense row
            const int a = order_s[p], w = rank_s[p] - 1;
            const int ws = wstart_s[w], n = wstart_s[w + 1] - ws;
            if (lane == 0) pos_win[p] = w;
            unsigned *row = drel + tb[w] + (long long)(p - ws) * n;
            if (swin) {
                const float xa = xq[(size_t)a * 3], ya = xq[(size_t)a * 3 + 1], za = xq[(size_t)a * 3 + 2];
                for (int t = lane; t < n; t += kWarp) {
                    const int b = order_s[ws + t];
                    row[t] = pack_bins((int)__fadd_rn(__fsub_rn(xa, xq[(size_t)b * 3]), swin_bias),
                                       (int)__fadd_rn(__fsub_rn(ya, xq[(size_t)b * 3 + 1]), swin_bias),
                                       (int)__fadd_rn(__fsub_rn(za, xq[(size_t)b * 3 + 2]), swin_bias), &bmin, &bmax);
                }
            } else {
                const float xa = __ldg(xyz + (size_t)a * 3), ya = __ldg(xyz + (size_t)a * 3 + 1), za = __ldg(xyz + (size_t)a * 3 + 2);
                for (int t = lane; t < n; t += kWarp) {
                    const int b = order_s[ws + t];
                    row[t] = pack_bins(rel_index_stratified(xa, __ldg(xyz + (size_t)b * 3), two_w, quant),
                                       rel_index_stratified(ya, __ldg(xyz + (size_t)b * 3 + 1), two_w, quant),
                                       rel_index_stratified(za, __ldg(xyz + (size_t)b * 3 + 2), two_w, quant), &bmin, &bmax);
                }
            }
        }
        if (has_sparse) {
            const int a = order_l[p], v = rank_l[p] - 1;
            const int vs = wstart_l[v];
            const int s0 = spos[vs], ns = spos[wstart_l[v + 1]] - s0;
            if (ns > 0) {
                const float xa = __ldg(xyz + (size_t)a * 3), ya = __ldg(xyz + (size_t)a * 3 + 1), za = __ldg(xyz + (size_t)a * 3 + 2);
                const int4 ca = wc[a];
                unsigned *row = srel + sb[v] + (long long)(p - vs) * ns;
                int d0 = 0, d1 = 0;   // the sparse pass stages the whole bin range: its extrema are not tracked
                for (int t = lane; t < ns; t += kWarp) {
                    const int b = samp[s0 + t];
                    unsigned wd = pack_bins(rel_index_stratified(xa, __ldg(xyz + (size_t)b * 3), two_w, quant),
                                            rel_index_stratified(ya, __ldg(xyz + (size_t)b * 3 + 1), two_w, quant),
                                            rel_index_stratified(za, __ldg(xyz + (size_t)b * 3 + 2), two_w, quant), &d0, &d1);
                    if (!wc_differs(ca, wc[b])) wd |= 0x80000000u;
                    row[t] = wd;
                }
            }
        }
    }
    bmin = __reduce_min_sync(0xffffffffu, bmin);
    bmax = __reduce_max_sync(0xffffffffu, bmax);
    if (lane == 0 && bmin <= bmax) {
        atomicMin(&totals[PT_BINMIN], bmin);
        atomicMax(&totals[PT_BINMAX], bmax);
    }
}

__global__ void plan_init_totals_kernel(int *totals) {
    const int i = threadIdx.x;
    if (i < PT_INTS) totals[i] = i == PT_BINMIN ? (1 << 30) : (i == PT_BINMAX ? -(1 << 30) : 0);
}

__global__ void swin_quant_gp_kernel(int N, const float *__restrict__ xyz, const GridParams *__restrict__ gp, float shift, float w,
                                     float quant, float *__restrict__ xq) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N * 3; i += gridDim.x * blockDim.x) {
        const float t = __fadd_rn(__fsub_rn(__ldg(xyz + i), gp->mn[i % 3]), shift);
        xq[i] = floor_div_f32(remainder_f32(t, w), quant);
    }
}


__global__ void plan_export_kernel(int N, const int *__restrict__ rank_s, const long long *__restrict__ tb, const int *__restrict__ order_s,
                                   const int *__restrict__ wstart_s, const int *__restrict__ order_l, const int *__restrict__ samp,
                                   int n_samp, int has_sparse, int *tile_base, int *o_order_s, int *o_wstart_s, int *o_order_l, int *o_samp) {
    const int n_win = rank_s[N - 1];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= N; i += gridDim.x * blockDim.x) {
        if (i < n_win) tile_base[i] = (int)tb[i];
        if (i <= n_win) o_wstart_s[i] = wstart_s[i];
        if (i < N) {
            o_order_s[i] = order_s[i];
            if (has_sparse) o_order_l[i] = order_l[i];
        }
        if (has_sparse && i < n_samp) o_samp[i] = samp[i];
    }
}

static int blocks_for(long long n, int per_block = 256, int cap = kNumSMs * 8) {
    return (int)max(1LL, min((n + per_block - 1) / per_block, (long long)cap));
}

}  // namespace stb200

using namespace stb200;

extern "C" {

size_t stb200_pair_builder_workspace_bytes(int N) {
    BuilderState st;
    return carve(st, nullptr, N < 1 ? 1 : N);
}

int stb200_stratified_pairs_count(int N, int b, const float *xyz, const int *offset, float window_size, int parity,
                                  const int *downsample_idx, int m, void *workspace, size_t workspace_bytes,
                                  int *index0_offsets, int *totals, void *stream) {
    STB200_REQUIRE(N > 0 && b > 0 && m >= 0, STB200_ERR_ARG, "bad sizes N=%d b=%d m=%d", N, b, m);
    STB200_REQUIRE(xyz && offset && workspace && index0_offsets && totals && (m == 0 || downsample_idx), STB200_ERR_ARG,
                   "null pointer");
    STB200_REQUIRE(workspace_bytes >= stb200_pair_builder_workspace_bytes(N), STB200_ERR_WORKSPACE,
                   "workspace too small: %zu < %zu", workspace_bytes, stb200_pair_builder_workspace_bytes(N));
    cudaStream_t s = (cudaStream_t)stream;
    BuilderState st;
    carve(st, (char *)(((uintptr_t)workspace + 255) & ~(uintptr_t)255), N);
    const int has_sparse = m > 0;
    const int gb = blocks_for(N);
    KernelScope ks("pair_builder_count[20 launches]", 0.0, s);
    count_launch(19);
    init_state_kernel<<<gb, 256, 0, s>>>(st.mm, st.gp, st.ds_mask, N);
    minmax_kernel<<<blocks_for(N, 256, kNumSMs * 2), 256, 0, s>>>(N, xyz, st.mm);
    if (has_sparse) mark_sampled_kernel<<<blocks_for(m), 256, 0, s>>>(m, downsample_idx, st.ds_mask, N);
    params_kernel<<<1, 32, 0, s>>>(st.mm, st.gp, window_size, parity & 1, b);
    keys_kernel<<<gb, 256, 0, s>>>(N, b, xyz, offset, st.gp, parity & 1, st.key_s, st.key_l, st.iota, st.wc);
    size_t tb = st.cub_bytes;
    cudaError_t ce;
#define STB200_CUB(call)                                                                                         \
    do {                                                                                                         \
        tb = st.cub_bytes;                                                                                       \
        if ((ce = (call)) != cudaSuccess) {                                                                      \
            set_error("pair builder: %s failed: %s", #call, cudaGetErrorString(ce));                             \
            return STB200_ERR_CUDA;                                                                              \
        }                                                                                                        \
    } while (0)
    STB200_CUB(cub::DeviceRadixSort::SortPairs(st.cub_tmp, tb, st.key_s, st.skey_s, st.iota, st.order_s, N, 0, 32, s));
    STB200_CUB(cub::DeviceRadixSort::SortPairs(st.cub_tmp, tb, st.key_l, st.skey_l, st.iota, st.order_l, N, 0, 32, s));
    flags_kernel<<<gb, 256, 0, s>>>(N, st.skey_s, st.skey_l, st.order_l, st.ds_mask, st.flag_s, st.flag_l, st.sflag);
    STB200_CUB(cub::DeviceScan::InclusiveSum(st.cub_tmp, tb, st.flag_s, st.rank_s, N + 1, s));
    STB200_CUB(cub::DeviceScan::InclusiveSum(st.cub_tmp, tb, st.flag_l, st.rank_l, N + 1, s));
    STB200_CUB(cub::DeviceScan::ExclusiveSum(st.cub_tmp, tb, st.sflag, st.spos, N + 1, s));
    scatter_windows_kernel<<<gb, 256, 0, s>>>(N, st.order_s, st.order_l, st.flag_s, st.flag_l, st.rank_s, st.rank_l,
                                              st.sflag, st.spos, st.wstart_s, st.wstart_l, st.win_s, st.win_l, st.samp);
    count_pairs_kernel<<<blocks_for((long long)N * kWarp), 256, 0, s>>>(N, st.win_s, st.win_l, st.wstart_s, st.wstart_l,
                                                                       st.spos, st.samp, st.wc, st.counts, st.gp,
                                                                       has_sparse);
    STB200_CUB(cub::DeviceScan::ExclusiveSum(st.cub_tmp, tb, st.counts, index0_offsets, N + 1, s));
#undef STB200_CUB
    finish_count_kernel<<<1, 32, 0, s>>>(N, index0_offsets, st.rank_s, st.gp, totals);
    return check_launch("stratified_pairs_count");
}

int stb200_stratified_pairs_fill(int N, const float *xyz, float window_size_x2, float quant_size, int has_sparse,
                                 void *workspace, size_t workspace_bytes, const int *index0_offsets, int *index_1,
                                 int *rel_idx, int *index_0, int *row_order, int *win_offsets, int n_win, int M, void *stream) {
    STB200_REQUIRE(N > 0 && xyz && workspace && index0_offsets && index_1, STB200_ERR_ARG, "null pointer / bad N");
    STB200_REQUIRE(workspace_bytes >= stb200_pair_builder_workspace_bytes(N), STB200_ERR_WORKSPACE, "workspace too small");
    BuilderState st;
    carve(st, (char *)(((uintptr_t)workspace + 255) & ~(uintptr_t)255), N);
    {
        // writes: index_1 (+ rel_idx, index_0); reads: xyz, offsets, window membership per point
        const double bytes = (double)M * (4 + (rel_idx ? 12 : 0) + (index_0 ? 4 : 0)) + (double)N * (12 + 4 + 8 + 16);
        KernelScope ks("pair_builder_fill", bytes, (cudaStream_t)stream);
        fill_pairs_kernel<<<blocks_for((long long)N * kWarp), 256, 0, (cudaStream_t)stream>>>(
            N, xyz, index0_offsets, st.win_s, st.win_l, st.wstart_s, st.wstart_l, st.order_s, st.spos, st.samp, st.wc,
            has_sparse, window_size_x2, quant_size, index_1, rel_idx, index_0);
    }
    cudaError_t ce = cudaSuccess;
    if (row_order) ce = cudaMemcpyAsync(row_order, st.order_s, (size_t)N * sizeof(int), cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
    if (ce == cudaSuccess && win_offsets && n_win > 0)
        ce = cudaMemcpyAsync(win_offsets, st.wstart_s, (size_t)(n_win + 1) * sizeof(int), cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
    STB200_REQUIRE(ce == cudaSuccess, STB200_ERR_CUDA, "pair builder: copy of the window order failed: %s", cudaGetErrorString(ce));
    return check_launch("stratified_pairs_fill");
}

int stb200_rel_pos_index_stratified(int N, const float *xyz, const int *index0_offsets, const int *index_1,
                                    float window_size_x2, float quant_size, int *rel_idx, void *stream) {
    STB200_REQUIRE(N >= 0, STB200_ERR_ARG, "bad N");
    if (N == 0) return STB200_OK;
    STB200_REQUIRE(xyz && index0_offsets && index_1 && rel_idx, STB200_ERR_ARG, "null pointer");
    {
        KernelScope ks("rel_pos_index_stratified", 0.0, (cudaStream_t)stream);
        rel_index_csr_kernel<<<blocks_for((long long)N * kWarp), 256, 0, (cudaStream_t)stream>>>(
            N, xyz, index0_offsets, index_1, window_size_x2, quant_size, rel_idx);
    }
    return check_launch("rel_pos_index_stratified");
}

int stb200_rel_pos_index_swin(int N, const float *xyz, const int *index0_offsets, const int *index_1,
                              float window_size, float quant_size, float shift_size, int quant_grid_length,
                              float *xq_scratch /*[N,3]*/, unsigned *mm_scratch /*[6]*/, int *rel_idx, void *stream) {
    STB200_REQUIRE(N >= 0, STB200_ERR_ARG, "bad N");
    if (N == 0) return STB200_OK;
    STB200_REQUIRE(xyz && index0_offsets && index_1 && rel_idx && xq_scratch && mm_scratch, STB200_ERR_ARG, "null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    KernelScope ks("rel_pos_index_swin[3 launches]", 0.0, s);
    count_launch(2);
    cudaMemsetAsync(mm_scratch, 0xff, 3 * sizeof(unsigned), s);
    cudaMemsetAsync(mm_scratch + 3, 0, 3 * sizeof(unsigned), s);
    minmax_kernel<<<blocks_for(N, 256, kNumSMs * 2), 256, 0, s>>>(N, xyz, mm_scratch);
    swin_quant_kernel<<<blocks_for((long long)N * 3), 256, 0, s>>>(N, xyz, mm_scratch, shift_size, window_size, quant_size, xq_scratch);
    swin_rel_kernel<<<blocks_for((long long)N * kWarp), 256, 0, s>>>(N, xq_scratch, index0_offsets, index_1,
                                                                    (float)(quant_grid_length - 1), rel_idx);
    return check_launch("rel_pos_index_swin");
}


int stb200_set_torch_semantics(int cuda) {
    const int v = cuda ? 1 : 0;
    const cudaError_t e = cudaMemcpyToSymbol(c_rel_cuda_division, &v, sizeof(int));
    STB200_REQUIRE(e == cudaSuccess, STB200_ERR_CUDA, "set_torch_semantics: %s", cudaGetErrorString(e));
    return STB200_OK;
}

// ---- fused work plan (include/stb200.h) ------------------------------------------------------------------------------
size_t stb200_fused_plan_scratch_bytes(int N) {
    PlanScratch ps;
    return plan_carve(ps, nullptr, N < 1 ? 1 : N);
}

int stb200_fused_plan_count(int N, void *workspace, size_t workspace_bytes, int has_sparse, int BQ, int BK, int BQS, int BKS,
                            void *scratch, size_t scratch_bytes, int *totals, void *stream) {
    STB200_REQUIRE(N > 0 && workspace && scratch && totals, STB200_ERR_ARG, "null pointer / bad N");
    STB200_REQUIRE(BQ == BK && BQ > 0 && BQS > 0 && BKS > 0, STB200_ERR_ARG, "dense blocks are square (BQ == BK)");
    STB200_REQUIRE(workspace_bytes >= stb200_pair_builder_workspace_bytes(N), STB200_ERR_WORKSPACE, "builder workspace too small");
    STB200_REQUIRE(scratch_bytes >= stb200_fused_plan_scratch_bytes(N), STB200_ERR_WORKSPACE, "plan scratch too small");
    cudaStream_t s = (cudaStream_t)stream;
    BuilderState st;
    carve(st, (char *)(((uintptr_t)workspace + 255) & ~(uintptr_t)255), N);
    PlanScratch ps;
    plan_carve(ps, (char *)(((uintptr_t)scratch + 255) & ~(uintptr_t)255), N);
    KernelScope ks("fused_plan_count[5 launches]", 0.0, s);
    count_launch(4);
    plan_init_totals_kernel<<<1, 64, 0, s>>>(totals);
    plan_count_kernel<<<blocks_for(N + 1), 256, 0, s>>>(N, st.rank_s, st.rank_l, st.wstart_s, st.wstart_l, st.spos, has_sparse ? 1 : 0,
                                                        BQ, BK, BQS, BKS, ps.sq, ps.sv, totals);
    size_t tb = ps.cub_bytes;
    cudaError_t e = cub::DeviceScan::ExclusiveSum(ps.cub_tmp, tb, ps.sq, ps.tb, N + 1, s);
    STB200_REQUIRE(e == cudaSuccess, STB200_ERR_CUDA, "cub scan: %s", cudaGetErrorString(e));
    tb = ps.cub_bytes;
    e = cub::DeviceScan::ExclusiveSum(ps.cub_tmp, tb, ps.sv, ps.sb, N + 1, s);
    STB200_REQUIRE(e == cudaSuccess, STB200_ERR_CUDA, "cub scan: %s", cudaGetErrorString(e));
    plan_totals_kernel<<<1, 32, 0, s>>>(N, ps.tb, ps.sb, totals);
    return check_launch("fused_plan_count");
}

int stb200_fused_plan_fill(int N, const float *xyz, float window_size_x2, float quant_size, int has_sparse, int BQ, int BK, int BQS,
                           int BKS, int swin, float swin_window, float swin_shift, void *workspace, size_t workspace_bytes,
                           void *scratch, size_t scratch_bytes, int *totals, unsigned *dense_rel, int *tile_base, int *pos_win,
                           int *order_s, int *wstart_s, void *dense_items, unsigned *sparse_rel, int *order_l, int *samp, int n_samp,
                           void *sparse_items, void *stream) {
    STB200_REQUIRE(N > 0 && xyz && workspace && scratch && totals && dense_rel && tile_base && pos_win && order_s && wstart_s && dense_items,
                   STB200_ERR_ARG, "null pointer / bad N");
    STB200_REQUIRE(!has_sparse || (sparse_rel && order_l && samp && sparse_items), STB200_ERR_ARG, "sparse outputs missing");
    STB200_REQUIRE(workspace_bytes >= stb200_pair_builder_workspace_bytes(N) && scratch_bytes >= stb200_fused_plan_scratch_bytes(N),
                   STB200_ERR_WORKSPACE, "workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    BuilderState st;
    carve(st, (char *)(((uintptr_t)workspace + 255) & ~(uintptr_t)255), N);
    PlanScratch ps;
    plan_carve(ps, (char *)(((uintptr_t)scratch + 255) & ~(uintptr_t)255), N);
    KernelScope ks("fused_plan_fill[4 launches]", 0.0, s);
    count_launch(3);
    if (swin) {
        const int qgl = (int)(swin_window / quant_size);
        swin_quant_gp_kernel<<<blocks_for((long long)N * 3), 256, 0, s>>>(N, xyz, st.gp, swin_shift, swin_window, quant_size, ps.xq);
        (void)qgl;
    }
    plan_items_kernel<<<blocks_for(N), 256, 0, s>>>(N, st.rank_s, st.rank_l, st.wstart_s, st.wstart_l, st.spos, has_sparse ? 1 : 0, BQ, BK,
                                                    BQS, BKS, ps.tb, ps.sb, totals, (PlanItem *)dense_items, (PlanItem *)sparse_items);
    const float swin_bias = swin ? (float)((int)(swin_window / quant_size) - 1) : 0.f;
    plan_tiles_kernel<<<blocks_for((long long)N * kWarp), 256, 0, s>>>(N, xyz, st.order_s, st.rank_s, st.wstart_s, ps.tb, st.order_l,
                                                                      st.rank_l, st.wstart_l, st.spos, st.samp, st.wc, ps.sb,
                                                                      has_sparse ? 1 : 0, swin ? 1 : 0, ps.xq, swin_bias, window_size_x2,
                                                                      quant_size, dense_rel, sparse_rel, pos_win, totals);
    // tile bases as int32 (checked < 2^31 by plan_totals_kernel) + the sorted orders the kernels index with
    plan_export_kernel<<<blocks_for(N + 1), 256, 0, s>>>(N, st.rank_s, ps.tb, st.order_s, st.wstart_s, st.order_l, st.samp, n_samp,
                                                         has_sparse ? 1 : 0, tile_base, order_s, wstart_s, order_l, samp);
    return check_launch("fused_plan_fill");
}

}  // extern "C"
