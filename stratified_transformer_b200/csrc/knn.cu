// k nearest neighbours per scene (TransitionDown / Upsample support op).
//
// Result specification = the reference kernel, including the order in which equal distances come out:
//   /root/reference/lib/pointops2/src/knnquery/knnquery_cuda_kernel.cu:65-108 — one thread per query scans its
//   scene in index order, keeps the k best in a binary max-heap (candidate accepted iff d2 < heap root, sift-down that
//   prefers the right child only when strictly larger and stops when the parent is strictly larger), heap-sorts at
//   the end.  d2 is contracted as fma(dz,dz, fma(dx,dx, dy*dy)) (SASS of the reference build).
// The same heap discipline is kept (it defines the tie order), but the scan is restructured: a CTA of 128 queries
// streams its scene through shared memory in tiles (coalesced, each point read once per CTA instead of once per
// thread), the heap root lives in a register so the common "reject" path is 8 instructions with no memory access,
// and the heap arrays are touched only on the ~k ln(n/k) accepted candidates.
// stb200_knnquery_ws adds a grid-pruned search in front of it (the brute-force scan is O(m n): 8 ms for TransitionDown on one
// 80k-point scene): support points are sorted by (scene, cell) with a per-scene cell edge of about two point spacings; a query
// collects the k+1 nearest candidates from the 3 x 3 x 3 (then 5^3, 7^3, 9^3) cells around it and is DONE when they all lie
// closer than the searched cube's nearest face and their distances are strictly increasing - then the k nearest points and
// their order are unique, so the result equals the reference's whatever its heap would have done.  Any tie (equal distances
// among the k+1, e.g. lattice data), a scene with at most k points, or a query whose neighbourhood is still too sparse at 9^3
// cells goes on a list that the exact heap kernel above processes afterwards.
#include "common.cuh"

#include <cstdlib>

#include <cub/cub.cuh>

namespace stb200 {

constexpr int kKnnThreads = 128;
constexpr int kKnnTile = 1024;
constexpr int kKnnMaxK = 100;   // the reference's fixed heap capacity

__device__ __forceinline__ void knn_sift_down(float *dist, int *idx, int k) {
    int root = 0, child = 1;
    while (child < k) {
        if (child + 1 < k && dist[child + 1] > dist[child]) ++child;
        if (dist[root] > dist[child]) return;
        const float td = dist[root]; dist[root] = dist[child]; dist[child] = td;
        const int ti = idx[root]; idx[root] = idx[child]; idx[child] = ti;
        root = child;
        child = 2 * root + 1;
    }
}

__global__ void __launch_bounds__(kKnnThreads) knn_kernel(int m, int b, int k, const float *__restrict__ xyz,
                                                          const float *__restrict__ new_xyz, const int *__restrict__ offset,
                                                          const int *__restrict__ new_offset, int *__restrict__ idx,
                                                          float *__restrict__ dist2, const int *__restrict__ qlist = nullptr,
                                                          const int *__restrict__ qcount = nullptr) {
    __shared__ float4 tile[kKnnTile];   // (x, y, z, -) per point: one LDS.128 broadcast per candidate
    __shared__ int range[2];
    // qlist: only the listed queries (the ones the grid search could not settle), *qcount of them
    const int slot = blockIdx.x * kKnnThreads + threadIdx.x;
    const int n_q = qlist ? min(*qcount, m) : m;
    if (blockIdx.x * kKnnThreads >= n_q) return;   // whole CTA idle (uniform)
    const bool live = slot < n_q;
    const int q = live ? (qlist ? __ldg(qlist + slot) : slot) : 0;
    int start = 0, end = 0;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (live) {
        int s = 0;
        while (s < b - 1 && q >= __ldg(new_offset + s)) ++s;   // scene of this query
        start = s ? __ldg(offset + s - 1) : 0;
        end = __ldg(offset + s);
        qx = __ldg(new_xyz + (size_t)q * 3);
        qy = __ldg(new_xyz + (size_t)q * 3 + 1);
        qz = __ldg(new_xyz + (size_t)q * 3 + 2);
    }
    // point range needed by this CTA (its queries are consecutive, so they span one scene or a few adjacent ones)
    if (threadIdx.x == 0) { range[0] = 0x7fffffff; range[1] = 0; }
    __syncthreads();
    if (live) { atomicMin(&range[0], start); atomicMax(&range[1], end); }
    __syncthreads();
    const int lo = range[0], hi = range[1];

    float hd[kKnnMaxK];
    int hi_[kKnnMaxK];
    for (int i = 0; i < k; ++i) { hd[i] = 1e10f; hi_[i] = start; }
    float root = 1e10f;

    for (int t0 = lo; t0 < hi; t0 += kKnnTile) {
        const int tn = min(kKnnTile, hi - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < tn; i += kKnnThreads) {
            const float *src = xyz + (size_t)(t0 + i) * 3;
            tile[i] = make_float4(__ldg(src), __ldg(src + 1), __ldg(src + 2), 0.f);
        }
        __syncthreads();
        const int a = max(start, t0) - t0, e = min(end, t0 + tn) - t0;
#pragma unroll 4
        for (int i = a; i < e; ++i) {
            const float4 c = tile[i];
            const float dx = __fsub_rn(qx, c.x), dy = __fsub_rn(qy, c.y), dz = __fsub_rn(qz, c.z);
            const float d2 = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
            if (d2 < root) {
                hd[0] = d2;
                hi_[0] = t0 + i;
                knn_sift_down(hd, hi_, k);
                root = hd[0];
            }
        }
    }
    if (!live) return;
    for (int i = k - 1; i > 0; --i) {   // heap sort, ascending
        const float td = hd[0]; hd[0] = hd[i]; hd[i] = td;
        const int ti = hi_[0]; hi_[0] = hi_[i]; hi_[i] = ti;
        knn_sift_down(hd, hi_, i);
    }
    for (int i = 0; i < k; ++i) {
        idx[(size_t)q * k + i] = hi_[i];
        dist2[(size_t)q * k + i] = hd[i];
    }
}


// ---- grid-pruned search --------------------------------------------------------------------------------------------------
struct KnnBox { float mnx, mny, mnz, cell, inv_cell; int n, start, pad; };
constexpr int kKnnMaxScenes = 256;
constexpr int kKnnRowDim = 260;   // cells per axis are capped at 256 (+1 shift, +1 for the neighbour of the last cell): row table [260][260]

__global__ void knn_scene_box_kernel(int b, const float *__restrict__ xyz, const int *__restrict__ offset, KnnBox *__restrict__ box,
                                     float pts_per_cell) {
    __shared__ float red[6][32];
    const int s = blockIdx.x;
    const int start = s ? offset[s - 1] : 0, n = offset[s] - start;
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = threadIdx.x; i < n; i += blockDim.x)
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float v = __ldg(xyz + (size_t)(start + i) * 3 + a);
            mn[a] = fminf(mn[a], v);
            mx[a] = fmaxf(mx[a], v);
        }
    const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
        }
        if (lane == 0) { red[a][warp] = mn[a]; red[3 + a][warp] = mx[a]; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nw = blockDim.x / 32;
        for (int a = 0; a < 3; ++a)
            for (int w = 1; w < nw; ++w) { red[a][0] = fminf(red[a][0], red[a][w]); red[3 + a][0] = fmaxf(red[3 + a][0], red[3 + a][w]); }
        KnnBox bx;
        bx.mnx = red[0][0]; bx.mny = red[1][0]; bx.mnz = red[2][0];
        const float ex = fmaxf(red[3][0] - red[0][0], 1e-6f), ey = fmaxf(red[4][0] - red[1][0], 1e-6f), ez = fmaxf(red[5][0] - red[2][0], 1e-6f);
        // cell edge: the cube that would hold 2 points at uniform density (measured best of 8 / 2 / 0.5 on room scenes: dense clutter makes
        // large cells expensive for the queries next to it), but at least extent / 255; scans and indoor rooms are surfaces, the ring expansion covers what this under-estimates
        float cell = cbrtf(ex * ey * ez * pts_per_cell / (float)max(n, 1));
        cell = fmaxf(cell, fmaxf(ex, fmaxf(ey, ez)) / 255.f);   // at most 256 cells per axis: the (z, y) row table stays small
        bx.cell = cell; bx.inv_cell = 1.0f / cell; bx.n = n; bx.start = start; bx.pad = 0;
        box[s] = bx;
    }
}

__device__ __forceinline__ int knn_cell(float p, float mn, float inv_cell) {
    const float c = floorf((p - mn) * inv_cell);
    return (int)fminf(fmaxf(c, -1.f), 65533.f) + 1;
}
__device__ __forceinline__ unsigned long long knn_key(int scene, int cz, int cy, int cx) {
    return ((unsigned long long)scene << 48) | ((unsigned long long)cz << 32) | ((unsigned long long)cy << 16) | (unsigned long long)cx;
}
__device__ __forceinline__ int knn_scene_of(int i, int b, const int *__restrict__ ends) {
    int lo = 0, hi = b - 1;   // first scene whose end exceeds i
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(ends + mid) > i) hi = mid; else lo = mid + 1;
    }
    return lo;
}

__global__ void knn_keys_kernel(int N, int b, const float *__restrict__ xyz, const int *__restrict__ offset, const KnnBox *__restrict__ box,
                                unsigned long long *__restrict__ keys, int *__restrict__ vals) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        const int s = knn_scene_of(i, b, offset);
        const KnnBox bx = box[s];
        keys[i] = knn_key(s, knn_cell(__ldg(xyz + (size_t)i * 3 + 2), bx.mnz, bx.inv_cell), knn_cell(__ldg(xyz + (size_t)i * 3 + 1), bx.mny, bx.inv_cell),
                          knn_cell(__ldg(xyz + (size_t)i * 3), bx.mnx, bx.inv_cell));
        vals[i] = i;
    }
}

__global__ void knn_gather_kernel(int N, const float *__restrict__ xyz, const int *__restrict__ vals, float4 *__restrict__ pts) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < N; j += gridDim.x * blockDim.x) {
        const int i = __ldg(vals + j);
        pts[j] = make_float4(__ldg(xyz + (size_t)i * 3), __ldg(xyz + (size_t)i * 3 + 1), __ldg(xyz + (size_t)i * 3 + 2), __int_as_float(i));
    }
}

// first / one-past-last sorted position of every (scene, z, y) row of cells: a query finds its 9 rows with two loads each
// instead of a binary search over all keys (which cost 9 x 17 dependent global loads per query)
__global__ void knn_rows_kernel(int N, const unsigned long long *__restrict__ keys, int *__restrict__ row_start, int *__restrict__ row_end) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < N; j += gridDim.x * blockDim.x) {
        const unsigned long long r = __ldg(keys + j) >> 16;
        const int scene = (int)(r >> 32), z = (int)((r >> 16) & 0xffff), y = (int)(r & 0xffff);
        const size_t slot = ((size_t)scene * kKnnRowDim + min(z, kKnnRowDim - 1)) * kKnnRowDim + min(y, kKnnRowDim - 1);
        if (j == 0 || (__ldg(keys + j - 1) >> 16) != r) row_start[slot] = j;
        if (j == N - 1 || (__ldg(keys + j + 1) >> 16) != r) row_end[slot] = j + 1;
    }
}

constexpr int kKnnGridMaxK = 32;     // the grid search keeps k + 1 candidates in registers / local memory; larger k: heap kernel
constexpr int kKnnMaxRing = 4;

__global__ void __launch_bounds__(128)
knn_grid_kernel(int m, int b, int k, int N, const float *__restrict__ new_xyz, const int *__restrict__ new_offset,
                const KnnBox *__restrict__ box, const unsigned long long *__restrict__ keys, const float4 *__restrict__ pts,
                const int *__restrict__ row_start, const int *__restrict__ row_end, int *__restrict__ idx, float *__restrict__ dist2, int *__restrict__ qlist, int *__restrict__ qcount) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= m) return;
    const int s = knn_scene_of(q, b, new_offset);
    const KnnBox bx = box[s];
    const float qx = __ldg(new_xyz + (size_t)q * 3), qy = __ldg(new_xyz + (size_t)q * 3 + 1), qz = __ldg(new_xyz + (size_t)q * 3 + 2);
    bool settled = false;
    float bd[kKnnGridMaxK + 1];
    int bi[kKnnGridMaxK + 1];
    const int K1 = k + 1;
    if (bx.n > k) {
        const int cx = knn_cell(qx, bx.mnx, bx.inv_cell), cy = knn_cell(qy, bx.mny, bx.inv_cell), cz = knn_cell(qz, bx.mnz, bx.inv_cell);
        for (int r = 1; r <= kKnnMaxRing && !settled; ++r) {
            int cnt = 0;
            for (int dz = -r; dz <= r; ++dz)
                for (int dy = -r; dy <= r; ++dy) {
                    const int z = cz + dz, y = cy + dy;
                    if (z < 0 || y < 0 || z >= kKnnRowDim || y >= kKnnRowDim) continue;
                    const unsigned long long k_lo = knn_key(s, z, y, max(cx - r, 0)), k_hi = knn_key(s, z, y, min(cx + r, 65535));
                    const size_t slot = ((size_t)s * kKnnRowDim + z) * kKnnRowDim + y;
                    int lo = __ldg(row_start + slot), hi = __ldg(row_end + slot);
                    const int row_hi = hi;
                    while (hi - lo > 8) {   // lower bound of the x range inside the row (rows are short: a few probes)
                        const int mid = (lo + hi) >> 1;
                        if (__ldg(keys + mid) < k_lo) lo = mid + 1; else hi = mid;
                    }
                    for (int j = lo; j < row_hi && __ldg(keys + j) <= k_hi; ++j) {
                        if (__ldg(keys + j) < k_lo) continue;
                        const float4 c = __ldg(pts + j);
                        const float ex = __fsub_rn(qx, c.x), ey = __fsub_rn(qy, c.y), ez = __fsub_rn(qz, c.z);
                        const float d = __fmaf_rn(ez, ez, __fmaf_rn(ex, ex, __fmul_rn(ey, ey)));
                        if (cnt == K1 && !(d < bd[K1 - 1])) continue;
                        int pos = cnt < K1 ? cnt : K1 - 1;
                        while (pos > 0 && d < bd[pos - 1]) { bd[pos] = bd[pos - 1]; bi[pos] = bi[pos - 1]; --pos; }
                        bd[pos] = d;
                        bi[pos] = __float_as_int(c.w);
                        if (cnt < K1) ++cnt;
                    }
                }
            if (cnt < K1) continue;   // not even k + 1 candidates in this cube: widen
            // every point closer than the cube's nearest face is inside the cube (0.999: rounding of the cell coordinate)
            const float reach = 0.999f * (float)r * bx.cell;
            if (!(bd[K1 - 1] < reach * reach)) continue;
            bool strict = true;
            for (int i = 1; i < K1; ++i) strict = strict && (bd[i - 1] < bd[i]);
            if (!strict) break;        // a tie: the order is the heap's business
            settled = true;
        }
    }
    if (settled) {
        for (int i = 0; i < k; ++i) {
            idx[(size_t)q * k + i] = bi[i];
            dist2[(size_t)q * k + i] = bd[i];
        }
    } else {
        qlist[atomicAdd(qcount, 1)] = q;
    }
}

struct KnnScratch {
    KnnBox *box;
    unsigned long long *keys_in, *keys_out;
    int *vals_in, *vals_out, *qlist, *qcount, *row_start, *row_end;
    float4 *pts;
    void *cub_tmp;
    size_t cub_bytes, total;
};

static KnnScratch knn_layout(int N, int m, int b, void *base) {
    KnnScratch st{};
    char *p = (char *)base;
    size_t o = 0;
    auto take = [&](size_t bytes) { char *r = p ? p + o : nullptr; o += (bytes + 255) / 256 * 256; return r; };
    st.box = (KnnBox *)take((size_t)b * sizeof(KnnBox));
    st.keys_in = (unsigned long long *)take((size_t)N * 8);
    st.keys_out = (unsigned long long *)take((size_t)N * 8);
    st.vals_in = (int *)take((size_t)N * 4);
    st.vals_out = (int *)take((size_t)N * 4);
    st.pts = (float4 *)take((size_t)N * 16);
    st.qlist = (int *)take((size_t)m * 4);
    st.qcount = (int *)take(16);
    const size_t table = b <= kKnnMaxScenes ? (size_t)b * kKnnRowDim * kKnnRowDim * 4 : 0;   // more scenes: heap kernel, no tables
    st.row_start = (int *)take(table);
    st.row_end = (int *)take(table);
    st.cub_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, st.cub_bytes, (const unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                    (const int *)nullptr, (int *)nullptr, N, 0, 64);
    st.cub_tmp = take(st.cub_bytes);
    st.total = o;
    return st;
}

}  // namespace stb200

using namespace stb200;

extern "C" int stb200_knnquery(int m, int b, int nsample, const float *xyz, const float *new_xyz, const int *offset,
                               const int *new_offset, int *idx, float *dist2, void *stream) {
    STB200_REQUIRE(m >= 0 && b > 0 && nsample > 0 && nsample <= kKnnMaxK, STB200_ERR_ARG, "bad sizes (nsample <= %d)", kKnnMaxK);
    if (m == 0) return STB200_OK;
    STB200_REQUIRE(xyz && new_xyz && offset && new_offset && idx && dist2, STB200_ERR_ARG, "null pointer");
    {
        KernelScope ks("knnquery", 0.0, (cudaStream_t)stream);
        knn_kernel<<<(m + kKnnThreads - 1) / kKnnThreads, kKnnThreads, 0, (cudaStream_t)stream>>>(m, b, nsample, xyz, new_xyz, offset,
                                                                                            new_offset, idx, dist2);
    }
    return check_launch("knnquery");
}

extern "C" size_t stb200_knnquery_workspace_bytes(int n, int m, int b) { return n > 0 && m > 0 && b > 0 ? knn_layout(n, m, b, nullptr).total : 0; }

extern "C" int stb200_knnquery_ws(int n, int m, int b, int nsample, const float *xyz, const float *new_xyz, const int *offset,
                                  const int *new_offset, int *idx, float *dist2, void *workspace, size_t workspace_bytes, void *stream) {
    STB200_REQUIRE(n >= 0 && m >= 0 && b > 0 && nsample > 0 && nsample <= kKnnMaxK, STB200_ERR_ARG, "bad sizes (nsample <= %d)", kKnnMaxK);
    if (m == 0) return STB200_OK;
    STB200_REQUIRE(xyz && new_xyz && offset && new_offset && idx && dist2, STB200_ERR_ARG, "null pointer");
    if (nsample > kKnnGridMaxK || n == 0 || b > kKnnMaxScenes || !workspace)   // outside the grid search's range: the heap kernel alone
        return stb200_knnquery(m, b, nsample, xyz, new_xyz, offset, new_offset, idx, dist2, stream);
    KnnScratch st = knn_layout(n, m, b, workspace);
    STB200_REQUIRE(workspace_bytes >= st.total, STB200_ERR_WORKSPACE, "knnquery workspace: %zu B given, %zu B needed", workspace_bytes, st.total);
    cudaStream_t s = (cudaStream_t)stream;
    {
        KernelScope ks("knnquery_grid_build", 0.0, s);
        static const float pts_per_cell = getenv("STB200_KNN_CELL_POINTS") ? (float)atof(getenv("STB200_KNN_CELL_POINTS")) : 2.f;
        knn_scene_box_kernel<<<b, 256, 0, s>>>(b, xyz, offset, st.box, pts_per_cell);
        knn_keys_kernel<<<min((n + 255) / 256, kNumSMs * 8), 256, 0, s>>>(n, b, xyz, offset, st.box, st.keys_in, st.vals_in);
        size_t tb = st.cub_bytes;
        cudaError_t e = cub::DeviceRadixSort::SortPairs(st.cub_tmp, tb, st.keys_in, st.keys_out, st.vals_in, st.vals_out, n, 0, 64, s);
        STB200_REQUIRE(e == cudaSuccess, STB200_ERR_CUDA, "knnquery sort: %s", cudaGetErrorString(e));
        knn_gather_kernel<<<min((n + 255) / 256, kNumSMs * 8), 256, 0, s>>>(n, xyz, st.vals_out, st.pts);
        e = cudaMemsetAsync(st.qcount, 0, sizeof(int), s);
        STB200_REQUIRE(e == cudaSuccess, STB200_ERR_CUDA, "memset: %s", cudaGetErrorString(e));
        e = cudaMemsetAsync(st.row_start, 0, (size_t)((char *)st.row_end - (char *)st.row_start) * 2, s);   // the two tables are adjacent
        STB200_REQUIRE(e == cudaSuccess, STB200_ERR_CUDA, "memset: %s", cudaGetErrorString(e));
        knn_rows_kernel<<<min((n + 255) / 256, kNumSMs * 8), 256, 0, s>>>(n, st.keys_out, st.row_start, st.row_end);
    }
    {
        KernelScope ks("knnquery_grid", 0.0, s);
        knn_grid_kernel<<<(m + 127) / 128, 128, 0, s>>>(m, b, nsample, n, new_xyz, new_offset, st.box, st.keys_out, st.pts, st.row_start, st.row_end, idx,
                                                       dist2, st.qlist, st.qcount);
    }
    {   // the queries the grid could not settle (ties, sparse neighbourhoods, tiny scenes): exact heap scan; idle CTAs exit at once
        KernelScope ks("knnquery", 0.0, s);
        knn_kernel<<<(m + kKnnThreads - 1) / kKnnThreads, kKnnThreads, 0, s>>>(m, b, nsample, xyz, new_xyz, offset, new_offset, idx, dist2,
                                                                               st.qlist, st.qcount);
    }
    return check_launch("knnquery_ws");
}
