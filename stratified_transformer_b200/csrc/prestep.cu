// Host pre-step of the training loop on the device: the per-point scene id and the radius neighbour search that
// /root/reference/train.py:319-325 computes on the CPU before every step
//     batch        = torch.cat([torch.tensor([ii] * o) for ii, o in enumerate(offset_)], 0).long()
//     neighbor_idx = tp.ball_query(radius, max_num_neighbors, coord, coord, mode="partial_dense", batch_x=batch, batch_y=batch)[0]
// tp = torch_points_kernels (third party, not vendored; parity unpinned): its CPU path keeps the first `max_num` matches of a
// nanoflann radius search in kd-tree traversal order, which is unspecified.  Specification here (oracle/prestep_oracle.py):
// for every query the support points of the same scene with d^2 < r^2 (d^2 = (dx*dx + dy*dy) + dz*dz in fp32, no contraction),
// ordered by (d^2, index), the first `max_num` of them, -1 padded.  Whenever at most `max_num` points are in range - the
// normal case at radius 2.5 x voxel size - that is the same SET the reference gets.
//
// Search structure: support points sorted by (scene, cell z, cell y, cell x) with cells of 1.001 r (the margin covers the
// fp32 rounding of the cell coordinate); a query scans the 9 x-runs of its 27 neighbour cells, each found by one binary
// search in the sorted keys.  One thread per query, sorted insertion into a local list of max_num entries.
#include "common.cuh"

#include <cub/cub.cuh>

namespace stb200 {

constexpr int kBallMaxK = 64;

__global__ void batch_from_offset_kernel(int N, int b, const int *__restrict__ offset, long long *__restrict__ batch) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        int lo = 0, hi = b;   // first scene whose cumulative end exceeds i
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(offset + mid) > i) hi = mid; else lo = mid + 1;
        }
        batch[i] = lo;
    }
}

// order-preserving float <-> int map for atomicMin / atomicMax
__device__ __forceinline__ int f2ord(float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void ball_min_init_kernel(int *mn) { if (threadIdx.x < 3) mn[threadIdx.x] = 0x7fffffff; }

__global__ void ball_min_kernel(int N, const float *__restrict__ xyz, int *mn) {
    float m[3] = {INFINITY, INFINITY, INFINITY};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x)
#pragma unroll
        for (int a = 0; a < 3; ++a) m[a] = fminf(m[a], __ldg(xyz + (size_t)i * 3 + a));
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o; o >>= 1) m[a] = fminf(m[a], __shfl_xor_sync(0xffffffffu, m[a], o));
        if ((threadIdx.x & 31) == 0 && m[a] < INFINITY) atomicMin(mn + a, f2ord(m[a]));
    }
}

// cell coordinate of a point, clamped to [0, 65534] after a +1 shift so that queries just outside the support box still map
__device__ __forceinline__ int ball_cell(float p, float mn, float inv_cell) {
    const float c = floorf(__fmul_rn(__fsub_rn(p, mn), inv_cell));
    return (int)fminf(fmaxf(c, -1.f), 65533.f) + 1;
}
__device__ __forceinline__ unsigned long long ball_key(long long scene, int cz, int cy, int cx) {
    return ((unsigned long long)scene << 48) | ((unsigned long long)cz << 32) | ((unsigned long long)cy << 16) | (unsigned long long)cx;
}

__global__ void ball_keys_kernel(int N, const float *__restrict__ xyz, const long long *__restrict__ batch, const int *__restrict__ mn,
                                 float inv_cell, unsigned long long *__restrict__ keys, int *__restrict__ vals) {
    const float m0 = ord2f(mn[0]), m1 = ord2f(mn[1]), m2 = ord2f(mn[2]);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        const float x = __ldg(xyz + (size_t)i * 3), y = __ldg(xyz + (size_t)i * 3 + 1), z = __ldg(xyz + (size_t)i * 3 + 2);
        keys[i] = ball_key(batch ? batch[i] : 0, ball_cell(z, m2, inv_cell), ball_cell(y, m1, inv_cell), ball_cell(x, m0, inv_cell));
        vals[i] = i;
    }
}

__global__ void __launch_bounds__(128)
ball_query_kernel(int Nx, int Ny, int K, const float *__restrict__ x, const float *__restrict__ y, const long long *__restrict__ batch_y,
                  const int *__restrict__ mn, float inv_cell, float r2, const unsigned long long *__restrict__ keys,
                  const int *__restrict__ vals, long long *__restrict__ idx, float *__restrict__ dist2) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Ny) return;
    const float m0 = ord2f(mn[0]), m1 = ord2f(mn[1]), m2 = ord2f(mn[2]);
    const float qx = __ldg(y + (size_t)q * 3), qy = __ldg(y + (size_t)q * 3 + 1), qz = __ldg(y + (size_t)q * 3 + 2);
    const long long scene = batch_y ? batch_y[q] : 0;
    const int cx = ball_cell(qx, m0, inv_cell), cy = ball_cell(qy, m1, inv_cell), cz = ball_cell(qz, m2, inv_cell);
    float bd[kBallMaxK];
    int bi[kBallMaxK];
    int cnt = 0;
    for (int dz = -1; dz <= 1; ++dz)
        for (int dy = -1; dy <= 1; ++dy) {
            const int z = cz + dz, yy = cy + dy;
            if (z < 0 || yy < 0 || z > 65535 || yy > 65535) continue;
            const unsigned long long k_lo = ball_key(scene, z, yy, max(cx - 1, 0)), k_hi = ball_key(scene, z, yy, min(cx + 1, 65535));
            int lo = 0, hi = Nx;   // first sorted position with key >= k_lo
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (__ldg(keys + mid) < k_lo) lo = mid + 1; else hi = mid;
            }
            for (int j = lo; j < Nx && __ldg(keys + j) <= k_hi; ++j) {
                const int p = __ldg(vals + j);
                const float ex = __fsub_rn(__ldg(x + (size_t)p * 3), qx), ey = __fsub_rn(__ldg(x + (size_t)p * 3 + 1), qy),
                            ez = __fsub_rn(__ldg(x + (size_t)p * 3 + 2), qz);
                const float d = __fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez));
                if (!(d < r2)) continue;
                if (cnt == K && !(d < bd[K - 1] || (d == bd[K - 1] && p < bi[K - 1]))) continue;
                int pos = cnt < K ? cnt : K - 1;          // sorted insertion by (d, index)
                while (pos > 0 && (d < bd[pos - 1] || (d == bd[pos - 1] && p < bi[pos - 1]))) {
                    bd[pos] = bd[pos - 1];
                    bi[pos] = bi[pos - 1];
                    --pos;
                }
                bd[pos] = d;
                bi[pos] = p;
                if (cnt < K) ++cnt;
            }
        }
    for (int s = 0; s < K; ++s) {
        idx[(size_t)q * K + s] = s < cnt ? bi[s] : -1;
        if (dist2) dist2[(size_t)q * K + s] = s < cnt ? bd[s] : -1.f;
    }
}

struct BallScratch {
    unsigned long long *keys_in, *keys_out;
    int *vals_in, *vals_out, *mn;
    void *cub_tmp;
    size_t cub_bytes, total;
};

static BallScratch ball_layout(int Nx, void *base) {
    BallScratch st{};
    char *p = (char *)base;
    size_t o = 0;
    auto take = [&](size_t bytes) { char *r = p ? p + o : nullptr; o += (bytes + 255) / 256 * 256; return r; };
    st.keys_in = (unsigned long long *)take((size_t)Nx * 8);
    st.keys_out = (unsigned long long *)take((size_t)Nx * 8);
    st.vals_in = (int *)take((size_t)Nx * 4);
    st.vals_out = (int *)take((size_t)Nx * 4);
    st.mn = (int *)take(16);
    st.cub_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, st.cub_bytes, (const unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                    (const int *)nullptr, (int *)nullptr, Nx, 0, 64);
    st.cub_tmp = take(st.cub_bytes);
    st.total = o;
    return st;
}

}  // namespace stb200

using namespace stb200;

extern "C" {

int stb200_batch_from_offset(int N, int b, const int *offset, long long *batch, void *stream) {
    STB200_REQUIRE(N >= 0 && b >= 0, STB200_ERR_ARG, "bad sizes");
    if (N == 0) return STB200_OK;
    STB200_REQUIRE(offset && batch && b > 0, STB200_ERR_ARG, "null pointer / no scenes");
    KernelScope ks("batch_from_offset", 8.0 * N, (cudaStream_t)stream);
    batch_from_offset_kernel<<<min((N + 255) / 256, kNumSMs * 8), 256, 0, (cudaStream_t)stream>>>(N, b, offset, batch);
    return check_launch("batch_from_offset");
}

size_t stb200_ball_query_workspace_bytes(int Nx) { return Nx > 0 ? ball_layout(Nx, nullptr).total : 0; }

int stb200_ball_query(int Nx, int Ny, float radius, int max_num, const float *x, const float *y, const long long *batch_x,
                      const long long *batch_y, void *workspace, size_t workspace_bytes, long long *idx, float *dist2, void *stream) {
    STB200_REQUIRE(Nx >= 0 && Ny >= 0 && max_num > 0 && max_num <= kBallMaxK && radius > 0.f, STB200_ERR_ARG,
                   "ball_query: bad sizes (max_num <= %d, radius > 0)", kBallMaxK);
    if (Ny == 0) return STB200_OK;
    STB200_REQUIRE(y && idx, STB200_ERR_ARG, "null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    if (Nx == 0) {   // nothing to find: all -1
        cudaError_t e = cudaMemsetAsync(idx, 0xff, (size_t)Ny * max_num * sizeof(long long), s);
        STB200_REQUIRE(e == cudaSuccess, STB200_ERR_CUDA, "memset: %s", cudaGetErrorString(e));
        return STB200_OK;
    }
    STB200_REQUIRE(x && workspace, STB200_ERR_ARG, "null pointer");
    STB200_REQUIRE((batch_x == nullptr) == (batch_y == nullptr), STB200_ERR_ARG, "batch_x and batch_y: both or neither");
    BallScratch st = ball_layout(Nx, workspace);
    STB200_REQUIRE(workspace_bytes >= st.total, STB200_ERR_WORKSPACE, "ball_query workspace: %zu B given, %zu B needed", workspace_bytes, st.total);
    const float cell = 1.001f * radius, inv_cell = 1.0f / cell, r2 = radius * radius;
    {
        KernelScope ks("ball_query_sort", 0.0, s);
        ball_min_init_kernel<<<1, 32, 0, s>>>(st.mn);
        ball_min_kernel<<<min((Nx + 255) / 256, kNumSMs * 4), 256, 0, s>>>(Nx, x, st.mn);
        ball_keys_kernel<<<min((Nx + 255) / 256, kNumSMs * 8), 256, 0, s>>>(Nx, x, batch_x, st.mn, inv_cell, st.keys_in, st.vals_in);
        size_t tb = st.cub_bytes;
        const cudaError_t e = cub::DeviceRadixSort::SortPairs(st.cub_tmp, tb, st.keys_in, st.keys_out, st.vals_in, st.vals_out, Nx, 0, 64, s);
        STB200_REQUIRE(e == cudaSuccess, STB200_ERR_CUDA, "ball_query sort: %s", cudaGetErrorString(e));
    }
    {
        KernelScope ks("ball_query", 12.0 * Nx + (double)Ny * (12.0 + 12.0 * max_num), s);
        ball_query_kernel<<<(Ny + 127) / 128, 128, 0, s>>>(Nx, Ny, max_num, x, y, batch_y, st.mn, inv_cell, r2, st.keys_out, st.vals_out, idx, dist2);
    }
    return check_launch("ball_query");
}

}  // extern "C"
