// Index-side kernels: transposed CSR (pairs grouped by key).
//
// The reference has no such structure: it scatters gradients to keys with float atomicAdd
// (/root/reference/lib/pointops2/src/attention_v2/attention_cuda_kernel_v2.cu:84,
//  rpe_v2/relative_pos_encoding_cuda_kernel_v2.cu:326,477).  Grouping the pairs by key once per index set
// turns those scatters into deterministic segment gathers (seg_reduce<PERM>).
#include <cub/cub.cuh>

#include "common.cuh"

namespace stb200 {

__global__ void expand_index0_kernel(int N, const int *__restrict__ offsets, int *__restrict__ index0,
                                     int *__restrict__ iota) {
    const int lane = threadIdx.x % kWarp;
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) / kWarp, nw = gridDim.x * blockDim.x / kWarp;
    for (int n = wid; n < N; n += nw) {
        const int s = __ldg(offsets + n), e = __ldg(offsets + n + 1);
        for (int m = s + lane; m < e; m += kWarp) {
            index0[m] = n;
            iota[m] = m;
        }
    }
}

// t_offsets[j] = first position in the key-sorted list whose key >= j
__global__ void key_offsets_kernel(int N, int M, const int *__restrict__ sorted_keys, int *__restrict__ t_offsets) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j > N) return;
    int lo = 0, hi = M;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(sorted_keys + mid) < j) lo = mid + 1; else hi = mid;
    }
    t_offsets[j] = lo;
}

__global__ void gather_int_kernel(int M, const int *__restrict__ src, const int *__restrict__ idx, int *__restrict__ dst) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) dst[i] = __ldg(src + __ldg(idx + i));
}

__global__ void pack_rel_kernel(int M, int L, const int *__restrict__ rel_idx, const int *__restrict__ perm,
                                unsigned *__restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
        const size_t m = perm ? __ldg(perm + i) : i;
        const unsigned r0 = min(max(__ldg(rel_idx + 3 * m + 0), 0), L - 1), r1 = min(max(__ldg(rel_idx + 3 * m + 1), 0), L - 1),
                       r2 = min(max(__ldg(rel_idx + 3 * m + 2), 0), L - 1);
        out[i] = r0 | (r1 << 10) | (r2 << 20);
    }
}

constexpr int kLengthKeyBits = 16, kLengthKeyMax = (1 << kLengthKeyBits) - 1;   // two radix passes

// keys[i] = pair count of row r_i (clamped), vals[i] = r_i, with r_i = base_order ? base_order[i] : i
__global__ void row_length_keys_kernel(int N, const int *__restrict__ offsets, const int *__restrict__ base_order,
                                       int *__restrict__ keys, int *__restrict__ vals) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        const int r = base_order ? __ldg(base_order + i) : i;
        keys[i] = min(__ldg(offsets + r + 1) - __ldg(offsets + r), kLengthKeyMax);   // a balance hint: longer rows may tie
        vals[i] = r;
    }
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static size_t length_sort_temp_bytes(int N) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const int *)nullptr, (int *)nullptr, (const int *)nullptr,
                                    (int *)nullptr, N, 0, kLengthKeyBits);
    return bytes;
}

static int key_bits(int N) {
    int b = 1;
    while ((1LL << b) < (long long)N) ++b;
    return b;
}

static size_t sort_temp_bytes(int N, int M) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const int *)nullptr, (int *)nullptr, (const int *)nullptr,
                                    (int *)nullptr, M, 0, key_bits(N));
    return bytes;
}

}  // namespace stb200

using namespace stb200;

extern "C" {

int stb200_pack_rel(int M, int L, const int *rel_idx, const int *perm, unsigned *out, void *stream) {
    STB200_REQUIRE(M >= 0 && L > 0 && L <= 1024, STB200_ERR_ARG, "bad M / L (L <= 1024)");
    if (M == 0) return STB200_OK;
    STB200_REQUIRE(rel_idx && out, STB200_ERR_ARG, "null pointer");
    {
        KernelScope ks("pack_rel", 16.0 * M + (perm ? 4.0 * M : 0.0), (cudaStream_t)stream);
        pack_rel_kernel<<<max(1, min((M + 255) / 256, kNumSMs * 8)), 256, 0, (cudaStream_t)stream>>>(M, L, rel_idx, perm, out);
    }
    return check_launch("pack_rel");
}

size_t stb200_length_order_workspace_bytes(int N) {
    if (N <= 0) return 256;
    return 3 * align256((size_t)N * sizeof(int)) + align256(length_sort_temp_bytes(N)) + 256;
}

int stb200_length_order(int N, const int *offsets, const int *base_order, int *order, void *workspace,
                        size_t workspace_bytes, void *stream) {
    STB200_REQUIRE(N >= 0, STB200_ERR_ARG, "bad size");
    if (N == 0) return STB200_OK;
    STB200_REQUIRE(offsets && order && workspace, STB200_ERR_ARG, "null pointer");
    STB200_REQUIRE(workspace_bytes >= stb200_length_order_workspace_bytes(N), STB200_ERR_WORKSPACE,
                   "workspace too small: %zu < %zu", workspace_bytes, stb200_length_order_workspace_bytes(N));
    cudaStream_t s = (cudaStream_t)stream;
    char *ws = (char *)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    const size_t ni = align256((size_t)N * sizeof(int));
    int *keys = (int *)ws, *vals = (int *)(ws + ni), *keys_out = (int *)(ws + 2 * ni);
    void *tmp = ws + 3 * ni;
    size_t tmp_bytes = length_sort_temp_bytes(N);
    KernelScope ks("length_order", 4.0 * (N + 1) + 4.0 * N * (base_order ? 2 : 1), s);
    count_launch(1);
    row_length_keys_kernel<<<max(1, min((N + 255) / 256, kNumSMs * 8)), 256, 0, s>>>(N, offsets, base_order, keys, vals);
    // stable: rows of equal length stay in base order (window order when the pair builder's row_order is passed)
    cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys_out, vals, order, N, 0, kLengthKeyBits, s);
    if (e != cudaSuccess) {
        set_error("cub radix sort: %s", cudaGetErrorString(e));
        return STB200_ERR_CUDA;
    }
    return check_launch("length_order");
}

size_t stb200_transpose_csr_workspace_bytes(int N, int M) {
    if (M <= 0) return 256;
    return 3 * align256((size_t)M * sizeof(int)) + align256(sort_temp_bytes(N, M)) + 256;
}

int stb200_transpose_csr(int N, int M, const int *index0_offsets, const int *index1, int *t_offsets, int *t_pair,
                         int *t_index0, void *workspace, size_t workspace_bytes, void *stream) {
    STB200_REQUIRE(N >= 0 && M >= 0, STB200_ERR_ARG, "bad sizes");
    STB200_REQUIRE(t_offsets, STB200_ERR_ARG, "null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    if (M == 0) {
        cudaMemsetAsync(t_offsets, 0, (size_t)(N + 1) * sizeof(int), s);
        return check_launch("transpose_csr memset");
    }
    STB200_REQUIRE(index0_offsets && index1 && t_pair && t_index0 && workspace, STB200_ERR_ARG, "null pointer");
    STB200_REQUIRE(workspace_bytes >= stb200_transpose_csr_workspace_bytes(N, M), STB200_ERR_WORKSPACE,
                   "workspace too small: %zu < %zu", workspace_bytes, stb200_transpose_csr_workspace_bytes(N, M));
    char *ws = (char *)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    const size_t mi = align256((size_t)M * sizeof(int));
    int *index0 = (int *)ws;
    int *iota = (int *)(ws + mi);
    int *sorted_keys = (int *)(ws + 2 * mi);
    void *tmp = ws + 3 * mi;
    size_t tmp_bytes = sort_temp_bytes(N, M);

    const int blocks = max(1, min((N + 7) / 8, kNumSMs * 8));
    KernelScope ks("transpose_csr[6 launches]", 4.0 * (N + 1) * 2 + 4.0 * M * 3, s);
    count_launch(5);
    expand_index0_kernel<<<blocks, 256, 0, s>>>(N, index0_offsets, index0, iota);
    // stable LSD radix sort over the significant key bits only: within a key, pairs stay in ascending pair id
    cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, index1, sorted_keys, iota, t_pair, M, 0,
                                                    key_bits(N), s);
    if (e != cudaSuccess) {
        set_error("cub radix sort: %s", cudaGetErrorString(e));
        return STB200_ERR_CUDA;
    }
    key_offsets_kernel<<<(N + 1 + 255) / 256, 256, 0, s>>>(N, M, sorted_keys, t_offsets);
    gather_int_kernel<<<max(1, min((M + 255) / 256, kNumSMs * 8)), 256, 0, s>>>(M, index0, t_pair, t_index0);
    return check_launch("transpose_csr");
}

}  // extern "C"
