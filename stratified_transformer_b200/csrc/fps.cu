// Exact furthest point sampling, one thread-block cluster per scene.
//
// Result specification = the reference kernel, bit for bit, including its tie behaviour
// (/root/reference/lib/pointops2/src/sampling/sampling_cuda_kernel.cu:14-129):
//   d = fma(dz,dz, fma(dx,dx, dy*dy))      (the contraction nvcc emits for its source expression)
//   running minimum per point, winner = maximum; among equal maxima the reference's halving tree over a
//   block of B = opt_n_threads(n) threads keeps the candidate whose owner thread has the smallest
//   bit-reversed id, and inside a thread the first (lowest index) strict maximum.
// That order is encoded as  rank(i) = bitrev_log2B(i mod B) << 20 | i / B  and every reduction below is
// "max distance, then min rank", so the parallel decomposition can be anything.
//
// The reference runs ONE CTA per scene, re-reads xyz and the min-distance array from global memory every
// iteration and does a 10-barrier shared-memory tree.  Here a cluster of up to 16 CTAs owns a scene:
// coordinates and running minima live in registers for the whole run, the per-iteration argmax is
// redux.sync in the warp -> one __syncthreads in the CTA -> one DSMEM exchange + cluster barrier, and the
// winner's coordinates travel with the candidate record so no global load sits on the critical path.
#include <cooperative_groups.h>
#include <cmath>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace stb200 {

constexpr int kFpsThreads = 1024;
constexpr int kMaxCluster = 16;

struct __align__(16) FpsRec {
    unsigned key, rank;
    float x, y, z;
    unsigned pad[3];
};

__device__ __forceinline__ unsigned fps_rank(int i, int logB) {
    const unsigned t = (unsigned)i & ((1u << logB) - 1u);
    const unsigned rev = logB ? (__brev(t) >> (32 - logB)) : 0u;
    return (rev << 20) | ((unsigned)i >> logB);
}
__device__ __forceinline__ int fps_unrank(unsigned rank, int logB) {
    const unsigned rev = rank >> 20;
    const unsigned t = logB ? (__brev(rev) >> (32 - logB)) : 0u;
    return (int)(t + ((rank & 0xFFFFFu) << logB));
}

// (key, rank) -> warp-wide winner: max key, then min rank
__device__ __forceinline__ void warp_argmax(unsigned &key, unsigned &rank) {
    const unsigned kmax = __reduce_max_sync(0xffffffffu, key);
    const unsigned r = key == kmax ? rank : 0xffffffffu;
    rank = __reduce_min_sync(0xffffffffu, r);
    key = kmax;
}

template <int P, bool CLUSTER>
__global__ void __launch_bounds__(kFpsThreads, 1)
fps_kernel(const float *__restrict__ xyz, const int *__restrict__ offset, const int *__restrict__ new_offset,
           int *__restrict__ idx, int logB, int cluster_size) {
    extern __shared__ float sxyz[];  // [3][P * T] coordinates of this CTA's points, for winner look-up
    __shared__ unsigned wkey[2][32], wrank[2][32];
    __shared__ FpsRec crec[2][kMaxCluster];

    const int T = blockDim.x, tid = threadIdx.x, lane = tid % kWarp, warp = tid / kWarp, nwarps = T / kWarp;
    const int crank = CLUSTER ? (int)cg::this_cluster().block_rank() : 0;
    const int scene = blockIdx.x / cluster_size;
    const int TT = T * cluster_size, gtid = crank * T + tid;

    const int start_n = scene ? offset[scene - 1] : 0, n = offset[scene] - start_n;
    const int start_m = scene ? new_offset[scene - 1] : 0, m = new_offset[scene] - start_m;

    float px[P], py[P], pz[P], mind[P];
#pragma unroll
    for (int u = 0; u < P; ++u) {
        const int i = gtid + u * TT;
        const bool valid = i < n;
        px[u] = valid ? xyz[(size_t)(start_n + i) * 3 + 0] : 0.f;
        py[u] = valid ? xyz[(size_t)(start_n + i) * 3 + 1] : 0.f;
        pz[u] = valid ? xyz[(size_t)(start_n + i) * 3 + 2] : 0.f;
        mind[u] = valid ? 1e10f : -1.f;
        sxyz[0 * P * T + u * T + tid] = px[u];
        sxyz[1 * P * T + u * T + tid] = py[u];
        sxyz[2 * P * T + u * T + tid] = pz[u];
    }
    float ox = 0.f, oy = 0.f, oz = 0.f;
    if (n > 0) {
        ox = xyz[(size_t)start_n * 3 + 0];
        oy = xyz[(size_t)start_n * 3 + 1];
        oz = xyz[(size_t)start_n * 3 + 2];
    }
    if (gtid == 0 && m > 0) idx[start_m] = start_n;
    __syncthreads();
    if (CLUSTER) cg::this_cluster().sync();  // every CTA of the cluster is resident before remote stores

    int buf = 0;
    for (int j = 1; j < m; ++j, buf ^= 1) {
        float best = -2.f;
        int bu = 0;
#pragma unroll
        for (int u = 0; u < P; ++u) {
            const float dx = __fsub_rn(px[u], ox), dy = __fsub_rn(py[u], oy), dz = __fsub_rn(pz[u], oz);
            const float d = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
            const float d2 = fminf(d, mind[u]);
            mind[u] = d2;
            if (d2 > best) {
                best = d2;
                bu = u;
            }
        }
        // distances are >= 0: their bit patterns order like unsigned ints; 0 is reserved for "no point"
        unsigned key = best >= 0.f ? __float_as_uint(best) + 1u : 0u;
        unsigned rank = fps_rank(gtid + bu * TT, logB);
        warp_argmax(key, rank);
        if (lane == 0) {
            wkey[buf][warp] = key;
            wrank[buf][warp] = rank;
        }
        __syncthreads();
        key = lane < nwarps ? wkey[buf][lane] : 0u;
        rank = lane < nwarps ? wrank[buf][lane] : 0xffffffffu;
        warp_argmax(key, rank);  // every warp now knows the CTA winner

        int win;
        if (CLUSTER) {
            if (tid < cluster_size) {
                FpsRec r;
                r.key = key;
                r.rank = rank;
                const int i = fps_unrank(rank, logB);
                const int slot = ((i - crank * T) / TT) * T + (i - crank * T) % TT;  // u * T + tid of the owner
                const bool mine = key != 0u && i >= crank * T && (i - crank * T) % TT < T;
                r.x = mine ? sxyz[0 * P * T + slot] : 0.f;
                r.y = mine ? sxyz[1 * P * T + slot] : 0.f;
                r.z = mine ? sxyz[2 * P * T + slot] : 0.f;
                r.pad[0] = r.pad[1] = r.pad[2] = 0u;
                FpsRec *dst = cg::this_cluster().map_shared_rank(&crec[buf][crank], tid);
                *dst = r;
            }
            cg::this_cluster().sync();
            const FpsRec r = crec[buf][lane < cluster_size ? lane : 0];
            unsigned k2 = lane < cluster_size ? r.key : 0u;
            unsigned r2 = lane < cluster_size ? r.rank : 0xffffffffu;
            const unsigned myk = k2, myr = r2;
            warp_argmax(k2, r2);
            const unsigned who = __ballot_sync(0xffffffffu, lane < cluster_size && myk == k2 && myr == r2);
            const int src = __ffs(who) - 1;
            ox = __shfl_sync(0xffffffffu, r.x, src);
            oy = __shfl_sync(0xffffffffu, r.y, src);
            oz = __shfl_sync(0xffffffffu, r.z, src);
            win = fps_unrank(r2, logB);
        } else {
            win = fps_unrank(rank, logB);
            const int slot = (win / T) * T + win % T;
            ox = sxyz[0 * P * T + slot];
            oy = sxyz[1 * P * T + slot];
            oz = sxyz[2 * P * T + slot];
        }
        if (gtid == 0) idx[start_m + j] = start_n + win;
    }
    if (CLUSTER) cg::this_cluster().sync();  // no CTA exits while a peer may still write into its smem
}

// Fallback for scenes too large for the register-resident kernel: one CTA per scene, coordinates and running
// minima streamed from global memory (the caller's `tmp` scratch), same (max distance, min rank) reduction.
__global__ void __launch_bounds__(kFpsThreads, 1)
fps_streaming_kernel(const float *__restrict__ xyz, const int *__restrict__ offset, const int *__restrict__ new_offset,
                     float *__restrict__ tmp, int *__restrict__ idx, int logB) {
    __shared__ unsigned wkey[2][32], wrank[2][32];
    const int T = blockDim.x, tid = threadIdx.x, lane = tid % kWarp, warp = tid / kWarp, nwarps = T / kWarp;
    const int scene = blockIdx.x;
    const int start_n = scene ? offset[scene - 1] : 0, n = offset[scene] - start_n;
    const int start_m = scene ? new_offset[scene - 1] : 0, m = new_offset[scene] - start_m;
    for (int i = tid; i < n; i += T) tmp[start_n + i] = 1e10f;
    if (tid == 0 && m > 0) idx[start_m] = start_n;
    int old = 0, buf = 0;
    for (int j = 1; j < m; ++j, buf ^= 1) {
        const float ox = xyz[(size_t)(start_n + old) * 3], oy = xyz[(size_t)(start_n + old) * 3 + 1],
                    oz = xyz[(size_t)(start_n + old) * 3 + 2];
        float best = -2.f;
        int bi = 0;
        for (int i = tid; i < n; i += T) {
            const float dx = __fsub_rn(xyz[(size_t)(start_n + i) * 3], ox);
            const float dy = __fsub_rn(xyz[(size_t)(start_n + i) * 3 + 1], oy);
            const float dz = __fsub_rn(xyz[(size_t)(start_n + i) * 3 + 2], oz);
            const float d2 = fminf(__fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy))), tmp[start_n + i]);
            tmp[start_n + i] = d2;
            if (d2 > best) {
                best = d2;
                bi = i;
            }
        }
        // T is a multiple of B, so all of a thread's points share the bit-reversed part of the rank:
        // "first strict maximum" inside the thread is exactly "min rank"
        unsigned key = best >= 0.f ? __float_as_uint(best) + 1u : 0u;
        unsigned rank = fps_rank(bi, logB);
        warp_argmax(key, rank);
        if (lane == 0) {
            wkey[buf][warp] = key;
            wrank[buf][warp] = rank;
        }
        __syncthreads();
        key = lane < nwarps ? wkey[buf][lane] : 0u;
        rank = lane < nwarps ? wrank[buf][lane] : 0xffffffffu;
        warp_argmax(key, rank);
        old = fps_unrank(rank, logB);
        if (tid == 0) idx[start_m + j] = start_n + old;
    }
}

// the reference's opt_n_threads (cuda_utils.h:10-13), evaluated in double exactly like it
static int ref_block_log2(int n) {
    if (n < 1) n = 1;
    int pow_2 = (int)(std::log((double)n) / std::log(2.0));
    if (pow_2 > 10) pow_2 = 10;
    if (pow_2 < 0) pow_2 = 0;
    return pow_2;
}

template <int P>
static int launch_fps(int b, int cs, const float *xyz, const int *offset, const int *new_offset, int *idx, int logB,
                      cudaStream_t s) {
    const size_t smem = (size_t)3 * P * kFpsThreads * sizeof(float);
    cudaError_t e;
    KernelScope ks("fps_cluster", 0.0, s);  // latency-bound by construction: bytes are not the meaningful unit
    if (cs == 1) {
        auto kern = fps_kernel<P, false>;
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) {
            set_error("fps smem attr: %s", cudaGetErrorString(e));
            return STB200_ERR_CUDA;
        }
        kern<<<b, kFpsThreads, smem, s>>>(xyz, offset, new_offset, idx, logB, 1);
    } else {
        auto kern = fps_kernel<P, true>;
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess ||
            (e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1)) != cudaSuccess) {
            set_error("fps attr: %s", cudaGetErrorString(e));
            return STB200_ERR_CUDA;
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(b * cs);
        cfg.blockDim = dim3(kFpsThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = cs;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        if ((e = cudaLaunchKernelEx(&cfg, kern, xyz, offset, new_offset, idx, logB, cs)) != cudaSuccess) {
            set_error("fps cluster launch (cluster=%d): %s", cs, cudaGetErrorString(e));
            return STB200_ERR_CUDA;
        }
    }
    return check_launch("fps");
}

}  // namespace stb200

using namespace stb200;

extern "C" int stb200_furthestsampling(int b, int n, const float *xyz, const int *offset, const int *new_offset,
                                       float *tmp, int *idx, void *stream) {
    STB200_REQUIRE(b >= 0 && n >= 0, STB200_ERR_ARG, "bad sizes b=%d n=%d", b, n);
    if (b == 0 || n == 0) return STB200_OK;
    STB200_REQUIRE(xyz && offset && new_offset && idx, STB200_ERR_ARG, "null pointer");
    STB200_REQUIRE(n < (1 << 30), STB200_ERR_ARG, "scene too large");
    cudaStream_t s = (cudaStream_t)stream;
    const int logB = ref_block_log2(n);
    // cluster size: as many CTAs per scene as keep all scenes co-resident (148 SMs), but never more threads than points
    int cs = kMaxCluster;
    while (cs > 1 && (b * cs > kNumSMs || (cs / 2) * kFpsThreads >= n)) cs >>= 1;
    const int need = (n + cs * kFpsThreads - 1) / (cs * kFpsThreads);
    if (need <= 1) return launch_fps<1>(b, cs, xyz, offset, new_offset, idx, logB, s);
    if (need <= 2) return launch_fps<2>(b, cs, xyz, offset, new_offset, idx, logB, s);
    if (need <= 3) return launch_fps<3>(b, cs, xyz, offset, new_offset, idx, logB, s);
    if (need <= 4) return launch_fps<4>(b, cs, xyz, offset, new_offset, idx, logB, s);
    if (need <= 5) return launch_fps<5>(b, cs, xyz, offset, new_offset, idx, logB, s);
    if (need <= 6) return launch_fps<6>(b, cs, xyz, offset, new_offset, idx, logB, s);
    if (need <= 8) return launch_fps<8>(b, cs, xyz, offset, new_offset, idx, logB, s);
    STB200_REQUIRE(tmp, STB200_ERR_ARG, "scene of %d points needs the tmp scratch (streaming path)", n);
    {
        KernelScope ks("fps_streaming", 0.0, s);
        fps_streaming_kernel<<<b, kFpsThreads, 0, s>>>(xyz, offset, new_offset, tmp, idx, logB);
    }
    return check_launch("fps_streaming");
}
