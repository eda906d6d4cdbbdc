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
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace stb200 {

constexpr int kFpsThreads = 256;   // few warps per CTA: the per-iteration exchange cost grows with the warp count
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

// ---- mbarrier / DSMEM primitives (PTX) ---------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ unsigned map_to_cta(unsigned local_addr, unsigned cta) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta));
    return r;
}
// remote shared-memory store that also completes `bytes` on the destination CTA's mbarrier
__device__ __forceinline__ void st_async_v4(unsigned raddr, unsigned a, unsigned b, unsigned c, unsigned d, unsigned rbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(raddr),
                 "r"(a), "r"(b), "r"(c), "r"(d), "r"(rbar)
                 : "memory");
}
__device__ __forceinline__ void st_async_b32(unsigned raddr, unsigned a, unsigned rbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(raddr), "r"(a), "r"(rbar)
                 : "memory");
}

// (key, rank) -> warp winner; returns the winning lane.  Exact ties on the distance are rare (continuous
// coordinates), so the rank reduction only runs when the ballot shows more than one lane at the maximum.
__device__ __forceinline__ int warp_argmax_lane(unsigned key, unsigned rank) {
    const unsigned kmax = __reduce_max_sync(0xffffffffu, key);
    unsigned bal = __ballot_sync(0xffffffffu, key == kmax);
    if (__popc(bal) > 1) {
        const unsigned rmin = __reduce_min_sync(0xffffffffu, key == kmax ? rank : 0xffffffffu);
        bal = __ballot_sync(0xffffffffu, key == kmax && rank == rmin);
    }
    return __ffs(bal) - 1;
}

// Packed fp32x2 arithmetic (FADD2 / FMUL2 / FFMA2 on sm_100a): two points per instruction in the distance update.
// Each half is an IEEE round-to-nearest fp32 operation, so the results are bit-identical to the scalar sequence.
__device__ __forceinline__ unsigned long long pack2(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float &a, float &b) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ unsigned long long sub2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// One iteration (no hardware cluster barrier and no global load inside the loop):
//   1. every thread updates its P running minima and keeps its best candidate;
//   2. redux argmax in the warp; lane 0 leaves the warp's record {key, rank, slot} in shared memory; ONE __syncthreads;
//   3. warp 0 reduces the <= 32 warp records and lane c sends the CTA's winner {key, rank, x, y, z} with st.async
//      (data + complete_tx) into slot [crank] of CTA c's inbox;
//   4. every warp waits on its own CTA's inbox mbarrier (armed by thread 0 with expect_tx for one record per CTA),
//      reduces the cluster_size records and continues with the global winner's coordinates.
// Warp records, inbox and mbarrier are double-buffered by iteration parity: a CTA cannot run two iterations ahead of a
// peer because it needs that peer's record of the next iteration first, and a warp cannot overwrite a record buffer
// before warp 0 has read it because two __syncthreads lie in between.
// (First version: every WARP sent its record to every CTA - no __syncthreads, but 4x the DSMEM stores and a 64-record
// inbox scan per warp; stamps showed 535 of 2000 cycles per iteration in that scan alone.)
constexpr int kMaxCluster = 16;

template <int P, bool CLUSTER, int MAXT>
__global__ void __launch_bounds__(MAXT, 1)
fps_kernel(const float *__restrict__ xyz, const int *__restrict__ offset, const int *__restrict__ new_offset,
           int *__restrict__ idx, int logB, int cluster_size, long long *dbg) {
    extern __shared__ float sxyz[];  // [3][P * T] coordinates of this CTA's points, for winner look-up
    __shared__ uint4 inbox_a[2][kMaxCluster];   // key, rank, x, y  - one record per CTA of the cluster
    __shared__ float inbox_z[2][kMaxCluster];
    __shared__ uint4 wrec[2][kWarp];            // key, rank, slot, -  - one record per warp of this CTA
    __shared__ __align__(8) unsigned long long cbar[2];

    const int T = blockDim.x, tid = threadIdx.x, lane = tid % kWarp, warp = tid / kWarp, nwarps = T / kWarp;
    const int crank = CLUSTER ? (int)cg::this_cluster().block_rank() : 0;
    const int scene = blockIdx.x / cluster_size;
    const int TT = T * cluster_size, gtid = crank * T + tid;

    const int start_n = scene ? offset[scene - 1] : 0, n = offset[scene] - start_n;
    const int start_m = scene ? new_offset[scene - 1] : 0, m = new_offset[scene] - start_m;

    if (tid == 0) {
        mbar_init(smem_u32(&cbar[0]), 1);
        mbar_init(smem_u32(&cbar[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    constexpr int P2 = P / 2;   // point pairs handled with packed arithmetic; an odd last point stays scalar
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
    unsigned long long x2[P2 > 0 ? P2 : 1], y2[P2 > 0 ? P2 : 1], z2[P2 > 0 ? P2 : 1];
#pragma unroll
    for (int k = 0; k < P2; ++k) {
        x2[k] = pack2(px[2 * k], px[2 * k + 1]);
        y2[k] = pack2(py[2 * k], py[2 * k + 1]);
        z2[k] = pack2(pz[2 * k], pz[2 * k + 1]);
    }
    float ox = 0.f, oy = 0.f, oz = 0.f;
    if (n > 0) {
        ox = xyz[(size_t)start_n * 3 + 0];
        oy = xyz[(size_t)start_n * 3 + 1];
        oz = xyz[(size_t)start_n * 3 + 2];
    }
    if (gtid == 0 && m > 0) idx[start_m] = start_n;
    // destination of warp 0's sends: slot of this CTA in the inbox of CTA `lane`
    unsigned dst_a[2] = {0, 0}, dst_z[2] = {0, 0}, dst_bar[2] = {0, 0};
    if (warp == 0 && lane < cluster_size) {
#pragma unroll
        for (int bf = 0; bf < 2; ++bf) {
            dst_a[bf] = map_to_cta(smem_u32(&inbox_a[bf][crank]), (unsigned)lane);
            dst_z[bf] = map_to_cta(smem_u32(&inbox_z[bf][crank]), (unsigned)lane);
            dst_bar[bf] = map_to_cta(smem_u32(&cbar[bf]), (unsigned)lane);
        }
    }
    __syncthreads();
    if (CLUSTER) cg::this_cluster().sync();  // barriers initialised and every CTA resident before any remote store

    const unsigned tx_bytes = 20u * (unsigned)cluster_size;
    for (int j = 1; j < m; ++j) {
        const int buf = j & 1;
        const unsigned parity = (unsigned)((j - 1) >> 1) & 1u;  // k-th use of this buffer, k = (j-1)/2
        if (tid == 0) mbar_arrive_expect_tx(smem_u32(&cbar[buf]), tx_bytes);
        // development aid: per-phase clock stamps of warps 0 and 1 of CTA 0 for iterations 64..79
        const bool stamp = dbg && blockIdx.x == 0 && lane == 0 && warp < 2 && j >= 64 && j < 80;
        long long *ds = dbg + ((j - 64) * 2 + warp) * 8;
        if (stamp) ds[0] = clock64();
        float best = -2.f;
        int bu = 0;
        const unsigned long long ox2 = pack2(ox, ox), oy2 = pack2(oy, oy), oz2 = pack2(oz, oz);
#pragma unroll
        for (int k = 0; k < P2; ++k) {
            const unsigned long long dx = sub2(x2[k], ox2), dy = sub2(y2[k], oy2), dz = sub2(z2[k], oz2);
            float da, db;
            unpack2(fma2(dz, dz, fma2(dx, dx, mul2(dy, dy))), da, db);
            const float ma = fminf(da, mind[2 * k]), mb = fminf(db, mind[2 * k + 1]);
            mind[2 * k] = ma;
            mind[2 * k + 1] = mb;
            if (ma > best) {
                best = ma;
                bu = 2 * k;
            }
            if (mb > best) {
                best = mb;
                bu = 2 * k + 1;
            }
        }
        if (P & 1) {
            constexpr int u = P - 1;
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
        const unsigned key = best >= 0.f ? __float_as_uint(best) + 1u : 0u;
        const unsigned rank = fps_rank(gtid + bu * TT, logB);
        if (stamp) ds[1] = clock64();
        {
            const int src = warp_argmax_lane(key, rank);
            if (lane == src) wrec[buf][warp] = make_uint4(key, rank, (unsigned)(bu * T + tid), 0u);
        }
        __syncthreads();
        if (warp == 0) {
            const uint4 r = lane < nwarps ? wrec[buf][lane] : make_uint4(0u, 0xffffffffu, 0u, 0u);
            const int src = warp_argmax_lane(r.x, r.y);
            const unsigned wk = __shfl_sync(0xffffffffu, r.x, src), wr = __shfl_sync(0xffffffffu, r.y, src);
            const int wslot = (int)__shfl_sync(0xffffffffu, r.z, src);
            if (lane < cluster_size) {
                const float wx = sxyz[0 * P * T + wslot], wy = sxyz[1 * P * T + wslot], wz = sxyz[2 * P * T + wslot];
                st_async_v4(dst_a[buf], wk, wr, __float_as_uint(wx), __float_as_uint(wy), dst_bar[buf]);
                st_async_b32(dst_z[buf], __float_as_uint(wz), dst_bar[buf]);
            }
        }
        if (stamp) ds[2] = clock64();
        mbar_wait(smem_u32(&cbar[buf]), parity);
        if (stamp) ds[3] = clock64();
        {
            const uint4 a = lane < cluster_size ? inbox_a[buf][lane] : make_uint4(0u, 0xffffffffu, 0u, 0u);
            const int src = warp_argmax_lane(a.x, a.y);
            const uint4 w = inbox_a[buf][src];
            ox = __uint_as_float(w.z);
            oy = __uint_as_float(w.w);
            oz = inbox_z[buf][src];
            if (gtid == 0) idx[start_m + j] = start_n + fps_unrank(w.y, logB);
        }
        if (stamp) ds[4] = clock64();
    }
    if (CLUSTER) cg::this_cluster().sync();  // no CTA exits while a peer may still write into its smem
}

// Fallback for scenes too large for the register-resident kernel: one CTA per scene, coordinates and running
// minima streamed from global memory (the caller's `tmp` scratch), same (max distance, min rank) reduction.
__global__ void __launch_bounds__(1024, 1)
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

constexpr int kFpsRetrySmallerCluster = -1;
constexpr int kFpsNotApplicable = -2;   // this (points per thread, CTA size) combination exceeds the register budget: use the streaming kernel
static long long *g_fps_dbg = nullptr;  // development aid, see stb200_fps_debug_buffer

static int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return v ? atoi(v) : dflt;
}

template <int P, int MAXT = (P <= 8 ? 1024 : (P <= 20 ? 512 : 256))>  // register budget: 4 P + ~40 per thread
static int launch_fps(int b, int cs, int threads, const float *xyz, const int *offset, const int *new_offset, int *idx,
                      int logB, cudaStream_t s) {
    // The tie order forces cs * threads to be a multiple of the reference's virtual block size, so many small-cluster
    // configurations (e.g. b = 20 scenes of 30k points: cluster 2 x 512 threads x 30 points) need more threads than the
    // register budget of P points per thread allows.  Not an error: the caller falls back to the streaming kernel.
    if (threads > MAXT) return kFpsNotApplicable;
    STB200_REQUIRE(threads % 32 == 0, STB200_ERR_ARG, "fps: %d threads with %d points per thread", threads, P);
    const size_t smem = (size_t)3 * P * threads * sizeof(float);
    cudaError_t e;
    KernelScope ks("fps_cluster", 0.0, s);  // latency-bound by construction: bytes are not the meaningful unit
    {
        auto kern = fps_kernel<P, true, MAXT>;
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess ||
            (e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1)) != cudaSuccess) {
            set_error("fps attr: %s", cudaGetErrorString(e));
            return STB200_ERR_CUDA;
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(b * cs);
        cfg.blockDim = dim3(threads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = cs;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        if (cs > 1) {   // all scenes must be co-resident, otherwise the clusters run in waves: retry with a smaller cluster
            int max_clusters = 0;
            const cudaError_t oe = cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg);
            if (getenv("STB200_FPS_DEBUG"))
                fprintf(stderr, "[stb200 fps] b=%d cluster=%d threads=%d P=%d smem=%zu -> max active clusters %d (%s)\n", b, cs,
                        threads, P, smem, max_clusters, cudaGetErrorString(oe));
            if (oe == cudaSuccess && max_clusters < b &&
                max_clusters * 2 > 0 && (b + max_clusters - 1) / max_clusters > (b * (cs / 2) > kNumSMs ? 2 : 1))
                return kFpsRetrySmallerCluster;
        }
        if ((e = cudaLaunchKernelEx(&cfg, kern, xyz, offset, new_offset, idx, logB, cs, g_fps_dbg)) != cudaSuccess) {
            set_error("fps cluster launch (cluster=%d): %s", cs, cudaGetErrorString(e));
            return STB200_ERR_CUDA;
        }
    }
    return check_launch("fps");
}

}  // namespace stb200

using namespace stb200;

extern "C" void stb200_fps_debug_buffer(long long *device_buffer /* >= 16*2*8 int64, or NULL */) { g_fps_dbg = device_buffer; }

extern "C" int stb200_furthestsampling(int b, int n, const float *xyz, const int *offset, const int *new_offset,
                                       float *tmp, int *idx, void *stream) {
    STB200_REQUIRE(b >= 0 && n >= 0, STB200_ERR_ARG, "bad sizes b=%d n=%d", b, n);
    if (b == 0 || n == 0) return STB200_OK;
    STB200_REQUIRE(xyz && offset && new_offset && idx, STB200_ERR_ARG, "null pointer");
    STB200_REQUIRE(n < (1 << 30), STB200_ERR_ARG, "scene too large");
    cudaStream_t s = (cudaStream_t)stream;
    const int logB = ref_block_log2(n);
    // cluster size: as many CTAs per scene as keep all scenes co-resident (148 SMs), but never more threads than points
    // tuning overrides (development): STB200_FPS_CLUSTER caps the cluster size, STB200_FPS_THREADS sets the CTA size
    const int threads = env_int("STB200_FPS_THREADS", kFpsThreads);
    // target points per thread when choosing the cluster size: small scenes spread over more CTAs (measured)
    const int per_thread = env_int("STB200_FPS_POINTS", n > 8192 ? 10 : 3);
    int cs = 1;
    // All clusters together take at most half of the SMs: the kernel is latency-bound (one CTA per SM, few issue
    // slots used) and normally runs beside the attention kernels of the previous batch (GeometryPrefetcher); 8 scenes
    // x 8 CTAs x 256 threads cost 8 % more FPS time than 16 x 128 but 2.4 % less step time in that pipeline.
    const int max_sms = env_int("STB200_FPS_MAX_SMS", kNumSMs / 2);
    while (cs < env_int("STB200_FPS_CLUSTER", kMaxCluster) && b * cs * 2 <= max_sms && cs * threads * per_thread < n) cs <<= 1;
    for (; cs >= 1; cs >>= 1) {
        // A thread's points are i = gtid + u * (cs * threads); "first strict maximum inside the thread" equals the
        // reference's tie order only if they all share i mod B, i.e. cs * threads must be a multiple of B.
        int threads_cs = threads;
        while (cs * threads_cs < (1 << logB) && threads_cs < 1024) threads_cs <<= 1;
        const int threads = threads_cs;
        const int need = (n + cs * threads - 1) / (cs * threads);
        int rc = STB200_ERR_ARG;
        if (need <= 1) rc = launch_fps<1>(b, cs, threads, xyz, offset, new_offset, idx, logB, s);
        else if (need <= 2) rc = launch_fps<2>(b, cs, threads, xyz, offset, new_offset, idx, logB, s);
        else if (need <= 3) rc = launch_fps<3>(b, cs, threads, xyz, offset, new_offset, idx, logB, s);
        else if (need <= 4) rc = launch_fps<4>(b, cs, threads, xyz, offset, new_offset, idx, logB, s);
        else if (need <= 5) rc = launch_fps<5>(b, cs, threads, xyz, offset, new_offset, idx, logB, s);
        else if (need <= 6) rc = launch_fps<6>(b, cs, threads, xyz, offset, new_offset, idx, logB, s);
        else if (need <= 8) rc = launch_fps<8>(b, cs, threads, xyz, offset, new_offset, idx, logB, s);
        else if (need <= 10) rc = launch_fps<10>(b, cs, threads, xyz, offset, new_offset, idx, logB, s);
        else if (need <= 12) rc = launch_fps<12>(b, cs, threads, xyz, offset, new_offset, idx, logB, s);
        else if (need <= 16) rc = launch_fps<16>(b, cs, threads, xyz, offset, new_offset, idx, logB, s);
        else if (need <= 20) rc = launch_fps<20>(b, cs, threads, xyz, offset, new_offset, idx, logB, s);
        else if (need <= 24) rc = launch_fps<24>(b, cs, threads, xyz, offset, new_offset, idx, logB, s);
        else if (need <= 32) rc = launch_fps<32>(b, cs, threads, xyz, offset, new_offset, idx, logB, s);
        else if (need <= 40) rc = launch_fps<40>(b, cs, threads, xyz, offset, new_offset, idx, logB, s);
        else break;   // too many points per thread for the register-resident kernel
        if (rc == kFpsNotApplicable) break;   // a smaller cluster only needs more points per thread
        if (rc != kFpsRetrySmallerCluster) return rc;
    }
    STB200_REQUIRE(tmp, STB200_ERR_ARG, "scene of %d points needs the tmp scratch (streaming path)", n);
    {
        KernelScope ks("fps_streaming", 0.0, s);
        fps_streaming_kernel<<<b, 1024, 0, s>>>(xyz, offset, new_offset, tmp, idx, logB);
    }
    return check_launch("fps_streaming");
}
