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

#include <cub/cub.cuh>

namespace cg = cooperative_groups;

namespace stb200 {

__device__ int g_fps_stamp_start = 64;   // development aid: first iteration of the 16-iteration clock-stamp window
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
        const bool stamp = dbg && blockIdx.x == 0 && lane == 0 && warp < 2 && j >= g_fps_stamp_start && j < g_fps_stamp_start + 16;
        long long *ds = dbg + ((j - g_fps_stamp_start) * 2 + warp) * 8;
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

// ---- pruned variant ----------------------------------------------------------------------------------------------------
// Same iteration structure, but the distance update only touches the points it can change.  The scene is first sorted along
// a Morton curve (fps_prepare) and dealt out in GROUPS of 256 consecutive sorted points = 8 slots x 32 lanes of one warp, so
// a group is a compact cell with a small bounding box.  A group is skipped in an iteration when the new sample is at least
// as far from its box as the group's largest running minimum: then fmin(d, mind) = mind for every point in it.  That is
// exact, not approximate: the box distance is built from the same fp32 operations as the point distance (subtract,
// multiply, two fma), each of which is monotone, so lb <= d for every point of the box in floating point as well.
// Neighbouring groups go to different CTAs / warps (group id modulo the warp count) so that the few groups around the new
// sample - about 2 % of them after the first hundred iterations - do not queue up in one warp.
// Inside a thread the points no longer share the bit-reversed part of the rank, so ties inside a thread are resolved by
// rank explicitly (rare with continuous coordinates; lattice scenes take that branch often and stay exact).
constexpr int kFpsGS = 8;   // slots per lane and group

__device__ __forceinline__ float fps_box_dist(const float4 lo, const float4 hi, float ox, float oy, float oz) {
    const float dx = fmaxf(fmaxf(__fsub_rn(lo.x, ox), __fsub_rn(ox, hi.x)), 0.f);
    const float dy = fmaxf(fmaxf(__fsub_rn(lo.y, oy), __fsub_rn(oy, hi.y)), 0.f);
    const float dz = fmaxf(fmaxf(__fsub_rn(lo.z, oz), __fsub_rn(oz, hi.z)), 0.f);
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

template <int NG, bool CLUSTER, int MAXT>
__global__ void __launch_bounds__(MAXT, 1)
fps_pruned_kernel(const float *__restrict__ xyz, const int *__restrict__ perm, const int *__restrict__ offset,
                  const int *__restrict__ new_offset, int *__restrict__ idx, int logB, int cluster_size, long long *dbg) {
    constexpr int P = NG * kFpsGS;
    extern __shared__ float sxyz[];  // [3][P * T] coordinates, then [P * T] original indices, then the boxes
    __shared__ uint4 inbox_a[2][kMaxCluster];
    __shared__ float inbox_z[2][kMaxCluster];
    __shared__ uint4 wrec[2][kWarp];
    __shared__ __align__(8) unsigned long long cbar[2];

    const int T = blockDim.x, tid = threadIdx.x, lane = tid % kWarp, warp = tid / kWarp, nwarps = T / kWarp;
    const int crank = CLUSTER ? (int)cg::this_cluster().block_rank() : 0;
    const int scene = blockIdx.x / cluster_size;
    int *sorig = reinterpret_cast<int *>(sxyz + 3 * P * T);
    float4 *sbox = reinterpret_cast<float4 *>(sxyz + 4 * P * T);   // [nwarps][NG + 1][2]: lo, hi; entry NG = the whole warp

    const int start_n = scene ? offset[scene - 1] : 0, n = offset[scene] - start_n;
    const int start_m = scene ? new_offset[scene - 1] : 0, m = new_offset[scene] - start_m;

    if (tid == 0) {
        mbar_init(smem_u32(&cbar[0]), 1);
        mbar_init(smem_u32(&cbar[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const int W = cluster_size * nwarps, wi = warp * cluster_size + crank;
    float px[P], py[P], pz[P], mind[P];
    float gmaxw[NG];   // warp-uniform: largest running minimum of the group
    float gbest[NG];   // this lane's best of the group
    int gbu[NG];
    float wlo[3] = {3e38f, 3e38f, 3e38f}, whi[3] = {-3e38f, -3e38f, -3e38f};
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f};
        bool any = false;
#pragma unroll
        for (int e = 0; e < kFpsGS; ++e) {
            const int u = g * kFpsGS + e;
            const long long sp = ((long long)(g * W + wi) * kFpsGS + e) * kWarp + lane;
            const bool valid = sp < n;
            int gi = 0;
            if (valid) gi = perm[start_n + (int)sp];
            px[u] = valid ? xyz[(size_t)gi * 3 + 0] : 0.f;
            py[u] = valid ? xyz[(size_t)gi * 3 + 1] : 0.f;
            pz[u] = valid ? xyz[(size_t)gi * 3 + 2] : 0.f;
            mind[u] = valid ? 1e10f : -1.f;
            sxyz[0 * P * T + u * T + tid] = px[u];
            sxyz[1 * P * T + u * T + tid] = py[u];
            sxyz[2 * P * T + u * T + tid] = pz[u];
            sorig[u * T + tid] = valid ? gi - start_n : 0;
            if (valid) {
                any = true;
                lo[0] = fminf(lo[0], px[u]); lo[1] = fminf(lo[1], py[u]); lo[2] = fminf(lo[2], pz[u]);
                hi[0] = fmaxf(hi[0], px[u]); hi[1] = fmaxf(hi[1], py[u]); hi[2] = fmaxf(hi[2], pz[u]);
            }
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) {
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
                hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
            }
            wlo[a] = fminf(wlo[a], lo[a]);
            whi[a] = fmaxf(whi[a], hi[a]);
        }
        const bool any_w = __any_sync(0xffffffffu, any);
        if (lane == 0) {
            sbox[(warp * (NG + 1) + g) * 2 + 0] = make_float4(lo[0], lo[1], lo[2], 0.f);
            sbox[(warp * (NG + 1) + g) * 2 + 1] = make_float4(hi[0], hi[1], hi[2], 0.f);
        }
        gmaxw[g] = any_w ? 1e10f : 0.f;   // 1e10: the first iteration updates every non-empty group
        gbest[g] = any ? 1e10f : -1.f;
        gbu[g] = g * kFpsGS;
    }
    if (lane == 0) {
        sbox[(warp * (NG + 1) + NG) * 2 + 0] = make_float4(wlo[0], wlo[1], wlo[2], 0.f);
        sbox[(warp * (NG + 1) + NG) * 2 + 1] = make_float4(whi[0], whi[1], whi[2], 0.f);
    }
    float wmaxw = 0.f;
#pragma unroll
    for (int g = 0; g < NG; ++g) wmaxw = fmaxf(wmaxw, gmaxw[g]);
    unsigned rk = 0u, rr = 0xffffffffu;   // this warp's candidate record, valid until one of its groups is updated
    int rslot = 0;

    float ox = 0.f, oy = 0.f, oz = 0.f;
    if (n > 0) {
        ox = xyz[(size_t)start_n * 3 + 0];
        oy = xyz[(size_t)start_n * 3 + 1];
        oz = xyz[(size_t)start_n * 3 + 2];
    }
    if (crank == 0 && tid == 0 && m > 0) idx[start_m] = start_n;
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
    if (CLUSTER) cg::this_cluster().sync();

    const unsigned tx_bytes = 20u * (unsigned)cluster_size;
    for (int j = 1; j < m; ++j) {
        const int buf = j & 1;
        const unsigned parity = (unsigned)((j - 1) >> 1) & 1u;
        if (tid == 0) mbar_arrive_expect_tx(smem_u32(&cbar[buf]), tx_bytes);
        const bool stamp = dbg && blockIdx.x == 0 && lane == 0 && warp < 2 && j >= g_fps_stamp_start && j < g_fps_stamp_start + 16;
        long long *ds = dbg + ((j - g_fps_stamp_start) * 2 + warp) * 8;
        if (stamp) ds[0] = clock64();
        bool changed = false;
        {
            // all box distances first (independent loads and arithmetic), then the warp-uniform decisions
            float lb[NG];
#pragma unroll
            for (int g = 0; g < NG; ++g)
                lb[g] = fps_box_dist(sbox[(warp * (NG + 1) + g) * 2], sbox[(warp * (NG + 1) + g) * 2 + 1], ox, oy, oz);
            const unsigned long long ox2 = pack2(ox, ox), oy2 = pack2(oy, oy), oz2 = pack2(oz, oz);
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                if (lb[g] < gmaxw[g]) {   // warp-uniform
                    float best = -2.f;
                    int bu = g * kFpsGS;
                    bool tie = false;
#pragma unroll
                    for (int e = 0; e < kFpsGS; e += 2) {
                        const int u = g * kFpsGS + e;
                        const unsigned long long dx = sub2(pack2(px[u], px[u + 1]), ox2), dy = sub2(pack2(py[u], py[u + 1]), oy2),
                                                 dz = sub2(pack2(pz[u], pz[u + 1]), oz2);
                        float da, db;
                        unpack2(fma2(dz, dz, fma2(dx, dx, mul2(dy, dy))), da, db);
                        const float ma = fminf(da, mind[u]), mb = fminf(db, mind[u + 1]);
                        mind[u] = ma;
                        mind[u + 1] = mb;
                        // branch-free candidate tracking; `tie` = another point equals the current best
                        tie = ma > best ? false : (tie || ma == best);
                        bu = ma > best ? u : bu;
                        best = fmaxf(best, ma);
                        tie = mb > best ? false : (tie || mb == best);
                        bu = mb > best ? u + 1 : bu;
                        best = fmaxf(best, mb);
                    }
                    if (__any_sync(0xffffffffu, tie && best >= 0.f)) {   // rare (lattice data): lowest rank among the equal ones
                        if (tie && best >= 0.f) {
                            unsigned br = 0xffffffffu;
#pragma unroll
                            for (int e = 0; e < kFpsGS; ++e) {
                                const int u = g * kFpsGS + e;
                                if (mind[u] == best) {
                                    const unsigned r = fps_rank(sorig[u * T + tid], logB);
                                    if (r < br) { br = r; bu = u; }
                                }
                            }
                        }
                    }
                    gbest[g] = best;
                    gbu[g] = bu;
                    gmaxw[g] = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(fmaxf(best, 0.f))));
                    changed = true;
                }
            }
        }
        if (stamp) ds[1] = clock64();
        if (changed) {   // warp-uniform: a group of this warp was touched, its candidate must be recomputed
            float best = gbest[0];
            int bu = gbu[0];
            bool tie = false;
            wmaxw = gmaxw[0];
#pragma unroll
            for (int g = 1; g < NG; ++g) {
                wmaxw = fmaxf(wmaxw, gmaxw[g]);
                tie = gbest[g] > best ? false : (tie || gbest[g] == best);
                bu = gbest[g] > best ? gbu[g] : bu;
                best = fmaxf(best, gbest[g]);
            }
            if (__any_sync(0xffffffffu, tie && best >= 0.f)) {
                if (tie && best >= 0.f) {
                    unsigned br = 0xffffffffu;
#pragma unroll
                    for (int g = 0; g < NG; ++g)
                        if (gbest[g] == best) {
                            const unsigned r = fps_rank(sorig[gbu[g] * T + tid], logB);
                            if (r < br) { br = r; bu = gbu[g]; }
                        }
                }
            }
            const unsigned key = best >= 0.f ? __float_as_uint(best) + 1u : 0u;
            const unsigned rank = best >= 0.f ? fps_rank(sorig[bu * T + tid], logB) : 0xffffffffu;
            const int src = warp_argmax_lane(key, rank);
            rk = __shfl_sync(0xffffffffu, key, src);
            rr = __shfl_sync(0xffffffffu, rank, src);
            rslot = __shfl_sync(0xffffffffu, bu * T + tid, src);
        }
        if (lane == 0) wrec[buf][warp] = make_uint4(rk, rr, (unsigned)rslot, 0u);
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
            if (crank == 0 && tid == 0) idx[start_m + j] = start_n + fps_unrank(w.y, logB);
        }
        if (stamp) ds[4] = clock64();
    }
    if (CLUSTER) cg::this_cluster().sync();
}

// Morton order of every scene's points (10 bits per axis inside the scene's bounding box): perm[start + s] = point at sorted
// position s of its scene.  Scenes stay contiguous because the scene id is the high part of the key.
__device__ __forceinline__ unsigned fps_spread3(unsigned v) {
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x30000ffu;
    v = (v | (v << 8)) & 0x300f00fu;
    v = (v | (v << 4)) & 0x30c30c3u;
    v = (v | (v << 2)) & 0x9249249u;
    return v;
}

__global__ void fps_scene_box_kernel(const float *__restrict__ xyz, const int *__restrict__ offset, float *__restrict__ box) {
    __shared__ float red[6][32];
    const int s = blockIdx.x;
    const int start = s ? offset[s - 1] : 0, n = offset[s] - start;
    float mn[3] = {3e38f, 3e38f, 3e38f}, mx[3] = {-3e38f, -3e38f, -3e38f};
    for (int i = threadIdx.x; i < n; i += blockDim.x)
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float v = __ldg(xyz + (size_t)(start + i) * 3 + a);
            mn[a] = fminf(mn[a], v);
            mx[a] = fmaxf(mx[a], v);
        }
    const int lane = threadIdx.x % 32, warp = threadIdx.x / 32, nw = blockDim.x / 32;
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
    if (threadIdx.x < 3) {
        const int a = threadIdx.x;
        float lo = red[a][0], hi = red[3 + a][0];
        for (int w = 1; w < nw; ++w) { lo = fminf(lo, red[a][w]); hi = fmaxf(hi, red[3 + a][w]); }
        box[s * 6 + a] = lo;
        box[s * 6 + 3 + a] = 1024.f / fmaxf(hi - lo, 1e-12f);
    }
}

__global__ void fps_morton_kernel(int N, int b, const float *__restrict__ xyz, const int *__restrict__ offset, const float *__restrict__ box,
                                  unsigned long long *__restrict__ keys, int *__restrict__ vals) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        int lo = 0, hi = b - 1;   // scene of point i
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(offset + mid) > i) hi = mid; else lo = mid + 1;
        }
        const float *bx = box + lo * 6;
        unsigned q[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) q[a] = (unsigned)fminf(fmaxf((__ldg(xyz + (size_t)i * 3 + a) - bx[a]) * bx[3 + a], 0.f), 1023.f);
        keys[i] = ((unsigned long long)lo << 30) | (fps_spread3(q[0]) | (fps_spread3(q[1]) << 1) | (fps_spread3(q[2]) << 2));
        vals[i] = i;
    }
}

struct FpsScratch {
    float *box;
    unsigned long long *keys_in, *keys_out;
    int *vals_in, *perm;
    void *cub_tmp;
    size_t cub_bytes, total;
};

static FpsScratch fps_layout(int N, int b, void *base) {
    FpsScratch st{};
    char *p = (char *)base;
    size_t o = 0;
    auto take = [&](size_t bytes) { char *r = p ? p + o : nullptr; o += (bytes + 255) / 256 * 256; return r; };
    st.box = (float *)take((size_t)b * 6 * sizeof(float));
    st.keys_in = (unsigned long long *)take((size_t)N * 8);
    st.keys_out = (unsigned long long *)take((size_t)N * 8);
    st.vals_in = (int *)take((size_t)N * 4);
    st.perm = (int *)take((size_t)N * 4);
    st.cub_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, st.cub_bytes, (const unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                    (const int *)nullptr, (int *)nullptr, N, 0, 64);
    st.cub_tmp = take(st.cub_bytes);
    st.total = o;
    return st;
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


template <int NG, int MAXT = (NG <= 2 ? 512 : 256)>   // register budget: 4 * 8 NG + ~60 per thread
static int launch_fps_pruned(int b, int cs, int threads, const float *xyz, const int *perm, const int *offset, const int *new_offset,
                             int *idx, int logB, cudaStream_t s) {
    if (threads > MAXT) return kFpsNotApplicable;
    constexpr int P = NG * kFpsGS;
    const size_t smem = (size_t)4 * P * threads * sizeof(float) + (size_t)(threads / 32) * (NG + 1) * 2 * sizeof(float4);
    if (smem > 200 * 1024) return kFpsNotApplicable;
    cudaError_t e;
    KernelScope ks("fps_cluster", 0.0, s);
    auto kern = fps_pruned_kernel<NG, true, MAXT>;
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
    if (cs > 1) {
        int max_clusters = 0;
        const cudaError_t oe = cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg);
        if (getenv("STB200_FPS_DEBUG"))
            fprintf(stderr, "[stb200 fps pruned] b=%d cluster=%d threads=%d P=%d smem=%zu -> max active clusters %d (%s)\n", b, cs, threads,
                    P, smem, max_clusters, cudaGetErrorString(oe));
        if (oe == cudaSuccess && max_clusters < b && (b + max(max_clusters, 1) - 1) / max(max_clusters, 1) > (b * (cs / 2) > kNumSMs ? 2 : 1))
            return kFpsRetrySmallerCluster;
    }
    if ((e = cudaLaunchKernelEx(&cfg, kern, xyz, perm, offset, new_offset, idx, logB, cs, g_fps_dbg)) != cudaSuccess) {
        set_error("fps pruned launch (cluster=%d): %s", cs, cudaGetErrorString(e));
        return STB200_ERR_CUDA;
    }
    return check_launch("fps_pruned");
}

}  // namespace stb200

using namespace stb200;

extern "C" void stb200_fps_debug_buffer(long long *device_buffer /* >= 16*2*8 int64, or NULL */) { g_fps_dbg = device_buffer; }
extern "C" void stb200_fps_debug_window(int first_iteration) { cudaMemcpyToSymbol(g_fps_stamp_start, &first_iteration, sizeof(int)); }

static int fps_run(int b, int n, const float *xyz, const int *offset, const int *new_offset, float *tmp, int *idx, const int *perm,
                   void *stream) {
    STB200_REQUIRE(b >= 0 && n >= 0, STB200_ERR_ARG, "bad sizes b=%d n=%d", b, n);
    if (b == 0 || n == 0) return STB200_OK;
    STB200_REQUIRE(xyz && offset && new_offset && idx, STB200_ERR_ARG, "null pointer");
    STB200_REQUIRE(n < (1 << 30), STB200_ERR_ARG, "scene too large");
    cudaStream_t s = (cudaStream_t)stream;
    const int logB = ref_block_log2(n);
    // cluster size: as many CTAs per scene as keep all scenes co-resident (148 SMs), but never more threads than points
    // tuning overrides (development): STB200_FPS_CLUSTER caps the cluster size, STB200_FPS_THREADS sets the CTA size
    // one or two large scenes get 16 CTAs each either way; 128-thread CTAs (40 points per thread) then measured 4.6 % faster than
    // 256-thread ones (fewer warps in the CTA-level argmax): 15.3 vs 16.1 ms for 20 001 samples of one 80k-point scene
    const int threads = env_int("STB200_FPS_THREADS", (b <= 2 && n >= 65536) ? 128 : kFpsThreads);
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
        if (perm && need >= 12 && need <= 40) {   // pruned kernel (points dealt out in sorted groups of 8 per lane)
            const int ng = (need + kFpsGS - 1) / kFpsGS;
            if (ng <= 2) rc = launch_fps_pruned<2>(b, cs, threads, xyz, perm, offset, new_offset, idx, logB, s);
            else if (ng == 3) rc = launch_fps_pruned<3>(b, cs, threads, xyz, perm, offset, new_offset, idx, logB, s);
            else if (ng == 4) rc = launch_fps_pruned<4>(b, cs, threads, xyz, perm, offset, new_offset, idx, logB, s);
            else rc = launch_fps_pruned<5>(b, cs, threads, xyz, perm, offset, new_offset, idx, logB, s);
            if (rc == kFpsRetrySmallerCluster) continue;
            if (rc != kFpsNotApplicable) return rc;
        }
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

extern "C" int stb200_furthestsampling(int b, int n, const float *xyz, const int *offset, const int *new_offset,
                                       float *tmp, int *idx, void *stream) {
    return fps_run(b, n, xyz, offset, new_offset, tmp, idx, nullptr, stream);
}

extern "C" size_t stb200_fps_workspace_bytes(int N, int b) { return N > 0 && b > 0 ? fps_layout(N, b, nullptr).total : 0; }

/* The same sampling with the exact bounding-box pruning of fps_pruned_kernel: N = total number of points; workspace from
 * stb200_fps_workspace_bytes(N, b).  Opt-in: STB200_FPS_PRUNE=1 in the environment, otherwise identical to stb200_furthestsampling. */
extern "C" int stb200_furthestsampling_ws(int b, int n, int N, const float *xyz, const int *offset, const int *new_offset, float *tmp,
                                          int *idx, void *workspace, size_t workspace_bytes, void *stream) {
    // Off by default: measured SLOWER than the plain kernel (8 x 80k points: 11.3 vs 9.5 ms).  The update it saves (825 cycles per
    // iteration for 40 points per thread) is replaced by five box tests (~390 cycles for the CTA's eight warps), the busiest warp's
    // one or two group updates and the recomputation of its candidate - and the iteration waits for the busiest warp, not for the
    // average one.  Kept as an exact, tested variant (STB200_FPS_PRUNE=1) with that evidence: DESIGN.md section 3.
    const int prune = env_int("STB200_FPS_PRUNE", 0);
    if (!workspace || !prune || b <= 0 || n <= 0 || N <= 0 || n < 12 * kFpsThreads)   // small scenes: too few points per thread
        return fps_run(b, n, xyz, offset, new_offset, tmp, idx, nullptr, stream);
    STB200_REQUIRE(xyz && offset && new_offset && idx, STB200_ERR_ARG, "null pointer");
    FpsScratch st = fps_layout(N, b, workspace);
    STB200_REQUIRE(workspace_bytes >= st.total, STB200_ERR_WORKSPACE, "fps workspace: %zu B given, %zu B needed", workspace_bytes, st.total);
    cudaStream_t s = (cudaStream_t)stream;
    {
        KernelScope ks("fps_prepare", 0.0, s);
        fps_scene_box_kernel<<<b, 256, 0, s>>>(xyz, offset, st.box);
        fps_morton_kernel<<<min((N + 255) / 256, kNumSMs * 8), 256, 0, s>>>(N, b, xyz, offset, st.box, st.keys_in, st.vals_in);
        int bits = 30;
        while ((1 << (bits - 30)) < b) ++bits;
        size_t tb = st.cub_bytes;
        const cudaError_t e = cub::DeviceRadixSort::SortPairs(st.cub_tmp, tb, st.keys_in, st.keys_out, st.vals_in, st.perm, N, 0, bits, s);
        STB200_REQUIRE(e == cudaSuccess, STB200_ERR_CUDA, "fps sort: %s", cudaGetErrorString(e));
    }
    return fps_run(b, n, xyz, offset, new_offset, tmp, idx, st.perm, stream);
}
