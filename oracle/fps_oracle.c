/* CPU oracle: exact furthest point sampling (TEST INFRASTRUCTURE; never linked into the product).
 *
 * Literal restatement of the reference kernel's arithmetic and tie behaviour:
 *   /root/reference/lib/pointops2/src/sampling/sampling_cuda_kernel.cu:14-129 (kernel),
 *   :5-10 (__update: max value, `v2 > v1 ? i2 : i1`), :131-171 (block size = opt_n_threads(n)),
 *   /root/reference/lib/pointops2/src/cuda_utils.h:10-13 (opt_n_threads).
 * The distance is contracted exactly like the SASS nvcc 12.9 emits for that source at sm_100a
 * (FMUL dy*dy; FFMA dx,dx; FFMA dz,dz  — checked with cuobjdump on oracle/_ref), so build this
 * file with -ffp-contract=off: the fmaf() calls below are the only fused operations.
 *
 * Pinning: tests/test_gpu_parity.py (-m gpu) compares against the reference kernel itself
 * (oracle/_ref/libpointops2_ref.so) on continuous and lattice-aligned (tie-heavy) scenes.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

int fps_oracle_block_size(int n_max) {
    /* opt_n_threads: 2^floor(log2 n) clamped to [1, 1024], computed in double like the reference */
    int pow_2 = (int)(log((double)n_max) / log(2.0));
    int t = 1 << pow_2;
    if (t > 1024) t = 1024;
    if (t < 1) t = 1;
    return t;
}

/* xyz [n,3] f32, offset/new_offset cumulative int32 [b], tmp [n] f32 scratch (filled with 1e10 here),
 * idx [new_offset[b-1]] int32 out.  n_max = largest scene size (what the Python wrapper passes). */
void fps_oracle(int b, int n_max, const float *xyz, const int *offset, const int *new_offset,
                float *tmp, int *idx) {
    const int B = fps_oracle_block_size(n_max);
    float *dists = (float *)malloc(sizeof(float) * B);
    int *dists_i = (int *)malloc(sizeof(int) * B);
    for (int s = 0; s < b; ++s) {
        const int start_n = s ? offset[s - 1] : 0, end_n = offset[s];
        const int start_m = s ? new_offset[s - 1] : 0, end_m = new_offset[s];
        int old = start_n;
        for (int k = start_n; k < end_n; ++k) tmp[k] = 1e10f;
        if (start_m < end_m) idx[start_m] = start_n;
        for (int j = start_m + 1; j < end_m; ++j) {
            const float x1 = xyz[old * 3 + 0], y1 = xyz[old * 3 + 1], z1 = xyz[old * 3 + 2];
            for (int t = 0; t < B; ++t) {           /* one "thread" at a time */
                int besti = start_n;
                float best = -1.0f;
                for (int k = start_n + t; k < end_n; k += B) {
                    const float dx = xyz[k * 3 + 0] - x1, dy = xyz[k * 3 + 1] - y1, dz = xyz[k * 3 + 2] - z1;
                    const float d = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
                    const float d2 = d < tmp[k] ? d : tmp[k];
                    tmp[k] = d2;
                    if (d2 > best) { besti = k; best = d2; }
                }
                dists[t] = best;
                dists_i[t] = besti;
            }
            for (int stride = B / 2; stride >= 1; stride >>= 1)      /* halving tree */
                for (int t = 0; t < stride; ++t) {
                    const float v1 = dists[t], v2 = dists[t + stride];
                    const int i1 = dists_i[t], i2 = dists_i[t + stride];
                    dists[t] = v1 > v2 ? v1 : v2;
                    dists_i[t] = v2 > v1 ? i2 : i1;
                }
            old = dists_i[0];
            idx[j] = old;
        }
    }
    free(dists);
    free(dists_i);
}

/* ------------------------------------------------------------------------------------------------
 * kNN oracle: literal restatement of /root/reference/lib/pointops2/src/knnquery/knnquery_cuda_kernel.cu:21-108
 * (per query: scan the scene in index order, binary max-heap of k, heap sort), distance contracted like the SASS
 * of the reference build: fma(dz,dz, fma(dx,dx, dy*dy)).  Pinned on the GPU against the reference kernel itself
 * (tests/test_gpu_knn.py). */
static void knn_reheap(float *dist, int *idx, int k) {
    int root = 0, child = 1;
    while (child < k) {
        if (child + 1 < k && dist[child + 1] > dist[child]) child++;
        if (dist[root] > dist[child]) return;
        float td = dist[root]; dist[root] = dist[child]; dist[child] = td;
        int ti = idx[root]; idx[root] = idx[child]; idx[child] = ti;
        root = child;
        child = root * 2 + 1;
    }
}

void knn_oracle(int m, int b, int k, const float *xyz, const float *new_xyz, const int *offset, const int *new_offset,
                int *idx, float *dist2) {
    float bd[100];
    int bi[100];
    for (int q = 0; q < m; ++q) {
        int s = 0;
        while (s < b - 1 && q >= new_offset[s]) ++s;
        const int start = s ? offset[s - 1] : 0, end = offset[s];
        const float nx = new_xyz[q * 3], ny = new_xyz[q * 3 + 1], nz = new_xyz[q * 3 + 2];
        for (int i = 0; i < k; ++i) { bd[i] = 1e10f; bi[i] = start; }
        for (int i = start; i < end; ++i) {
            const float dx = nx - xyz[i * 3], dy = ny - xyz[i * 3 + 1], dz = nz - xyz[i * 3 + 2];
            const float d2 = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
            if (d2 < bd[0]) { bd[0] = d2; bi[0] = i; knn_reheap(bd, bi, k); }
        }
        for (int i = k - 1; i > 0; --i) {
            float td = bd[0]; bd[0] = bd[i]; bd[i] = td;
            int ti = bi[0]; bi[0] = bi[i]; bi[i] = ti;
            knn_reheap(bd, bi, i);
        }
        for (int i = 0; i < k; ++i) { idx[q * k + i] = bi[i]; dist2[q * k + i] = bd[i]; }
    }
}
