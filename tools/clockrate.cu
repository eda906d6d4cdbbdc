// Development aid: ratio of clock64() ticks to globaltimer nanoseconds for a busy and a mostly-waiting kernel.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void spin(long long *out, int iters, int mode) {
    unsigned long long t0, t1;
    long long c0 = clock64();
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    float x = threadIdx.x;
    for (int i = 0; i < iters; ++i) {
        if (mode == 0) { x = x * 1.0001f + 0.5f; }
        else { __nanosleep(100); }
    }
    long long c1 = clock64();
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = c1 - c0; out[1] = (long long)(t1 - t0); out[2] = (long long)x; }
}
int main() {
    long long *d, h[3];
    cudaMalloc(&d, 24);
    for (int mode = 0; mode < 2; ++mode)
        for (int blocks : {1, 148, 148 * 8}) {
            spin<<<blocks, 256>>>(d, mode ? 20000 : 4000000, mode);
            cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
            printf("mode %d blocks %4d: %lld cycles in %lld ns -> %.3f GHz\n", mode, blocks, h[0], h[1], (double)h[0] / h[1]);
        }
    return 0;
}
