// Host emulation of the fused window-attention CTA body (TEST INFRASTRUCTURE).
// Compiles stratified_transformer_b200/csrc/fused_phases.cuh with FW_HOST_EMU: every barrier-separated phase
// becomes a loop over the thread ids of one CTA.  Lets `-m "not gpu"` tests check the index arithmetic, the shared
// memory aliasing and the math of the CUDA kernels against the CPU oracle.  Never loaded by the product package.
#define FW_HOST_EMU 1
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>
#include "../../stratified_transformer_b200/csrc/fused_phases.cuh"

using namespace stb200::fw;

template <int BQ, int BK>
static void run(const PassParams &P, int backward, int n_cta) {
    const Layout y = make_layout(BQ, BK, P.Rpad, backward != 0);
    float *sm = (float *)aligned_alloc(64, (size_t)y.total * sizeof(float));
    for (int head = 0; head < P.h; ++head)
        for (int cta = 0; cta < n_cta; ++cta) {
            for (int i = 0; i < y.total; ++i) sm[i] = __builtin_nanf("");   // uninitialised shared memory
            if (backward) backward_cta<BQ, BK>(P, head, cta, n_cta, sm);
            else forward_cta<BQ, BK>(P, head, cta, n_cta, sm);
        }
    free(sm);
}

extern "C" int fw_emu_smem_floats(int BQ, int BK, int Rpad, int backward) { return make_layout(BQ, BK, Rpad, backward != 0).total; }

extern "C" int fw_emu_run(const PassParams *P, int BQ, int BK, int backward, int n_cta) {
    if (BQ == 64 && BK == 64) run<64, 64>(*P, backward, n_cta);
    else if (BQ == 48 && BK == 32) run<48, 32>(*P, backward, n_cta);
    else if (BQ == 32 && BK == 32) run<32, 32>(*P, backward, n_cta);
    else if (BQ == 16 && BK == 16) run<16, 16>(*P, backward, n_cta);
    else if (BQ == 16 && BK == 8) run<16, 8>(*P, backward, n_cta);
    else return 1;
    return 0;
}
