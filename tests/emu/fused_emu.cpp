// Host emulation of the fused window-attention CTA body (TEST INFRASTRUCTURE).
// Compiles stratified_transformer_b200/csrc/fused_phases.cuh with FW_HOST_EMU: every barrier-separated phase
// becomes a loop over the thread ids of one CTA.  Lets `-m "not gpu"` tests check the index arithmetic, the shared
// memory aliasing and the math of the CUDA kernels against the CPU oracle.  Never loaded by the product package.
#define FW_HOST_EMU 1
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>
#include "../../stratified_transformer_b200/csrc/fused_phases.cuh"
#include "../../stratified_transformer_b200/csrc/fused_tc.cuh"

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

// ---- tensor-core version (fused_tc.cuh): TMEM and the MMAs are emulated functionally -------------------------------------
using namespace stb200::fwtc;

template <int BQ, int BK>
static void run_tc(const PassParams &P, int backward, int n_cta) {
    const TcLayout y = make_tc_layout(BQ, BK, P.Rpad, backward != 0);
    unsigned char *sm = (unsigned char *)aligned_alloc(128, y.total);
    float *tmem = (float *)malloc(sizeof(float) * 128 * TMEM_COLS);
    for (int head = 0; head < P.h; ++head)
        for (int cta = 0; cta < n_cta; ++cta) {
            for (uint32_t i = 0; i < y.total / 4; ++i) ((float *)sm)[i] = __builtin_nanf("");
            for (int i = 0; i < 128 * TMEM_COLS; ++i) tmem[i] = __builtin_nanf("");
            TcCtx ctx{tmem, sm};
            if (backward) {
                if (y.HR == 128) backward_cta_tc<BQ, BK, 128>(P, head, cta, n_cta, sm, ctx);
                else if (y.HR == 192) backward_cta_tc<BQ, BK, 192>(P, head, cta, n_cta, sm, ctx);
                else backward_cta_tc<BQ, BK, 256>(P, head, cta, n_cta, sm, ctx);
            } else {
                forward_cta_tc<BQ, BK>(P, head, cta, n_cta, sm, ctx);
            }
        }
    free(sm);
    free(tmem);
}

extern "C" int fw_emu_tc_smem_bytes(int BQ, int BK, int Rpad, int backward) { return (int)make_tc_layout(BQ, BK, Rpad, backward != 0).total; }

extern "C" int fw_emu_run_tc(const PassParams *P, int BQ, int BK, int backward, int n_cta) {
    if (P->Rpad % 16) return 2;
    if (BQ == 64 && BK == 64) run_tc<64, 64>(*P, backward, n_cta);
    else if (BQ == 48 && BK == 32) run_tc<48, 32>(*P, backward, n_cta);
    else if (BQ == 32 && BK == 32) run_tc<32, 32>(*P, backward, n_cta);
    else if (BQ == 16 && BK == 16) run_tc<16, 16>(*P, backward, n_cta);
    else if (BQ == 16 && BK == 8) run_tc<16, 8>(*P, backward, n_cta);
    else return 1;
    return 0;
}
