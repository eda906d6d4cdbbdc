// Fused window attention, forward and backward: kernels around the CTA bodies of fused_phases.cuh and their C ABI.
//
// Replaces the pair path of WindowAttention.forward (/root/reference/model/stratified_transformer.py:183-210) and its
// backward for pair lists that come from the device builder (pair_builder.cu): one launch per pass
// (dense windows, key chunk 0, 1, ...; then sparse keys per large window, key chunk 0, 1, ...).  Persistent CTAs, each bound
// to one head (its three table slices stay in shared memory) and striding over the pass's work items.
#include "common.cuh"
#include "fused_phases.cuh"
#include "fused_tc.cuh"

#include <cstdlib>
#include <cstring>

namespace stb200 {

using fw::PassParams;

template <int BQ, int BK, bool BWD>
__global__ void __launch_bounds__(fw::NT, 1) fused_window_kernel(const PassParams P, int ctas_per_head) {
    extern __shared__ __align__(16) float fw_smem[];
    const int head = blockIdx.x / ctas_per_head, cta = blockIdx.x - head * ctas_per_head;
    if (BWD) fw::backward_cta<BQ, BK>(P, head, cta, ctas_per_head, fw_smem);
    else fw::forward_cta<BQ, BK>(P, head, cta, ctas_per_head, fw_smem);
}

constexpr size_t kMaxSmem = 232448;   // 227 KB opt-in limit per CTA on sm_100a
static long long *g_phase_prof = nullptr;   // development aid, see stb200_fused_phase_profile

// tensor-core version (fused_tc.cuh): 512 threads, 512 TMEM columns per CTA, one mbarrier for MMA completion
template <int BQ, int BK, int HRT, bool BWD>
__global__ void __launch_bounds__(fwtc::NTC, 1) fused_window_tc_kernel(const PassParams P, int ctas_per_head) {
    extern __shared__ __align__(128) unsigned char fw_smem_tc[];
    const fwtc::TcLayout y = fwtc::make_tc_layout(BQ, BK, P.Rpad, BWD);
    uint32_t *slot = reinterpret_cast<uint32_t *>(fw_smem_tc + y.slot);
    uint64_t *bar = reinterpret_cast<uint64_t *>(fw_smem_tc + y.slot + 8);
    const int head = blockIdx.x / ctas_per_head, cta = blockIdx.x - head * ctas_per_head;
    if (threadIdx.x < 32) tc::tmem_alloc(slot, fwtc::TMEM_COLS);
    if (threadIdx.x == 0) tc::mbar_init(bar, 1);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    fwtc::TcCtx ctx{*slot, tc::smem_u32(fw_smem_tc), bar, 0u};
    if (BWD) fwtc::backward_cta_tc<BQ, BK, HRT>(P, head, cta, ctas_per_head, fw_smem_tc, ctx);
    else fwtc::forward_cta_tc<BQ, BK>(P, head, cta, ctas_per_head, fw_smem_tc, ctx);
    tc::fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) tc::tmem_dealloc(ctx.tmem, fwtc::TMEM_COLS);
}

// STB200_FUSED_IMPL = fma: always the FMA kernels (A/B measurements); default: tensor-core kernels whenever they fit
static bool use_tc() {
    static int v = -1;
    if (v < 0) {
        const char *e = std::getenv("STB200_FUSED_IMPL");
        v = (e && !std::strcmp(e, "fma")) ? 0 : 1;
    }
    return v == 1;
}

template <int BQ, int BK, int HRT, bool BWD>
static int launch_pass_tc(const PassParams &P, const char *name, double bytes, cudaStream_t s) {
    const size_t smem = fwtc::make_tc_layout(BQ, BK, P.Rpad, BWD).total;
    static bool attr_set = false;
    if (!attr_set) {
        const cudaError_t e = cudaFuncSetAttribute(fused_window_tc_kernel<BQ, BK, HRT, BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem);
        STB200_REQUIRE(e == cudaSuccess, STB200_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        attr_set = true;
    }
    int per_head = kNumSMs / P.h;
    if (per_head < 1) per_head = 1;
    if (per_head > P.n_items) per_head = P.n_items;
    KernelScope ks(name, bytes, s);
    fused_window_tc_kernel<BQ, BK, HRT, BWD><<<per_head * P.h, fwtc::NTC, smem, s>>>(P, per_head);
    return STB200_OK;
}

template <int BQ, int BK, bool BWD>
static int launch_pass(const PassParams &P, const char *name, double bytes, cudaStream_t s) {
    const fw::Layout y = fw::make_layout(BQ, BK, P.Rpad, BWD);
    const size_t smem = (size_t)y.total * sizeof(float);
    STB200_REQUIRE(smem <= kMaxSmem, STB200_ERR_ARG, "fused window attention: %zu B of shared memory needed for blocks %dx%d with %d staged bins",
                   smem, BQ, BK, P.RB);
    static bool attr_set = false;   // per instantiation
    if (!attr_set) {
        const cudaError_t e = cudaFuncSetAttribute(fused_window_kernel<BQ, BK, BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem);
        STB200_REQUIRE(e == cudaSuccess, STB200_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        attr_set = true;
    }
    // one CTA per SM; CTAs of a head are consecutive
    int per_head = kNumSMs / P.h;
    if (per_head < 1) per_head = 1;
    if (per_head > P.n_items) per_head = P.n_items;
    KernelScope ks(name, bytes, s);
    fused_window_kernel<BQ, BK, BWD><<<per_head * P.h, fw::NT, smem, s>>>(P, per_head);
    return STB200_OK;
}

template <bool BWD>
static int dispatch_pass(const stb200_fused_pass &ps, PassParams &P, const char *name, double bytes, cudaStream_t s) {
    P.items = (const fw::Item *)ps.items;
    P.n_items = ps.n_items;
    P.q_order = ps.q_order; P.k_order = ps.k_order; P.rel = ps.rel;
    P.pos_win = ps.pos_win; P.wstart = ps.wstart; P.tile_base = ps.tile_base;
    P.bin_lo = ps.bin_lo; P.RB = ps.RB;
    { const char *e = std::getenv("STB200_FUSED_DBG"); P.dbg = e ? std::atoi(e) : 0; }
    P.prof = g_phase_prof;
    P.Rpad = (3 * ps.RB + 15) / 16 * 16;
    STB200_REQUIRE(ps.RB > 0 && P.Rpad <= 256, STB200_ERR_ARG, "fused window attention stages at most 85 bins per axis (got %d)", ps.RB);
    if (use_tc()) {   // tensor-core kernels where their shared-memory plan fits (S3DIS table length: both passes)
        const int hr = fwtc::hist_rows(P.Rpad);
        if (ps.BQ == 64 && ps.BK == 64 && hr == 128 && fwtc::make_tc_layout(64, 64, P.Rpad, BWD).total <= kMaxSmem)
            return launch_pass_tc<64, 64, 128, BWD>(P, name, bytes, s);
        if (ps.BQ == 48 && ps.BK == 32 && hr == 192 && fwtc::make_tc_layout(48, 32, P.Rpad, BWD).total <= kMaxSmem)
            return launch_pass_tc<48, 32, 192, BWD>(P, name, bytes, s);
    }
    if (ps.BQ == 64 && ps.BK == 64) return launch_pass<64, 64, BWD>(P, name, bytes, s);
    if (ps.BQ == 48 && ps.BK == 32) return launch_pass<48, 32, BWD>(P, name, bytes, s);
    if (ps.BQ == 32 && ps.BK == 32) return launch_pass<32, 32, BWD>(P, name, bytes, s);
    set_error("fused window attention: no kernel for blocks %dx%d (built: 64x64, 48x32, 32x32)", ps.BQ, ps.BK);
    return STB200_ERR_ARG;
}

}  // namespace stb200

using namespace stb200;

extern "C" {

/* development aid: device buffer of 64 int64 that CTA 0 / head 0 of every later tcgen05 fused launch adds its per-phase cycle
 * counts to (NULL switches it off); tools/fused_phase_prof.py */
void stb200_fused_phase_profile(long long *device_buffer) { g_phase_prof = device_buffer; }

int stb200_fused_attention_forward(const stb200_fused_pass *passes, int n_passes, int N, int h, int L, const float *q, const float *k,
                                   const float *v, const float *table_q, const float *table_k, const float *table_v, float *out,
                                   float *m, float *l, void *stream) {
    STB200_REQUIRE(passes && n_passes > 0 && N > 0 && h > 0 && L > 0, STB200_ERR_ARG, "bad sizes");
    STB200_REQUIRE(q && k && v && table_q && table_k && table_v && out && m && l, STB200_ERR_ARG, "null pointer");
    PassParams P{};
    P.L = L; P.h = h;
    P.q = q; P.k = k; P.v = v; P.tq = table_q; P.tk = table_k; P.tv = table_v;
    P.out = out; P.m = m; P.l = l;
    for (int i = 0; i < n_passes; ++i) {
        if (passes[i].n_items <= 0) continue;
        // fused accounting: q, k, v read + out written once per point, spread over the passes by their share of rows
        const double bytes = 4.0 * N * h * 16 * 4 / n_passes;
        const int rc = dispatch_pass<false>(passes[i], P, passes[i].pos_win ? "fused_fwd[dense]" : "fused_fwd[sparse]", bytes, (cudaStream_t)stream);
        if (rc) return rc;
    }
    return check_launch("fused_attention_forward");
}

int stb200_fused_attention_backward(const stb200_fused_pass *passes, int n_passes, int N, int h, int L, const float *grad_out,
                                    const float *out, const float *lse, const float *q, const float *k, const float *v,
                                    const float *table_q, const float *table_k, const float *table_v, float *grad_q, float *grad_k,
                                    float *grad_v, float *grad_table_q, float *grad_table_k, float *grad_table_v, void *stream) {
    STB200_REQUIRE(passes && n_passes > 0 && N > 0 && h > 0 && L > 0, STB200_ERR_ARG, "bad sizes");
    STB200_REQUIRE(grad_out && out && lse && q && k && v && table_q && table_k && table_v && grad_q && grad_k && grad_v && grad_table_q &&
                       grad_table_k && grad_table_v, STB200_ERR_ARG, "null pointer");
    PassParams P{};
    P.L = L; P.h = h;
    P.q = q; P.k = k; P.v = v; P.tq = table_q; P.tk = table_k; P.tv = table_v;
    P.out = const_cast<float *>(out); P.g = grad_out; P.lse = lse;
    P.gq = grad_q; P.gk = grad_k; P.gv = grad_v; P.gtq = grad_table_q; P.gtk = grad_table_k; P.gtv = grad_table_v;
    for (int i = 0; i < n_passes; ++i) {
        if (passes[i].n_items <= 0) continue;
        const double bytes = 4.0 * N * h * 16 * 8 / n_passes;
        const int rc = dispatch_pass<true>(passes[i], P, passes[i].pos_win ? "fused_bwd[dense]" : "fused_bwd[sparse]", bytes, (cudaStream_t)stream);
        if (rc) return rc;
    }
    return check_launch("fused_attention_backward");
}

}  // extern "C"
