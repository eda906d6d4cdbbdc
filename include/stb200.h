/* stb200.h — C ABI of the B200-native window-attention hot path (libstb200.so).
 *
 * Drop-in boundary for the `extern "C"` launchers of the reference's pointops2 extension
 * (paths under /root/reference/lib/pointops2/src).  Every entry point cites the reference
 * declaration it replaces.  Conventions shared by all entry points:
 *
 *   - plain C: device pointers + sizes, no torch types.  float = fp32, int = int32, contiguous.
 *   - trailing `void *stream` is a cudaStream_t (NULL = legacy default stream, what the
 *     reference always used: `<<<blocks, threads, 0>>>`).  Calls are asynchronous.
 *   - return value: 0 on success, an STB200_ERR_* code otherwise (the reference returned void and
 *     threw a `const char*` for unsupported head dims); stb200_last_error() gives the message.
 *   - outputs are caller-allocated like in the reference (functions/pointops.py:157,188-189,...).
 *     Outputs that the reference overwrites are overwritten; outputs it accumulates into
 *     (grad_k, grad_v, grad_table*) are ACCUMULATED INTO here too, so callers keep zero-filling
 *     them exactly as functions/pointops.py does.
 *   - `n_max` is accepted for signature compatibility and ignored (no launch shape depends on it,
 *     so segments longer than 1024 pairs are fine).
 *   - head dim (C/h or hdim) must be 16 or 32, as in the reference (attention_cuda_kernel_v2.cu:108-117).
 *   - rel_idx values must lie in [0, L); out-of-range values are clamped for memory safety
 *     (the reference reads out of bounds; its Python asserts the range, stratified_transformer.py:189-190).
 *
 * Additions relative to the reference ABI (needed by a scatter-free design, documented in DESIGN.md):
 *   - `L` (table length) on every entry point that takes a table: tables are staged in shared memory.
 *   - the "transposed CSR" (pairs grouped by key) on backward entry points: grad_k / grad_v are
 *     computed by gathering over a key's incoming pairs instead of float atomics.  Build it once per
 *     index set with stb200_transpose_csr().
 */
#ifndef STB200_H
#define STB200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STB200_OK 0
#define STB200_ERR_HEAD_DIM 1   /* head dim not in {16, 32} */
#define STB200_ERR_ARG 2        /* null pointer / negative size / inconsistent sizes */
#define STB200_ERR_CUDA 3       /* a CUDA runtime call failed; see stb200_last_error() */
#define STB200_ERR_WORKSPACE 4  /* workspace too small; query the size with the *_workspace_bytes call */

const char *stb200_last_error(void);
/* number of kernels this library launched since load (bench.py's "gpu_launches") */
long long stb200_launch_count(void);
/* ABI version: 100 = first release, 101 = stb200_index has len_order / t_len_order (append-only struct growth),
 * 102 = fused work plan + fused window attention entry points; 103 = stb200_qkv_split / _merge, layer norm, pre-step (batch vector, ball query), *_ws variants of FPS / kNN */
int stb200_version(void);
/* Optional per-kernel profiler: when enabled every launch is bracketed by CUDA events on its stream.
 * stb200_profile_dump writes a JSON object {"kernel name": {"launches", "ms", "bytes"}} (bytes = algorithmic bytes
 * of those launches, DESIGN.md "Roofline accounting") into buf, clears the records, returns the size needed. */
void stb200_profile_enable(int on);
size_t stb200_profile_dump(char *buf, size_t cap);

/* ------------------------------------------------------------------------------------------------
 * Transposed CSR: pairs grouped by key.  For key j, t in [t_offsets[j], t_offsets[j+1]) enumerates its
 * incoming pairs in ascending pair id: t_pair[t] = m, t_index0[t] = query id of pair m.
 * (No reference counterpart: the reference scatters with atomicAdd, e.g. attention_cuda_kernel_v2.cu:84.)
 * workspace: stb200_transpose_csr_workspace_bytes(N, M) bytes of device scratch. */
size_t stb200_transpose_csr_workspace_bytes(int N, int M);
int stb200_transpose_csr(int N, int M, const int *index0_offsets, const int *index1,
                         int *t_offsets /*[N+1]*/, int *t_pair /*[M]*/, int *t_index0 /*[M]*/,
                         void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------------
 * attention_step1 v2 — replaces attention_step1_forward_cuda_launcher_v2 /
 * attention_step1_backward_cuda_launcher_v2 (attention_v2/attention_cuda_kernel_v2.h:18-19).
 * attn[m,h] = <q[i0(m),h,:], k[index1[m],h,:]>,  m in [index0_offsets[n], index0_offsets[n+1]). */
int stb200_attention_step1_forward_v2(int N, int M, int h, int C, unsigned int n_max,
                                      const float *q, const float *k, const int *index0_offsets,
                                      const int *index1, float *attn, void *stream);
int stb200_attention_step1_backward_v2(int N, int M, int h, int C, unsigned int n_max,
                                       const float *grad_out, const int *index0_offsets, const int *index1,
                                       const float *q, const float *k, float *grad_q, float *grad_k,
                                       const int *t_offsets, const int *t_pair, const int *t_index0,
                                       void *stream);

/* dot_prod_with_idx v3 — replaces dot_prod_with_idx_forward_cuda_launcher_v3 / ..._backward_..._v3
 * (rpe_v2/relative_pos_encoding_cuda_kernel_v2.h:23-24).
 * out[m,h] = <q[i0],Eq(m,h)> + <k[i1],Ek(m,h)>, E_t = T_t[r0,h,:,0] + T_t[r1,h,:,1] + T_t[r2,h,:,2]. */
int stb200_dot_prod_with_idx_forward_v3(int N, int M, int h, int hdim, int n_max, int L,
                                        const float *q, const int *index_q_offsets, const float *k,
                                        const int *index_k, const float *table_q, const float *table_k,
                                        const int *rel_idx, float *output, void *stream);
int stb200_dot_prod_with_idx_backward_v3(int N, int M, int h, int hdim, int n_max, int L,
                                         const float *grad_out, const float *q, const int *index_q_offsets,
                                         const float *k, const int *index_k, const float *table_q,
                                         const float *table_k, const int *rel_idx, float *grad_q,
                                         float *grad_k, float *grad_table_q, float *grad_table_k,
                                         const int *t_offsets, const int *t_pair, const int *t_index0,
                                         void *stream);

/* attention_step2_with_rel_pos_value v2 — replaces attention_step2_with_rel_pos_value_forward_cuda_launcher_v2
 * / ..._backward_..._v2 (rpe_v2/relative_pos_encoding_cuda_kernel_v2.h:26-27).
 * out[n,h,:] = sum_seg attn[m,h] * (v[index1[m],h,:] + Ev(m,h,:)). */
int stb200_attention_step2_with_rel_pos_value_forward_v2(int N, int M, int h, int hdim, int n_max, int L,
                                                         const float *attn, const float *v,
                                                         const int *index0_offsets, const int *index1,
                                                         const float *table, const int *rel_idx,
                                                         float *output, void *stream);
int stb200_attention_step2_with_rel_pos_value_backward_v2(int N, int M, int h, int hdim, int n_max, int L,
                                                          const float *grad_out, const int *index0_offsets,
                                                          const int *index1, const float *attn, const float *v,
                                                          const float *table, const int *rel_idx,
                                                          float *grad_attn, float *grad_v, float *grad_table,
                                                          const int *t_offsets, const int *t_pair,
                                                          const int *t_index0, void *stream);

/* Segment softmax over the pairs of each query, per head — replaces the third-party
 * torch_scatter.scatter_softmax(src=attn_flat, index=index_0, dim=0) call and the preceding
 * `attn_flat + relative_position_bias` (model/stratified_transformer.py:203,205).
 * p = softmax_seg(a + b); b may be NULL.  Backward: gs = p * (gp - sum_seg p*gp). */
int stb200_segment_softmax_forward(int N, int M, int h, const float *a, const float *b,
                                   const int *index0_offsets, float *p, void *stream);
int stb200_segment_softmax_backward(int N, int M, int h, const float *p, const float *grad_p,
                                    const int *index0_offsets, float *grad_s, void *stream);

/* ------------------------------------------------------------------------------------------------
 * v1 entry points (explicit, possibly unsorted index0/index1; API completeness).
 * Replace attention/attention_cuda_kernel.h:17-21 and rpe/relative_pos_encoding_cuda_kernel.h launchers. */
int stb200_attention_step1_forward(int N, int M, int h, int C, const float *q, const float *k,
                                   const int *index0, const int *index1, float *attn, void *stream);
int stb200_attention_step1_backward(int N, int M, int h, int C, const float *grad_out, const int *index0,
                                    const int *index1, const float *q, const float *k, float *grad_q,
                                    float *grad_k, void *stream);
int stb200_attention_step2_forward(int N, int M, int h, int C, const float *attn, const float *v,
                                   const int *index0, const int *index1, float *output, void *stream);
int stb200_attention_step2_backward(int N, int M, int h, int C, const float *grad_out, const int *index0,
                                    const int *index1, const float *attn, const float *v, float *grad_attn,
                                    float *grad_v, void *stream);
int stb200_dot_prod_with_idx_forward(int N, int M, int h, int hdim, int L, const float *q, const int *index,
                                     const float *table, const int *rel_idx, float *output, void *stream);
int stb200_dot_prod_with_idx_backward(int N, int M, int h, int hdim, int L, const float *grad_out,
                                      const float *q, const int *index, const float *table,
                                      const int *rel_idx, float *grad_q, float *grad_table, void *stream);
int stb200_attention_step2_with_rel_pos_value_forward(int N, int M, int h, int hdim, int L, const float *attn,
                                                      const float *v, const int *index0, const int *index1,
                                                      const float *table, const int *rel_idx, float *output,
                                                      void *stream);
int stb200_attention_step2_with_rel_pos_value_backward(int N, int M, int h, int hdim, int L,
                                                       const float *grad_out, const int *index0,
                                                       const int *index1, const float *attn, const float *v,
                                                       const float *table, const int *rel_idx,
                                                       float *grad_attn, float *grad_v, float *grad_table,
                                                       void *stream);

/* ------------------------------------------------------------------------------------------------
 * Furthest point sampling — replaces furthestsampling_cuda_launcher (sampling/sampling_cuda_kernel.h:10).
 * Bit-exact with the reference kernel including its tie behaviour.  `n` = largest scene size (selects the
 * reference's virtual block size that defines tie order); `tmp` [N] f32 scratch is accepted and left
 * untouched (the running minimum distances live on chip); idx [new_offset[b-1]] int32 out. */
int stb200_furthestsampling(int b, int n, const float *xyz, const int *offset, const int *new_offset,
                            float *tmp, int *idx, void *stream);
/* The same sampling, bit for bit, with exact bounding-box pruning of the distance update (csrc/fps.cu, fps_pruned_kernel): the points
 * of a scene are Morton-sorted into groups of 256 and a group is skipped while the new sample is at least as far from its box as
 * its largest running minimum.  N = total number of points (all scenes); workspace from stb200_fps_workspace_bytes(N, b).
 * Opt-in (STB200_FPS_PRUNE=1 in the environment; measured slower than the plain kernel on B200); without it,
 * with workspace NULL or for scenes below 3072 points: identical to stb200_furthestsampling. */
size_t stb200_fps_workspace_bytes(int N, int b);
int stb200_furthestsampling_ws(int b, int n, int N, const float *xyz, const int *offset, const int *new_offset, float *tmp, int *idx,
                               void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Pair-index construction (new: the reference does this in Python with torch ops + torch_geometric's
 * voxel_grid, model/stratified_transformer.py:10-65, 267-317).  Two calls because the caller allocates
 * index_1 / rel_idx and therefore has to learn M in between:
 *
 *   1. stb200_stratified_pairs_count: window partition for one block parity (0: unshifted, 1: shifted by
 *      w/2), sampled keys from `downsample_idx` (m = 0 -> dense pairs only, the Swin variant), per-query
 *      key counts -> index0_offsets [N+1] and totals[4] = {M, n_max, error flag, n_windows} (device memory).
 *      totals[3] = number of small windows.  Intermediate state stays in `workspace` (stb200_pair_builder_workspace_bytes(N) bytes).
 *   2. stb200_stratified_pairs_fill: emits index_1 [M] (per query: dense keys ascending by point id, then
 *      sparse keys ascending by point id), rel_idx [M,3] (may be NULL) and index_0 [M] (may be NULL) from the
 *      same workspace.  window_size_x2 = (float)(2.0 * window_size), quant_size as fp32 — the scalars of
 *      model/stratified_transformer.py:188.
 * offset: cumulative point counts per scene [b] (int32), xyz [N,3] fp32. */
size_t stb200_pair_builder_workspace_bytes(int N);
int stb200_stratified_pairs_count(int N, int b, const float *xyz, const int *offset, float window_size, int parity,
                                  const int *downsample_idx, int m, void *workspace, size_t workspace_bytes,
                                  int *index0_offsets, int *totals, void *stream);
int stb200_stratified_pairs_fill(int N, const float *xyz, float window_size_x2, float quant_size, int has_sparse,
                                 void *workspace, size_t workspace_bytes, const int *index0_offsets, int *index_1,
                                 int *rel_idx, int *index_0, int *row_order /* [N] points sorted by window, may be NULL */,
                                 int *win_offsets /* [n_win+1] window boundaries in row_order, may be NULL */, int n_win,
                                 int M, void *stream);

/* Which torch device's arithmetic the Stratified rel-pos index reproduces bit for bit.  1 (default): CUDA tensors, i.e. the
 * reference as it runs -- `round(r * 1e5) / 1e5` is a multiplication by the fp32 reciprocal there; 0: CPU tensors (IEEE
 * division), for fixtures generated with CPU torch.  Applies to every later call of this process (pair builder, fused plan,
 * stb200_rel_pos_index_stratified); synchronises the device. */
int stb200_set_torch_semantics(int cuda);

/* Relative-position index of an existing CSR pair list.
 * Stratified: idx = ((round((xyz[i0]-xyz[i1])*1e5)/1e5) + 2w - 1e-4) // quant  (stratified_transformer.py:186-188)
 * Swin:       xq = ((xyz - min + shift) % w) // quant; idx = xq[i0] - xq[i1] + qgl - 1 (swin3d_transformer.py:151-154)
 * both with torch's fp32 semantics (round-half-even, fmod-based floor division). */
int stb200_rel_pos_index_stratified(int N, const float *xyz, const int *index0_offsets, const int *index_1,
                                    float window_size_x2, float quant_size, int *rel_idx, void *stream);
int stb200_rel_pos_index_swin(int N, const float *xyz, const int *index0_offsets, const int *index_1,
                              float window_size, float quant_size, float shift_size, int quant_grid_length,
                              float *xq_scratch, unsigned *mm_scratch, int *rel_idx, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Fused entry points (new; no reference counterpart).  The reference's WindowAttention issues attention_step1_v2 and
 * dot_prod_with_idx_v3 back to back on the same pairs (model/stratified_transformer.py:183,194) and adds the results
 * (:203); both gather the same key rows.  stb200_window_logits_* does q.k + rel-pos bias in one pass (and the matching
 * single-pass gradients); stb200_window_aggregate_* is attention_step2_with_rel_pos_value_v2 driven by the same index
 * descriptor.  rel_packed / t_rel_packed (optional, stb200_pack_rel) replace the 12-byte rel_idx rows by one 32-bit
 * word per pair (three 10-bit bins) in query-segment / key-segment order.
 * Output convention of the fused entry points: grad_q, grad_k, grad_v, grad_attn, logits, output are OVERWRITTEN (no
 * zero-fill needed); only the table gradients are accumulated into (caller zero-fills them, they are a few KB). */
typedef struct stb200_index {
    int N, M;
    const int *index0_offsets; /* [N+1] */
    const int *index1;         /* [M]   */
    const int *rel_idx;        /* [M,3] (may be NULL when rel_packed is given and no transposed kernel needs it) */
    const int *t_offsets;      /* [N+1] transposed CSR, backward only */
    const int *t_pair;         /* [M]   */
    const int *t_index0;       /* [M]   */
    const unsigned *rel_packed;   /* [M] optional */
    const unsigned *t_rel_packed; /* [M] optional, bins of pair t_pair[t] */
    const int *row_order;         /* [N] optional: process rows in this order (e.g. points sorted by window, from the
                                     pair builder) so that neighbouring warps gather the same k/v rows */
    const int *len_order;         /* [N] optional: queries sorted by pair count (stb200_length_order on index0_offsets) */
    const int *t_len_order;       /* [N] optional: keys sorted by incoming pair count (stb200_length_order on t_offsets);
                                     both are balance hints for the table-gradient kernels, whose 32-row tiles give one
                                     lane per row: rows of equal length keep all lanes busy */
} stb200_index;

/* order[] = the rows listed in base_order (NULL: 0..N-1) stably sorted by pair count offsets[r+1]-offsets[r]
 * (counts above 65535 compare equal: the order is a balance hint, any permutation is valid input for the kernels). */
size_t stb200_length_order_workspace_bytes(int N);
int stb200_length_order(int N, const int *offsets, const int *base_order, int *order, void *workspace,
                        size_t workspace_bytes, void *stream);

/* out[i] = r0 | r1 << 10 | r2 << 20 of pair (perm ? perm[i] : i), bins clamped to [0, L) */
int stb200_pack_rel(int M, int L, const int *rel_idx, const int *perm, unsigned *out, void *stream);
int stb200_window_logits_forward(const stb200_index *ix, int h, int hdim, int L, const float *q, const float *k,
                                 const float *table_q, const float *table_k, float *logits, void *stream);
int stb200_window_logits_backward(const stb200_index *ix, int h, int hdim, int L, const float *grad_logits,
                                  const float *q, const float *k, const float *table_q, const float *table_k,
                                  float *grad_q, float *grad_k, float *grad_table_q, float *grad_table_k, void *stream);
/* Same, with scratch space: when `workspace` holds stb200_window_logits_backward_workspace_bytes(M, h) bytes, grad_logits
 * is first brought into transposed-CSR order so the two key-side kernels stream it instead of gathering through t_pair. */
size_t stb200_window_logits_backward_workspace_bytes(int M, int h);
int stb200_window_logits_backward_ws(const stb200_index *ix, int h, int hdim, int L, const float *grad_logits,
                                     const float *q, const float *k, const float *table_q, const float *table_k,
                                     float *grad_q, float *grad_k, float *grad_table_q, float *grad_table_k,
                                     void *workspace, size_t workspace_bytes, void *stream);
int stb200_window_aggregate_forward(const stb200_index *ix, int h, int hdim, int L, const float *attn, const float *v,
                                    const float *table_v, float *output, void *stream);
int stb200_window_aggregate_backward(const stb200_index *ix, int h, int hdim, int L, const float *grad_out,
                                     const float *attn, const float *v, const float *table_v, float *grad_attn,
                                     float *grad_v, float *grad_table_v, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Fused window-attention forward: logits + rel-pos bias + softmax + aggregation per window on the
 * tensor cores, no M-sized intermediate besides the probabilities `attn` [M,h] that the backward pass consumes.
 * stb200_classify_windows marks the windows whose queries all share one key list of at most stb200_fused_max_keys()
 * keys (flags[w] = 1) and lists the rows of all other windows (fallback_rows[0..*fallback_count)); the caller runs
 * the per-pair entry points on those rows (stb200_index.row_order = fallback_rows, N = count).  Requires
 * rel_packed, row_order (points sorted by window) and win_offsets from the pair builder; head dim 16, L <= 85. */
int stb200_fused_max_keys(void);
int stb200_classify_windows(int n_win, const int *win_offsets, const int *row_order, const int *index0_offsets,
                            const int *index1, unsigned char *flags, int *fallback_rows, int *fallback_count, void *stream);
int stb200_window_attention_forward_fused(const stb200_index *ix, int n_win, const int *win_offsets,
                                          const unsigned char *win_flags, int h, int hdim, int L, const float *q,
                                          const float *k, const float *v, const float *table_q, const float *table_k,
                                          const float *table_v, float *output, float *attn, void *stream);
/* segment softmax restricted to a list of rows (rows = NULL: all N rows) */
int stb200_segment_softmax_forward_rows(int n_rows, const int *rows, int h, const float *a, const float *b,
                                        const int *index0_offsets, float *p, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Window-centric fused attention (no reference counterpart as an op): the whole pair path of
 * WindowAttention.forward (model/stratified_transformer.py:183-210: attention_step1_v2 + dot_prod_with_idx_v3 + add +
 * scatter_softmax + attention_step2_with_rel_pos_value_v2) and its backward, without any [M,h] tensor.  It consumes a
 * WORK PLAN instead of the CSR pair list: the same pairs (get_indice_pairs, :10-42) expressed as
 *   dense tiles  : per small window the [n x n] matrix of packed rel-pos bins (r0 | r1<<8 | r2<<16), queries x keys;
 *   sparse tiles : per large window [n_V x n_s] (queries = its points, keys = its FPS-sampled points), bit 31 set where
 *                  the reference drops the pair (equal window_coord, :28-35);
 *   items        : blocks of at most BQ x BK of those matrices (struct of 8 ints: q_pos, nq, k_pos, nk, rel_off,
 *                  rel_pitch, flags, pad), grouped by key-chunk ordinal; one pass (= one launch) per ordinal.
 * Build it right after stb200_stratified_pairs_count on the same builder workspace (the CSR arrays are not needed):
 *   stb200_fused_plan_count -> totals[40] (device; ints): [0] dense rel words, [1] sparse rel words, [2] large windows,
 *       [3] largest small window, [4] most sampled keys of a large window, [5] error bits (1/2: a window needs more than 8
 *       key chunks, 4: more than 2^31 rel words), [6],[7] min / max dense bin (after fill), [8..15] dense items per ordinal,
 *       [16..23] sparse items per ordinal;   the caller reads it, allocates, then calls
 *   stb200_fused_plan_fill.  swin != 0: rel-pos index of model/swin3d_transformer.py:151-154 (dense only).
 * Dense blocks are square (BQ == BK).  Kernels are built for blocks 64x64, 48x32 and 32x32; head dim 16. */
typedef struct stb200_fused_pass {
    const void *items;       /* device, n_items x 8 ints, all of one ordinal */
    int n_items;
    const int *q_order;      /* sorted position -> point id of the query rows (dense: by small window; sparse: by large window) */
    const int *k_order;      /* ... of the key rows (dense: same array; sparse: sampled points grouped by large window) */
    const unsigned *rel;     /* tiles */
    const int *pos_win;      /* dense only (NULL marks a sparse pass): small-window rank of a sorted position [N] */
    const int *wstart;       /* dense only: window boundaries [n_win+1] */
    const int *tile_base;    /* dense only: first rel word of every window's tile [n_win] */
    int bin_lo, RB;          /* bins [bin_lo, bin_lo+RB) of every axis are staged (products are only computed for those) */
    int BQ, BK;
} stb200_fused_pass;

size_t stb200_fused_plan_scratch_bytes(int N);
int stb200_fused_plan_count(int N, void *builder_workspace, size_t builder_workspace_bytes, int has_sparse, int BQ, int BK, int BQS,
                            int BKS, void *scratch, size_t scratch_bytes, int *totals, void *stream);
int stb200_fused_plan_fill(int N, const float *xyz, float window_size_x2, float quant_size, int has_sparse, int BQ, int BK, int BQS,
                           int BKS, int swin, float swin_window, float swin_shift, void *builder_workspace,
                           size_t builder_workspace_bytes, void *scratch, size_t scratch_bytes, int *totals, unsigned *dense_rel,
                           int *tile_base, int *pos_win, int *order_s, int *wstart_s, void *dense_items, unsigned *sparse_rel,
                           int *order_l, int *samp, int n_samp, void *sparse_items, void *stream);
/* Forward: passes in launch order (dense ordinals, then sparse ordinals).  out [N,h,16]; m, l [N,h] scratch: after the call m
 * holds the log-sum-exp of every (query, head) row, which the backward pass takes as `lse`.  Nothing needs zero-filling. */
int stb200_fused_attention_forward(const stb200_fused_pass *passes, int n_passes, int N, int h, int L, const float *q, const float *k,
                                   const float *v, const float *table_q, const float *table_k, const float *table_v, float *out,
                                   float *m, float *l, void *stream);
/* Backward: grad_q is overwritten; grad_k / grad_v are overwritten by the dense pass except for windows larger than BK, whose
 * key rows are accumulated into (zero-fill grad_k / grad_v when totals[3] > BK); table gradients are accumulated into. */
int stb200_fused_attention_backward(const stb200_fused_pass *passes, int n_passes, int N, int h, int L, const float *grad_out,
                                    const float *out, const float *lse, const float *q, const float *k, const float *v,
                                    const float *table_q, const float *table_k, const float *table_v, float *grad_q, float *grad_k,
                                    float *grad_v, float *grad_table_q, float *grad_table_k, float *grad_table_v, void *stream);

/* Self-test of the tcgen05 building blocks the fused kernels use (csrc/tc_umma.cuh): one CTA computes a 3xTF32 GEMM through
 * tensor memory.  mode 0: A [M,K], B [N,K] -> A B^T (both K-major); mode 1: A [K,M], B [K,N] -> A^T B (both MN-major);
 * mode 2: A [M,K] (K-major), B [K,N] (MN-major) -> A B.  M in {64, 128}, N, K multiples of 8.  out [128, N] receives TMEM
 * lanes 0..127 of the accumulator; *status becomes 1 when the MMA never signalled completion. */
int stb200_tc_selftest(int mode, int M, int N, int K, const float *A, const float *B, float *out, int *status, void *stream);

/* ------------------------------------------------------------------------------------------------
 * The elementwise passes of WindowAttention.forward around the pair ops (model/stratified_transformer.py:172-175 and the
 * `.float()` casts at lines 193-216), one kernel each way.
 * stb200_qkv_split: qkv [N, 3C] (dtype 0 = fp32, 1 = bf16, 2 = fp16; the output of the projection GEMM WITHOUT its bias)
 *     -> q, k, v fp32 [N, C] contiguous (= [N, h, C/h]), bias [3C] added (NULL: none).  The caller folds `scale` into the q rows
 *     of the weight and bias.
 * stb200_qkv_merge: grad_q, grad_k, grad_v fp32 [N, C] -> grad_qkv [N, 3C] in `dtype`, and, when bias_partial != NULL,
 *     per-CTA column sums bias_partial [stb200_qkv_partial_rows(N, C), 3C] whose sum over rows is the bias gradient
 *     (deterministic: no atomics).  C must be a multiple of 8, 3C <= 2048. */
int stb200_qkv_partial_rows(int N, int C);
int stb200_qkv_split(int N, int C, int dtype, const void *qkv, const float *bias, float *q, float *k, float *v, void *stream);
int stb200_qkv_merge(int N, int C, int dtype, const float *grad_q, const float *grad_k, const float *grad_v, void *grad_qkv,
                     float *bias_partial, void *stream);

/* nn.LayerNorm(C) over the last dimension of fp32 [N, C], 1 <= C <= 384 (model/stratified_transformer.py:227,233 and the norms of
 * TransitionDown / Upsample): y = (x - mean) * rstd * gamma + beta, biased variance, rstd = rsqrt(var + eps); mean / rstd [N] are kept
 * for the backward pass.  Backward: grad_x, and per-CTA partial sums partial [stb200_layer_norm_partial_rows(N, C), 2C] whose sum over
 * rows is (grad_gamma | grad_beta) (NULL: not wanted).  gamma / beta may be NULL (no affine part). */
int stb200_layer_norm_partial_rows(long long N, int C);
int stb200_layer_norm_forward(long long N, int C, float eps, const float *x, const float *gamma, const float *beta, float *y, float *mean,
                              float *rstd, void *stream);
int stb200_layer_norm_backward(long long N, int C, const float *grad_y, const float *x, const float *gamma, const float *mean,
                               const float *rstd, float *grad_x, float *partial, void *stream);

/* Neighbourhood aggregation of the KPConv stem (torch_points3d KPConvLayer: rigid, linear influence, sum; third party, outside the
 * hot path): weighted[i, k, :] = sum_j max(0, 1 - |s_xyz[nbr[i,j]] - q_xyz[i] - kpts[k]| / extent) * feats[nbr[i,j], :], nbr int64
 * [n, nn] with entries < 0 or >= n_sup meaning "no neighbour"; K <= 16 kernel points, C <= 16 channels.  The projection
 * out = sum_k weighted[:, k, :] W_k stays a library GEMM.  Backward accumulates into grad_feats [n_sup, C] (zero it first). */
int stb200_kpconv_weighted(int n, int n_sup, int nn, int K, int C, float extent, const float *q_xyz, const float *s_xyz,
                           const long long *nbr, const float *kpts, const float *feats, float *weighted, void *stream);
int stb200_kpconv_weighted_backward(int n, int n_sup, int nn, int K, int C, float extent, const float *q_xyz, const float *s_xyz,
                                    const long long *nbr, const float *kpts, const float *grad_weighted, float *grad_feats, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Host pre-step of the training loop on the device — replaces train.py:319-325.
 * stb200_batch_from_offset: batch[i] = scene of point i (the reference builds it with a Python list per scene); offset = cumulative
 *     scene ends [b], int32; batch int64 [N] like the reference's `.long()`.
 * stb200_ball_query: tp.ball_query(radius, max_num, x, y, mode="partial_dense", batch_x, batch_y) of torch_points_kernels (third
 *     party, absent; its CPU path keeps max_num matches in kd-tree traversal order = unspecified).  Here: for query y[i] the
 *     support points x[j] of the same scene with d^2 < radius^2 (fp32, (dx*dx + dy*dy) + dz*dz), ordered by (d^2, j), the first
 *     max_num (<= 64) of them in idx [Ny, max_num] (int64, -1 padded) and dist2 [Ny, max_num] (-1 padded; may be NULL).
 *     batch_x / batch_y NULL: one scene. */
int stb200_batch_from_offset(int N, int b, const int *offset, long long *batch, void *stream);
size_t stb200_ball_query_workspace_bytes(int Nx);
int stb200_ball_query(int Nx, int Ny, float radius, int max_num, const float *x, const float *y, const long long *batch_x,
                      const long long *batch_y, void *workspace, size_t workspace_bytes, long long *idx, float *dist2, void *stream);

/* ------------------------------------------------------------------------------------------------
 * k nearest neighbours per scene — replaces knnquery_cuda_launcher
 * (knnquery/knnquery_cuda_kernel.h).  `b` = number of scenes (the reference walks new_offset until it finds the
 * query's scene).  idx [m, nsample] int32 ascending by squared distance, dist2 [m, nsample] squared distances;
 * identical to the reference including the order of equal distances.  nsample <= 100 (the reference's heap size). */
int stb200_knnquery(int m, int b, int nsample, const float *xyz, const float *new_xyz, const int *offset,
                    const int *new_offset, int *idx, float *dist2, void *stream);
/* The same result through a grid-pruned search (support points sorted by (scene, cell); a query is settled from the cells around
 * it when its k+1 nearest candidates are closer than the searched cube's faces and strictly ordered - then the answer is unique)
 * with the exact heap scan above for whatever remains (ties, sparse neighbourhoods, scenes of at most k points, nsample > 32).
 * n = number of support points; workspace from stb200_knnquery_workspace_bytes(n, m, b) (NULL: heap scan only). */
size_t stb200_knnquery_workspace_bytes(int n, int m, int b);
int stb200_knnquery_ws(int n, int m, int b, int nsample, const float *xyz, const float *new_xyz, const int *offset,
                       const int *new_offset, int *idx, float *dist2, void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------------
 * bf16-storage forward path (inference, BASELINE config 3): q / k / v are bf16 [N,h,16] (device pointers to
 * __nv_bfloat16), the fp32 tables are rounded to bf16 while being staged, products and sums are fp32, logits and
 * output are fp32.  Halves the gathered row bytes and the shared-memory bytes per table look-up.  Stated tolerance
 * against the fp32 path: 2e-2 of the output scale.  Forward only; rel_packed required. */
int stb200_window_logits_forward_bf16(const stb200_index *ix, int h, int hdim, int L, const void *q_bf16, const void *k_bf16,
                                      const float *table_q, const float *table_k, float *logits, void *stream);
int stb200_window_aggregate_forward_bf16(const stb200_index *ix, int h, int hdim, int L, const float *attn, const void *v_bf16,
                                         const float *table_v, float *output, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* STB200_H */
