"""Extension-level mirror of the reference's `pointops2_cuda` pybind module.

Same function names, argument order and caller-allocates-outputs convention as
/root/reference/lib/pointops2/src/pointops_api.cpp:17-44 (signatures: SURVEY Appendix F), so code
written against `import pointops2_cuda as pointops_cuda` keeps working.  Each function forwards raw
device pointers to the C ABI of libstb200.so (include/stb200.h) on torch's CURRENT stream (the
reference always launched on the legacy default stream).  No CPU fallback: tensors must be CUDA.

Only the hot-path subset is provided: the 6 v2/v3 attention entry points, the 8 v1 ones,
attention_step2_*_v2 (bound by the reference but never called from its Python) and furthestsampling.
"""
from __future__ import annotations

from collections import OrderedDict

import os

import torch

from . import _cabi


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _f(t: torch.Tensor, name: str):
    if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
        raise TypeError(f"{name}: expected a contiguous CUDA float32 tensor, got {t.dtype} {t.device} contiguous={t.is_contiguous()}")
    return t.data_ptr()


def _i(t: torch.Tensor, name: str):
    if not (t.is_cuda and t.dtype == torch.int32 and t.is_contiguous()):
        raise TypeError(f"{name}: expected a contiguous CUDA int32 tensor, got {t.dtype} {t.device} contiguous={t.is_contiguous()}")
    return t.data_ptr()


# ------------------------------------------------------------------------------------------------
# transposed CSR (pairs grouped by key), built once per index set and cached
class TransposedCSR:
    __slots__ = ("t_offsets", "t_pair", "t_index0")

    def __init__(self, t_offsets, t_pair, t_index0):
        self.t_offsets, self.t_pair, self.t_index0 = t_offsets, t_pair, t_index0


def build_transposed_csr(index0_offsets: torch.Tensor, index1: torch.Tensor, n_keys: int | None = None) -> TransposedCSR:
    """Group the pair list by key.  n_keys = number of rows of k/v (defaults to the number of queries)."""
    N = index0_offsets.numel() - 1
    M = index1.numel()
    NK = N if n_keys is None else n_keys
    dev = index1.device
    t_offsets = torch.empty(NK + 1, dtype=torch.int32, device=dev)
    t_pair = torch.empty(M, dtype=torch.int32, device=dev)
    t_index0 = torch.empty(M, dtype=torch.int32, device=dev)
    if NK != N:
        # the C entry point sizes t_offsets by its first argument; expand_index0 walks the query offsets
        raise NotImplementedError("k/v with a different row count than q is not used by any model path")
    nbytes = _cabi.load().stb200_transpose_csr_workspace_bytes(N, M)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    _cabi.call("stb200_transpose_csr", N, M, _i(index0_offsets, "index0_offsets"), _i(index1, "index1"),
               t_offsets.data_ptr(), t_pair.data_ptr(), t_index0.data_ptr(), ws.data_ptr(), nbytes, _stream())
    return TransposedCSR(t_offsets, t_pair, t_index0)


_TCSR_CACHE: "OrderedDict[tuple, tuple]" = OrderedDict()
_TCSR_CACHE_SIZE = 4


def transposed_csr_for(index0_offsets: torch.Tensor, index1: torch.Tensor) -> TransposedCSR:
    """Cache keyed by storage identity + version; the cached entry keeps the index tensors alive, so a
    recycled device pointer can never alias a stale entry."""
    key = (index0_offsets.data_ptr(), index1.data_ptr(), index0_offsets.numel(), index1.numel(),
           index0_offsets._version, index1._version)
    hit = _TCSR_CACHE.get(key)
    if hit is not None:
        _TCSR_CACHE.move_to_end(key)
        return hit[0]
    t = build_transposed_csr(index0_offsets, index1)
    _TCSR_CACHE[key] = (t, index0_offsets, index1)
    while len(_TCSR_CACHE) > _TCSR_CACHE_SIZE:
        _TCSR_CACHE.popitem(last=False)
    return t


def clear_caches():
    _TCSR_CACHE.clear()


# ------------------------------------------------------------------------------------------------
# v2 / v3 (CSR) — the entry points the shipped models call
def attention_step1_forward_cuda_v2(N, M, h, C, n_max, q, k, index0_offsets, index1, attn):
    _cabi.call("stb200_attention_step1_forward_v2", N, M, h, C, int(n_max), _f(q, "q"), _f(k, "k"),
               _i(index0_offsets, "index0_offsets"), _i(index1, "index1"), _f(attn, "attn"), _stream())


def attention_step1_backward_cuda_v2(N, M, h, C, n_max, grad_out, index0_offsets, index1, q, k, grad_q, grad_k,
                                     tcsr: TransposedCSR | None = None):
    t = tcsr or transposed_csr_for(index0_offsets, index1)
    _cabi.call("stb200_attention_step1_backward_v2", N, M, h, C, int(n_max), _f(grad_out, "grad_out"),
               _i(index0_offsets, "index0_offsets"), _i(index1, "index1"), _f(q, "q"), _f(k, "k"),
               _f(grad_q, "grad_q"), _f(grad_k, "grad_k"), t.t_offsets.data_ptr(), t.t_pair.data_ptr(),
               t.t_index0.data_ptr(), _stream())


def dot_prod_with_idx_forward_cuda_v3(N, M, h, hdim, n_max, q, index_q_offsets, k, index_k, table_q, table_k,
                                      rel_idx, output):
    _cabi.call("stb200_dot_prod_with_idx_forward_v3", N, M, h, hdim, int(n_max), table_q.shape[0], _f(q, "q"),
               _i(index_q_offsets, "index_q_offsets"), _f(k, "k"), _i(index_k, "index_k"), _f(table_q, "table_q"),
               _f(table_k, "table_k"), _i(rel_idx, "rel_idx"), _f(output, "output"), _stream())


def dot_prod_with_idx_backward_cuda_v3(N, M, h, hdim, n_max, grad_out, q, index_q_offsets, k, index_k, table_q,
                                       table_k, rel_idx, grad_q, grad_k, grad_table_q, grad_table_k,
                                       tcsr: TransposedCSR | None = None):
    t = tcsr or transposed_csr_for(index_q_offsets, index_k)
    _cabi.call("stb200_dot_prod_with_idx_backward_v3", N, M, h, hdim, int(n_max), table_q.shape[0],
               _f(grad_out, "grad_out"), _f(q, "q"), _i(index_q_offsets, "index_q_offsets"), _f(k, "k"),
               _i(index_k, "index_k"), _f(table_q, "table_q"), _f(table_k, "table_k"), _i(rel_idx, "rel_idx"),
               _f(grad_q, "grad_q"), _f(grad_k, "grad_k"), _f(grad_table_q, "grad_table_q"),
               _f(grad_table_k, "grad_table_k"), t.t_offsets.data_ptr(), t.t_pair.data_ptr(),
               t.t_index0.data_ptr(), _stream())


def attention_step2_with_rel_pos_value_forward_cuda_v2(N, M, h, hdim, n_max, attn, v, index0_offsets, index1, table,
                                                       rel_idx, output):
    _cabi.call("stb200_attention_step2_with_rel_pos_value_forward_v2", N, M, h, hdim, int(n_max), table.shape[0],
               _f(attn, "attn"), _f(v, "v"), _i(index0_offsets, "index0_offsets"), _i(index1, "index1"),
               _f(table, "table"), _i(rel_idx, "rel_idx"), _f(output, "output"), _stream())


def attention_step2_with_rel_pos_value_backward_cuda_v2(N, M, h, hdim, n_max, grad_out, index0_offsets, index1, attn,
                                                        v, table, rel_idx, grad_attn, grad_v, grad_table,
                                                        tcsr: TransposedCSR | None = None):
    t = tcsr or transposed_csr_for(index0_offsets, index1)
    _cabi.call("stb200_attention_step2_with_rel_pos_value_backward_v2", N, M, h, hdim, int(n_max), table.shape[0],
               _f(grad_out, "grad_out"), _i(index0_offsets, "index0_offsets"), _i(index1, "index1"),
               _f(attn, "attn"), _f(v, "v"), _f(table, "table"), _i(rel_idx, "rel_idx"), _f(grad_attn, "grad_attn"),
               _f(grad_v, "grad_v"), _f(grad_table, "grad_table"), t.t_offsets.data_ptr(), t.t_pair.data_ptr(),
               t.t_index0.data_ptr(), _stream())


# segment softmax: new entry points (replace torch_scatter.scatter_softmax on the path)
def segment_softmax_forward_cuda(N, M, h, a, b, index0_offsets, p):
    _cabi.call("stb200_segment_softmax_forward", N, M, h, _f(a, "a"), None if b is None else _f(b, "b"),
               _i(index0_offsets, "index0_offsets"), _f(p, "p"), _stream())


def segment_softmax_backward_cuda(N, M, h, p, grad_p, index0_offsets, grad_s):
    _cabi.call("stb200_segment_softmax_backward", N, M, h, _f(p, "p"), _f(grad_p, "grad_p"),
               _i(index0_offsets, "index0_offsets"), _f(grad_s, "grad_s"), _stream())


# ------------------------------------------------------------------------------------------------
# v1 (explicit index0 / index1)
def attention_step1_forward_cuda(N, M, h, C, q, k, index0, index1, attn):
    _cabi.call("stb200_attention_step1_forward", N, M, h, C, _f(q, "q"), _f(k, "k"), _i(index0, "index0"),
               _i(index1, "index1"), _f(attn, "attn"), _stream())


def attention_step1_backward_cuda(N, M, h, C, grad_out, index0, index1, q, k, grad_q, grad_k):
    _cabi.call("stb200_attention_step1_backward", N, M, h, C, _f(grad_out, "grad_out"), _i(index0, "index0"),
               _i(index1, "index1"), _f(q, "q"), _f(k, "k"), _f(grad_q, "grad_q"), _f(grad_k, "grad_k"), _stream())


def attention_step2_forward_cuda(N, M, h, C, attn, v, index0, index1, output):
    _cabi.call("stb200_attention_step2_forward", N, M, h, C, _f(attn, "attn"), _f(v, "v"), _i(index0, "index0"),
               _i(index1, "index1"), _f(output, "output"), _stream())


def attention_step2_backward_cuda(N, M, h, C, grad_out, index0, index1, attn, v, grad_attn, grad_v):
    _cabi.call("stb200_attention_step2_backward", N, M, h, C, _f(grad_out, "grad_out"), _i(index0, "index0"),
               _i(index1, "index1"), _f(attn, "attn"), _f(v, "v"), _f(grad_attn, "grad_attn"), _f(grad_v, "grad_v"),
               _stream())


# the reference binds *_v2 duplicates of step2 that behave exactly like v1 (attention_cuda_kernel_v2.cu:148-195)
attention_step2_forward_cuda_v2 = attention_step2_forward_cuda
attention_step2_backward_cuda_v2 = attention_step2_backward_cuda


def dot_prod_with_idx_forward_cuda(N, M, h, hdim, q, index, table, rel_idx, output):
    _cabi.call("stb200_dot_prod_with_idx_forward", N, M, h, hdim, table.shape[0], _f(q, "q"), _i(index, "index"),
               _f(table, "table"), _i(rel_idx, "rel_idx"), _f(output, "output"), _stream())


def dot_prod_with_idx_backward_cuda(N, M, h, hdim, grad_out, q, index, table, rel_idx, grad_q, grad_table):
    _cabi.call("stb200_dot_prod_with_idx_backward", N, M, h, hdim, table.shape[0], _f(grad_out, "grad_out"),
               _f(q, "q"), _i(index, "index"), _f(table, "table"), _i(rel_idx, "rel_idx"), _f(grad_q, "grad_q"),
               _f(grad_table, "grad_table"), _stream())


def attention_step2_with_rel_pos_value_forward_cuda(N, M, h, hdim, attn, v, index0, index1, table, rel_idx, output):
    _cabi.call("stb200_attention_step2_with_rel_pos_value_forward", N, M, h, hdim, table.shape[0], _f(attn, "attn"),
               _f(v, "v"), _i(index0, "index0"), _i(index1, "index1"), _f(table, "table"), _i(rel_idx, "rel_idx"),
               _f(output, "output"), _stream())


def attention_step2_with_rel_pos_value_backward_cuda(N, M, h, hdim, grad_out, index0, index1, attn, v, table, rel_idx,
                                                     grad_attn, grad_v, grad_table):
    _cabi.call("stb200_attention_step2_with_rel_pos_value_backward", N, M, h, hdim, table.shape[0],
               _f(grad_out, "grad_out"), _i(index0, "index0"), _i(index1, "index1"), _f(attn, "attn"), _f(v, "v"),
               _f(table, "table"), _i(rel_idx, "rel_idx"), _f(grad_attn, "grad_attn"), _f(grad_v, "grad_v"),
               _f(grad_table, "grad_table"), _stream())


# dot_prod_with_idx v2 (pairs pre-sorted by merged rel idx, test-only in the reference): same math as v3 on a
# pair list given by (index_q, index_k); routed through the v1 single-table kernels.
def dot_prod_with_idx_forward_cuda_v2(N, M, h, hdim, n_max, T, q, index_q, k, index_k, table_q, table_k, rel_idx,
                                      rel_idx_offsets, sort_indices, output):
    tmp = torch.empty_like(output)
    dot_prod_with_idx_forward_cuda(N, M, h, hdim, q, index_q, table_q, rel_idx, output)
    dot_prod_with_idx_forward_cuda(N, M, h, hdim, k, index_k, table_k, rel_idx, tmp)
    output.add_(tmp)


def dot_prod_with_idx_backward_cuda_v2(N, M, h, hdim, n_max, T, grad_out, q, index_q, k, index_k, table_q, table_k,
                                       rel_idx, rel_idx_offsets, sort_indices, grad_q, grad_k, grad_table_q,
                                       grad_table_k):
    dot_prod_with_idx_backward_cuda(N, M, h, hdim, grad_out, q, index_q, table_q, rel_idx, grad_q, grad_table_q)
    dot_prod_with_idx_backward_cuda(N, M, h, hdim, grad_out, k, index_k, table_k, rel_idx, grad_k, grad_table_k)


# ------------------------------------------------------------------------------------------------
def furthestsampling_cuda(b, n, xyz, offset, new_offset, tmp, idx):
    """same positional arguments as the reference pybind function (sampling/sampling_cuda_kernel.h:7).  `tmp` (the
    reference's 1e10-filled scratch) is only read by the streaming fallback for scenes that do not fit the
    register-resident cluster kernel; None allocates it here."""
    if tmp is None:
        tmp = torch.full((xyz.shape[0],), 1e10, dtype=torch.float32, device=xyz.device)
    # exact bounding-box pruning (include/stb200.h: stb200_furthestsampling_ws); the library itself falls back to the plain
    # kernel for small scenes or STB200_FPS_PRUNE=0
    N = int(xyz.shape[0])
    nbytes = int(_cabi.load().stb200_fps_workspace_bytes(N, int(b)))
    ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=xyz.device)
    _cabi.call("stb200_furthestsampling_ws", int(b), int(n), N, _f(xyz, "xyz"), _i(offset, "offset"), _i(new_offset, "new_offset"),
               None if tmp is None else _f(tmp, "tmp"), _i(idx, "idx"), ws.data_ptr(), nbytes, _stream())


def knnquery_cuda(m, nsample, xyz, new_xyz, offset, new_offset, idx, dist2):
    """same positional arguments as the reference pybind function (knnquery/knnquery_cuda_kernel.h:7); grid-pruned search with
    the exact heap scan as its completion (include/stb200.h); STB200_KNN_BRUTE=1: heap scan only"""
    n, b = int(xyz.shape[0]), int(offset.numel())
    if os.environ.get("STB200_KNN_BRUTE"):
        _cabi.call("stb200_knnquery", int(m), b, int(nsample), _f(xyz, "xyz"), _f(new_xyz, "new_xyz"),
                   _i(offset, "offset"), _i(new_offset, "new_offset"), _i(idx, "idx"), _f(dist2, "dist2"), _stream())
        return
    nbytes = int(_cabi.load().stb200_knnquery_workspace_bytes(n, int(m), b))
    ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=xyz.device)
    _cabi.call("stb200_knnquery_ws", n, int(m), b, int(nsample), _f(xyz, "xyz"), _f(new_xyz, "new_xyz"), _i(offset, "offset"),
               _i(new_offset, "new_offset"), _i(idx, "idx"), _f(dist2, "dist2"), ws.data_ptr(), nbytes, _stream())
