"""ctypes binding of oracle/_ref/libpointops2_ref.so (TEST INFRASTRUCTURE, GPU only).

That library is the REFERENCE's own kernels: oracle/Makefile compiles
/root/reference/lib/pointops2/src/{attention,attention_v2,rpe,rpe_v2,sampling}/*_kernel*.cu where they
lie (no source is copied) and exports their `extern "C"` launchers
(e.g. attention_v2/attention_cuda_kernel_v2.h:18-19, rpe_v2/relative_pos_encoding_cuda_kernel_v2.h:16-30,
sampling/sampling_cuda_kernel.h:10).  The launchers use the legacy default stream and return void, so every
wrapper here synchronises around the call.  Used by tests/test_gpu_parity.py to pin both the CPU oracle and
the new kernels against the reference itself, and by tools/ref_gpu_compare.py as the "reference kernels on
B200" timing baseline.
"""
from __future__ import annotations

import ctypes
import os

import torch

_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libpointops2_ref.so")
_lib = None


def available() -> bool:
    return os.path.exists(_SO)


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_SO)
    return _lib


def _p(t):
    assert t.is_cuda and t.is_contiguous()
    return ctypes.c_void_p(t.data_ptr())


def _call(name, *args):
    torch.cuda.synchronize()
    fn = getattr(lib(), name)
    fn.restype = None
    conv = []
    for a in args:
        if isinstance(a, ctypes.c_void_p):
            conv.append(a)
        elif isinstance(a, _U):
            conv.append(ctypes.c_uint(int(a)))
        else:
            conv.append(ctypes.c_int(int(a)))
    fn(*conv)
    torch.cuda.synchronize()


class _U(int):
    """marks an `unsigned int` argument"""


def n_max_of(offsets):
    return int((offsets[1:] - offsets[:-1]).max().item())


def step1_fwd(q, k, offsets, index1):
    N, h, d = q.shape
    M = index1.numel()
    out = torch.zeros(M, h, device=q.device)
    _call("attention_step1_forward_cuda_launcher_v2", N, M, h, h * d, _U(n_max_of(offsets)), _p(q), _p(k), _p(offsets), _p(index1), _p(out))
    return out


def step1_bwd(g, q, k, offsets, index1):
    N, h, d = q.shape
    M = index1.numel()
    gq, gk = torch.zeros_like(q), torch.zeros_like(k)
    _call("attention_step1_backward_cuda_launcher_v2", N, M, h, h * d, _U(n_max_of(offsets)), _p(g), _p(offsets), _p(index1), _p(q), _p(k), _p(gq), _p(gk))
    return gq, gk


def rpe_fwd(q, k, offsets, index1, tq, tk, rel):
    N, h, d = q.shape
    M = index1.numel()
    out = torch.zeros(M, h, device=q.device)
    _call("dot_prod_with_idx_forward_cuda_launcher_v3", N, M, h, d, n_max_of(offsets), _p(q), _p(offsets), _p(k), _p(index1), _p(tq), _p(tk), _p(rel), _p(out))
    return out


def rpe_bwd(g, q, k, offsets, index1, tq, tk, rel):
    N, h, d = q.shape
    M = index1.numel()
    gq, gk, gtq, gtk = torch.zeros_like(q), torch.zeros_like(k), torch.zeros_like(tq), torch.zeros_like(tk)
    _call("dot_prod_with_idx_backward_cuda_launcher_v3", N, M, h, d, n_max_of(offsets), _p(g), _p(q), _p(offsets), _p(k), _p(index1), _p(tq), _p(tk), _p(rel), _p(gq), _p(gk), _p(gtq), _p(gtk))
    return gq, gk, gtq, gtk


def step2_rpv_fwd(p, v, offsets, index1, tv, rel):
    N, h, d = v.shape
    M = index1.numel()
    out = torch.zeros(N, h, d, device=v.device)
    _call("attention_step2_with_rel_pos_value_forward_cuda_launcher_v2", N, M, h, d, n_max_of(offsets), _p(p), _p(v), _p(offsets), _p(index1), _p(tv), _p(rel), _p(out))
    return out


def step2_rpv_bwd(g, p, v, offsets, index1, tv, rel):
    N, h, d = v.shape
    M = index1.numel()
    gp, gv, gt = torch.zeros_like(p), torch.zeros_like(v), torch.zeros_like(tv)
    _call("attention_step2_with_rel_pos_value_backward_cuda_launcher_v2", N, M, h, d, n_max_of(offsets), _p(g), _p(offsets), _p(index1), _p(p), _p(v), _p(tv), _p(rel), _p(gp), _p(gv), _p(gt))
    return gp, gv, gt


# v1 (explicit index0/index1)
def v1_step1_fwd(q, k, i0, i1):
    N, h, d = q.shape
    out = torch.zeros(i0.numel(), h, device=q.device)
    _call("attention_step1_forward_cuda_launcher", k.shape[0], i0.numel(), h, h * d, _p(q), _p(k), _p(i0), _p(i1), _p(out))
    return out


def v1_step2_fwd(p, v, i0, i1, N):
    _, h, d = v.shape
    out = torch.zeros(N, h, d, device=v.device)
    _call("attention_step2_forward_cuda_launcher", N, i0.numel(), h, h * d, _p(p), _p(v), _p(i0), _p(i1), _p(out))
    return out


def v1_rpe_fwd(x, index, table, rel):
    N, h, d = x.shape
    out = torch.zeros(index.numel(), h, device=x.device)
    _call("dot_prod_with_idx_forward_cuda_launcher", N, index.numel(), h, d, _p(x), _p(index), _p(table), _p(rel), _p(out))
    return out


def v1_step2_rpv_fwd(p, v, i0, i1, table, rel, N):
    _, h, d = v.shape
    out = torch.zeros(N, h, d, device=v.device)
    _call("attention_step2_with_rel_pos_value_forward_cuda_launcher", N, i0.numel(), h, d, _p(p), _p(v), _p(i0), _p(i1), _p(table), _p(rel), _p(out))
    return out


def furthestsampling(xyz, offset, new_offset):
    b = offset.numel()
    sizes = torch.diff(offset, prepend=offset.new_zeros(1))
    n_max = int(sizes.max().item())
    idx = torch.zeros(int(new_offset[-1].item()), dtype=torch.int32, device=xyz.device)
    tmp = torch.full((xyz.shape[0],), 1e10, device=xyz.device)
    _call("furthestsampling_cuda_launcher", b, n_max, _p(xyz), _p(offset), _p(new_offset), _p(tmp), _p(idx))
    return idx


def knnquery(nsample, xyz, new_xyz, offset, new_offset):
    m = new_xyz.shape[0]
    idx = torch.zeros(m, nsample, dtype=torch.int32, device=xyz.device)
    d2 = torch.zeros(m, nsample, device=xyz.device)
    _call("knnquery_cuda_launcher", m, nsample, _p(xyz), _p(new_xyz), _p(offset), _p(new_offset), _p(idx), _p(d2))
    return idx, d2


def timed(name, args, reps=10, warmup=2):
    """Average device time in ms of the reference launcher `name` (same argument convention as _call), measured with
    CUDA events on the legacy default stream the launchers use (torch's default stream).  bench.py's ref_cuda_baseline leg."""
    fn = getattr(lib(), name)
    fn.restype = None
    conv = []
    for a in args:
        if isinstance(a, ctypes.c_void_p):
            conv.append(a)
        elif isinstance(a, _U):
            conv.append(ctypes.c_uint(int(a)))
        else:
            conv.append(ctypes.c_int(int(a)))
    torch.cuda.synchronize()
    for _ in range(warmup):
        fn(*conv)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(torch.cuda.default_stream())
    for _ in range(reps):
        fn(*conv)
    e1.record(torch.cuda.default_stream())
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
