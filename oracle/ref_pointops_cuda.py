"""TEST / BASELINE INFRASTRUCTURE: a stand-in for the reference's pybind module `pointops2_cuda`
(/root/reference/lib/pointops2/src/pointops_api.cpp:17-44) on top of oracle/_ref/libpointops2_ref.so, i.e. the reference's own
kernels compiled where they lie (oracle/Makefile).  Same function names and argument orders as the pybind functions; each
one passes the tensors' device pointers to the reference's `extern "C"` launcher exactly like the reference's host shims
(src/*/*_cuda*.cpp) do.  With it the reference's `functions/pointops.py` runs unmodified (oracle/_ref/ref_pointops.py, generated),
which is what the layer-level baseline of tests/test_gpu_layers.py and bench.py's `ref_layer_baseline` leg time.
The launchers use the legacy default stream, which is torch's default stream: no extra synchronisation is added.
Only the entry points the model's layers reach are bound."""
import ctypes

from . import ref_cuda


def _p(t):
    assert t.is_cuda and t.is_contiguous(), "reference launchers take contiguous CUDA tensors"
    return ctypes.c_void_p(t.data_ptr())


def _launch(name, *args):
    fn = getattr(ref_cuda.lib(), name)
    fn.restype = None
    fn(*[a if isinstance(a, (ctypes.c_void_p, ctypes.c_uint)) else ctypes.c_int(int(a)) for a in args])


def _u(v):
    return ctypes.c_uint(int(v))


def furthestsampling_cuda(b, n, xyz, offset, new_offset, tmp, idx):
    _launch("furthestsampling_cuda_launcher", b, n, _p(xyz), _p(offset), _p(new_offset), _p(tmp), _p(idx))


def knnquery_cuda(m, nsample, xyz, new_xyz, offset, new_offset, idx, dist2):
    _launch("knnquery_cuda_launcher", m, nsample, _p(xyz), _p(new_xyz), _p(offset), _p(new_offset), _p(idx), _p(dist2))


def attention_step1_forward_cuda_v2(N, M, h, C, n_max, q, k, index0_offsets, index1, attn):
    _launch("attention_step1_forward_cuda_launcher_v2", N, M, h, C, _u(n_max), _p(q), _p(k), _p(index0_offsets), _p(index1), _p(attn))


def attention_step1_backward_cuda_v2(N, M, h, C, n_max, grad_out, index0_offsets, index1, q, k, grad_q, grad_k):
    _launch("attention_step1_backward_cuda_launcher_v2", N, M, h, C, _u(n_max), _p(grad_out), _p(index0_offsets), _p(index1), _p(q),
            _p(k), _p(grad_q), _p(grad_k))


def dot_prod_with_idx_forward_cuda_v3(N, M, h, hdim, n_max, q, index_q_offsets, k, index_k, table_q, table_k, rel_idx, output):
    _launch("dot_prod_with_idx_forward_cuda_launcher_v3", N, M, h, hdim, n_max, _p(q), _p(index_q_offsets), _p(k), _p(index_k),
            _p(table_q), _p(table_k), _p(rel_idx), _p(output))


def dot_prod_with_idx_backward_cuda_v3(N, M, h, hdim, n_max, grad_out, q, index_q_offsets, k, index_k, table_q, table_k, rel_idx,
                                       grad_q, grad_k, grad_table_q, grad_table_k):
    _launch("dot_prod_with_idx_backward_cuda_launcher_v3", N, M, h, hdim, n_max, _p(grad_out), _p(q), _p(index_q_offsets), _p(k),
            _p(index_k), _p(table_q), _p(table_k), _p(rel_idx), _p(grad_q), _p(grad_k), _p(grad_table_q), _p(grad_table_k))


def attention_step2_with_rel_pos_value_forward_cuda_v2(N, M, h, hdim, n_max, attn, v, index0_offsets, index1, table, rel_idx, output):
    _launch("attention_step2_with_rel_pos_value_forward_cuda_launcher_v2", N, M, h, hdim, n_max, _p(attn), _p(v), _p(index0_offsets),
            _p(index1), _p(table), _p(rel_idx), _p(output))


def attention_step2_with_rel_pos_value_backward_cuda_v2(N, M, h, hdim, n_max, grad_out, index0_offsets, index1, attn, v, table,
                                                        rel_idx, grad_attn, grad_v, grad_table):
    _launch("attention_step2_with_rel_pos_value_backward_cuda_launcher_v2", N, M, h, hdim, n_max, _p(grad_out), _p(index0_offsets),
            _p(index1), _p(attn), _p(v), _p(table), _p(rel_idx), _p(grad_attn), _p(grad_v), _p(grad_table))


def __getattr__(name):
    raise AttributeError(f"oracle.ref_pointops_cuda: the reference entry point {name!r} is not bound (only what the model's layers call is)")
