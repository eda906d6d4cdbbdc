"""Oracle (attention math) self-pinning on CPU: explicit backward == autograd of the authors'
pure-torch formulation (lib/pointops2/functions/test_relative_pos_encoding_op_step2.py:32-45),
vectorised form == literal kernel loops, v1 identities the reference's test scripts assert."""
import numpy as np
import pytest
import torch

from oracle import attention_oracle as ao


def make_case(N=60, M=900, h=2, d=16, L=9, seed=0, dtype=torch.float64, empty=True):
    g = torch.Generator().manual_seed(seed)
    q, k, v = (torch.rand(N, h, d, generator=g, dtype=dtype) for _ in range(3))
    tq, tk, tv = (torch.rand(L, h, d, 3, generator=g, dtype=dtype) for _ in range(3))
    i0 = torch.sort((torch.rand(M, generator=g) * N).long()).values
    if empty:  # make a few queries empty, including the last one
        i0 = i0[(i0 != 3) & (i0 != N - 1)]
    M = i0.numel()
    i1 = (torch.rand(M, generator=g) * N).long()
    rel = (torch.rand(M, 3, generator=g) * L).long()
    offsets = torch.cat([torch.zeros(1, dtype=torch.long), torch.bincount(i0, minlength=N).cumsum(0)])
    g_out = torch.rand(N, h, d, generator=g, dtype=dtype)
    return q, k, v, offsets, i1, tq, tk, tv, rel, g_out


def test_explicit_backward_equals_autograd():
    args = make_case()
    e = ao.layer_fwd_bwd(*args)
    a = ao.layer_autograd(*args)
    for key in ("out", "p", "gq", "gk", "gv", "gtq", "gtk", "gtv"):
        assert torch.allclose(e[key], a[key], rtol=1e-10, atol=1e-12), key


def test_vectorised_equals_kernel_loops_fp32():
    args = make_case(N=25, M=300, dtype=torch.float32, seed=3)
    f = ao.layer_fwd(*args[:9])
    l = ao.layer_loops(*args[:9])
    for key in ("a", "b", "p", "out"):
        assert torch.allclose(f[key], l[key], rtol=2e-5, atol=2e-6), key


def test_reference_script_identities():
    """v2 == v1(q) + v1(k)  (test_relative_pos_encoding_op_step1_v2.py:63, ..._v3.py:89) and
    step2 with value table == v1 form attn*(v/3 + T) summed over axes (rpe kernel :87)."""
    q, k, v, offsets, i1, tq, tk, tv, rel, g_out = make_case(seed=5)
    i0 = ao.index0_from_offsets(offsets)
    both = ao.rpe_fwd(q, k, i0, i1, tq, tk, rel)
    split = ao.rpe_single_fwd(q, i0, tq, rel) + ao.rpe_single_fwd(k, i1, tk, rel)
    assert torch.allclose(both, split, rtol=1e-12)
    p = torch.rand(i1.numel(), q.shape[1], dtype=q.dtype)
    out = ao.step2_rpv_fwd(p, v, i0, i1, tv, rel, q.shape[0])
    r = rel.long()
    v1 = torch.zeros_like(out)
    for a in range(3):
        v1.index_add_(0, i0, p.unsqueeze(-1) * (v[i1] / 3.0 + tv[r[:, a], :, :, a]))
    assert torch.allclose(out, v1, rtol=1e-12)
    gq, gk, gtq, gtk = ao.rpe_bwd(p, q, k, i0, i1, tq, tk, rel)
    gq1, gtq1 = ao.rpe_single_bwd(p, q, i0, tq, rel)
    gk1, gtk1 = ao.rpe_single_bwd(p, k, i1, tk, rel)
    for x, y in ((gq, gq1), (gk, gk1), (gtq, gtq1), (gtk, gtk1)):
        assert torch.allclose(x, y, rtol=1e-12)


def test_softmax_properties():
    q, k, v, offsets, i1, tq, tk, tv, rel, g_out = make_case(seed=7)
    N = q.shape[0]
    i0 = ao.index0_from_offsets(offsets)
    s = torch.randn(i1.numel(), 2, dtype=torch.float64) * 5
    p = ao.softmax_fwd(s, i0, N)
    sums = torch.zeros(N, 2, dtype=torch.float64).index_add_(0, i0, p)
    nonempty = (offsets[1:] - offsets[:-1]) > 0
    assert torch.allclose(sums[nonempty], torch.ones_like(sums[nonempty]))
    assert torch.allclose(p, ao.softmax_fwd(s + 100.0, i0, N))  # shift invariance
    s.requires_grad_(True)
    gp = torch.rand_like(p)
    # autograd of a dense per-segment softmax
    ref = torch.cat([torch.softmax(s[offsets[n]:offsets[n + 1]], 0) for n in range(N)])
    ref.backward(gp)
    assert torch.allclose(ao.softmax_bwd(p, gp, i0, N), s.grad, rtol=1e-9, atol=1e-12)


def test_plain_step2_matches_scatter_form():
    q, k, v, offsets, i1, *_ = make_case(seed=9)
    i0 = ao.index0_from_offsets(offsets)
    p = torch.rand(i1.numel(), 2, dtype=torch.float64, requires_grad=True)
    v = v.clone().requires_grad_(True)
    out = torch.zeros_like(q).index_add(0, i0, p.unsqueeze(-1) * v[i1])   # test_attention_op_step2.py:25-27
    assert torch.allclose(out, ao.step2_fwd(p.detach(), v.detach(), i0, i1, q.shape[0]))
    g = torch.rand_like(out)
    out.backward(g)
    gp, gv = ao.step2_bwd(g, p.detach(), v.detach(), i0, i1)
    assert torch.allclose(gp, p.grad) and torch.allclose(gv, v.grad)
