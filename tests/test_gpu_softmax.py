"""Segment softmax (a13: bias add + torch_scatter.scatter_softmax, /root/reference/model/stratified_transformer.py:203-205) forward
and backward through the C ABI against a dense fp64 softmax per segment: ragged and empty segments, every head-count class,
chunks that exceed the staging buffer of the span kernels (per-row path), unaligned views."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ragged(N, max_len, seed, empty_every=7, long_rows=()):
    rng = np.random.default_rng(seed)
    lens = rng.integers(1, max_len + 1, N)
    lens[::empty_every] = 0
    for r, l in long_rows:
        lens[r] = l
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    return lens, off


def _dense_reference(s, g, off):
    """fp64: p = softmax per segment, gs = p * (g - <p, g>)"""
    p = torch.empty_like(s)
    gs = torch.empty_like(s)
    for n in range(len(off) - 1):
        a, b = int(off[n]), int(off[n + 1])
        if b > a:
            p[a:b] = torch.softmax(s[a:b], 0)
            gs[a:b] = p[a:b] * (g[a:b] - (p[a:b] * g[a:b]).sum(0, keepdim=True))
    return p, gs


@pytest.mark.parametrize("h", [1, 2, 3, 4, 6, 8, 12, 16, 24, 32])
@pytest.mark.parametrize("with_bias", [False, True])
def test_segment_softmax_fwd_bwd_vs_dense(h, with_bias):
    from stratified_transformer_b200 import pointops
    N = 700
    # rows 100 and 400 alone exceed the span buffer for h >= 3 (3000 * h floats): per-row path inside the span kernel
    lens, off = _ragged(N, 70, seed=h, long_rows=((100, 3000), (400, 2900)))
    M = int(off[-1])
    g = torch.Generator().manual_seed(h)
    a = (torch.randn(M, h, generator=g) * 3).double()
    b = (torch.randn(M, h, generator=g)).double() if with_bias else None
    gp = torch.randn(M, h, generator=g).double()
    want_p, want_gs = _dense_reference(a if b is None else a + b, gp, off)
    a_d = a.float().cuda().requires_grad_(True)
    b_d = None if b is None else b.float().cuda().requires_grad_(True)
    p = pointops.segment_softmax(a_d, torch.from_numpy(off).cuda(), b_d)
    assert (p.double().cpu() - want_p).abs().max() < 1e-5
    p.backward(gp.float().cuda())
    assert (a_d.grad.double().cpu() - want_gs).abs().max() < 3e-5
    if b_d is not None:
        assert torch.equal(b_d.grad, a_d.grad)
    # every non-empty segment sums to one per head
    sums = torch.zeros(N, h, dtype=torch.float64).index_add(0, torch.repeat_interleave(torch.arange(N), torch.from_numpy(lens)), p.double().cpu())
    assert (sums[torch.from_numpy(lens) > 0] - 1).abs().max() < 1e-5


def test_segment_softmax_unaligned_views_and_variants_agree():
    """A view that starts at an odd float offset takes the per-row kernels; both give the same numbers to 1 ulp."""
    from stratified_transformer_b200 import pointops2_cuda as ext
    h, N = 3, 5000
    lens, off = _ragged(N, 64, seed=11)
    M = int(off[-1])
    torch.manual_seed(3)
    a0 = torch.randn(M, h, device="cuda") * 2
    off_d = torch.from_numpy(off).cuda()
    outs = []
    for shift in (0, 1, 4):
        a = torch.empty(M * h + 8, device="cuda")[shift:shift + M * h].view(M, h)
        a.copy_(a0)
        p = torch.empty(M * h + 8, device="cuda")[shift:shift + M * h].view(M, h)
        ext.segment_softmax_forward_cuda(N, M, h, a, None, off_d, p)
        outs.append(p.clone())
    # aligned views take the span kernel (one reciprocal per row), the odd one the per-row kernel (a division per element)
    assert torch.equal(outs[0], outs[2]) and torch.allclose(outs[0], outs[1], rtol=3e-7, atol=0)
    want, _ = _dense_reference(a0.double().cpu(), a0.double().cpu(), off)
    assert (outs[0].double().cpu() - want).abs().max() < 1e-5
