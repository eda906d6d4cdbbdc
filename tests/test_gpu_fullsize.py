"""Full-size checks (-m gpu) at BASELINE shapes (one 80k-point S3DIS-shape scene, layer-0 geometry) through
size-independent properties: the oracle is too slow at this size, invariants are not."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def scene():
    from stratified_transformer_b200 import index
    from stratified_transformer_b200.synthetic import make_batch
    xyz, _, offset = make_batch(1, 80000, seed0=5)
    xd, od = torch.from_numpy(xyz).cuda(), torch.from_numpy(offset).cuda()
    li = index.build_layer_index(xd, od, 0.16, 0.01, 8, want_index_0=True)
    return xd, od, li


def test_index_invariants_full_size(scene):
    xd, od, li = scene
    N = xd.shape[0]
    assert li.downsample_idx.numel() == N // 8 + 1 and li.downsample_idx.unique().numel() == li.downsample_idx.numel()
    for parity in (0, 1):
        pi = li.for_block(parity)
        off = pi.index_0_offsets.long()
        cnt = off[1:] - off[:-1]
        assert int(off[0]) == 0 and int(off[-1]) == pi.M and int(cnt.max()) == pi.n_max and int(cnt.min()) >= 1
        i0, i1 = pi.index_0.long(), pi.index_1.long()
        assert torch.equal(i0, torch.repeat_interleave(torch.arange(N, device="cuda"), cnt))
        assert bool(((i1 >= 0) & (i1 < N)).all())
        # every point attends to itself, and the rel-pos index of that pair is the centre bin in all three axes
        self_pair = i0 == i1
        assert int(torch.zeros(N, dtype=torch.long, device="cuda").index_add_(0, i0[self_pair], torch.ones_like(i0[self_pair])).min()) >= 1
        assert bool((pi.rel_idx[self_pair] == 31).all())
        assert int(pi.rel_idx.min()) >= 0 and int(pi.rel_idx.max()) <= 63
        # window bookkeeping: windows partition the points; the dense block of a query is its whole window
        wo = pi.win_offsets.long()
        assert int(wo[0]) == 0 and int(wo[-1]) == N and bool((wo[1:] > wo[:-1]).all())
        assert torch.equal(pi.row_order.long().sort().values, torch.arange(N, device="cuda"))
        win_size = torch.repeat_interleave(wo[1:] - wo[:-1], wo[1:] - wo[:-1])
        assert bool((cnt[pi.row_order.long()] >= win_size).all())
        # transposed CSR is a permutation of the pairs, grouped by key
        t = pi.tcsr
        assert torch.equal(t.t_pair.long().sort().values, torch.arange(pi.M, device="cuda"))
        assert torch.equal(i1[t.t_pair.long()], torch.repeat_interleave(torch.arange(N, device="cuda"), (t.t_offsets[1:] - t.t_offsets[:-1]).long()))
        assert torch.equal(i0[t.t_pair.long()], t.t_index0.long())


def test_attention_properties_full_size(scene):
    from stratified_transformer_b200 import pointops
    xd, od, li = scene
    N, h, d, L = xd.shape[0], 3, 16, 64
    pi = li.for_block(1)
    g = torch.Generator(device="cuda").manual_seed(0)
    q, k, v, v2 = (torch.randn(N, h, d, device="cuda", generator=g) for _ in range(4))
    tq, tk, tv = (torch.randn(L, h, d, 3, device="cuda", generator=g) * 0.1 for _ in range(3))
    s = pointops.window_logits(q, k, tq, tk, pi)
    p = pointops.segment_softmax(s, pi.index_0_offsets)
    # probabilities: non-negative, each query's segment sums to one per head
    sums = torch.zeros(N, h, device="cuda").index_add_(0, pi.index_0.long(), p)
    assert float(p.min()) >= 0.0 and float((sums - 1).abs().max()) < 1e-4
    # shift invariance of the softmax, linearity of the aggregation in v, and the constant-v identity
    p2 = pointops.segment_softmax(s + 3.0, pi.index_0_offsets)
    assert float((p - p2).abs().max()) < 1e-5
    o1, o2 = pointops.window_aggregate(p, v, tv, pi), pointops.window_aggregate(p, v2, tv, pi)
    o12, o0 = pointops.window_aggregate(p, v + v2, tv, pi), pointops.window_aggregate(p, torch.zeros_like(v), tv, pi)
    assert float((o12 - (o1 + o2 - o0)).abs().max()) < 2e-4
    ones = pointops.window_aggregate(p, torch.ones_like(v), torch.zeros_like(tv), pi)
    assert float((ones - 1).abs().max()) < 1e-4
    # per-op API == fused entry points at full size
    a = pointops.attention_step1_v2(q, k, pi.index_1, pi.index_0_offsets, pi.n_max)
    b = pointops.dot_prod_with_idx_v3(q, pi.index_0_offsets, pi.n_max, k, pi.index_1, tq, tk, pi.rel_idx)
    assert float((a + b - s).abs().max()) < 1e-4 * max(1.0, float(s.abs().max()))
    # gradient identity: d/dt sum(out * G) along a random direction of v equals <grad_v, dv>
    vv = v.clone().requires_grad_(True)
    G = torch.randn(N, h, d, device="cuda", generator=g)
    (pointops.window_aggregate(p, vv, tv, pi) * G).sum().backward()
    dv = torch.randn_like(v)
    lhs = float((vv.grad * dv).sum())
    rhs = float(((pointops.window_aggregate(p, v + dv, tv, pi) - o1) * G).sum())   # exact: the op is linear in v
    assert abs(lhs - rhs) <= 2e-3 * max(1.0, abs(rhs))
