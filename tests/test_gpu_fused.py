"""GPU parity of the window-centric fused path (-m gpu), all through the C ABI.

* the device plan builder against its numpy oracle (oracle/fused_plan_oracle.py, itself pinned to the reference's pair
  construction in tests/test_fused_emu.py): tiles, orders and window tables bit-exact, items as sets;
* stb200_fused_attention_forward/backward on builder-produced plans against
    - the fp64 CPU oracle (small scenes, incl. the lattice scene with duplicate keys and scenes whose windows need chunking),
    - the REFERENCE's own CUDA kernels (oracle/_ref) on one 80k-point scene at the four layer shapes of BASELINE cfg2
      (C/h 48/3, 96/6, 192/12, 384/24, L=64) — every output and every gradient, on the index the bench times;
* the WindowAttention module with a plan against the same module on the per-op path.
fp32 tolerances: 1e-4 * max(1, |ref|) (2e-4 for table gradients, sums of > 1e5 terms), as in tests/test_gpu_parity.py.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _check(got, want, tol, name, scale_floor=0.0):
    """'' when |got - want| <= tol * max(1, |want|) (+ scale_floor * max|want|) everywhere, else a description"""
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    if not torch.isfinite(got).all():
        return f"{name}: non-finite values"
    err = (got - want).abs()
    bound = tol * torch.clamp(want.abs(), min=1.0) + scale_floor * want.abs().max()
    bad = err > bound
    if bool(bad.any()):
        return f"{name}: max err {float(err.max()):.3e} at |ref| max {float(want.abs().max()):.2f}, {int(bad.sum())} of {bad.numel()} bad"
    return ""


def _close(got, want, tol, name):
    msg = _check(got, want, tol, name)
    assert not msg, msg


def _small_scene(n_pts, seed, lattice=False, scenes=2):
    from oracle import fps_oracle, index_oracle as io
    from stratified_transformer_b200.synthetic import make_scene
    xs = [make_scene(seed + s, n_pts, n_raw=60000, lattice=lattice)[0] for s in range(scenes)]
    xyz = np.concatenate(xs)
    offset = np.cumsum([x.shape[0] for x in xs]).astype(np.int32)
    ds = fps_oracle.furthestsampling(xyz, offset, io.fps_new_offset(offset, 8))
    return xyz, offset, ds


def _items_set(items, counts):
    out, off = [], 0
    for c in counts:
        out.append(sorted(map(tuple, items[off:off + c, :7].tolist())))
        off += c
    return out


@pytest.mark.parametrize("parity", [0, 1])
@pytest.mark.parametrize("lattice", [False, True])
def test_plan_matches_oracle(parity, lattice):
    from oracle import fused_plan_oracle as fpo
    from stratified_transformer_b200 import index
    xyz, offset, ds = _small_scene(2500, 3, lattice)
    window, quant = 0.32, 0.02
    b = index.FUSED_BLOCKS
    want = fpo.build(xyz, offset, window, quant, parity, ds, BQ=b["BQ"], BK=b["BK"], BQS=b["BQS"], BKS=b["BKS"])
    pi = index.build_stratified_index(torch.from_numpy(xyz).cuda(), torch.from_numpy(offset).cuda(), window, quant,
                                      torch.from_numpy(ds).cuda(), parity, fused=True)
    plan = pi.plan
    d, s = want["dense"], want["sparse"]
    assert plan.totals[0] == d["rel"].shape[0] and plan.totals[1] == s["rel"].shape[0]
    assert plan.totals[2] == s["n_win"] and plan.totals[3] == d["max_win"] and plan.totals[4] == s["max_ns"]
    assert plan.totals[8:16] == d["counts"].tolist() and plan.totals[16:24] == s["counts"].tolist()
    assert np.array_equal(plan.order_s.cpu().numpy(), d["q_order"])
    assert np.array_equal(plan.wstart_s.cpu().numpy(), d["wstart"])
    assert np.array_equal(plan.pos_win.cpu().numpy(), d["pos_win"])
    assert np.array_equal(plan.tile_base.cpu().numpy(), d["tile_base"])
    assert np.array_equal(plan.dense_rel.cpu().numpy().view(np.uint32), d["rel"])
    assert np.array_equal(plan.order_l.cpu().numpy(), s["q_order"])
    assert np.array_equal(plan.samp.cpu().numpy(), s["k_order"])
    assert np.array_equal(plan.sparse_rel.cpu().numpy().view(np.uint32), s["rel"])
    assert _items_set(plan.dense_items.cpu().numpy(), d["counts"]) == _items_set(d["items"], d["counts"])
    assert _items_set(plan.sparse_items.cpu().numpy(), s["counts"]) == _items_set(s["items"], s["counts"])
    # the dense bins stay inside the range the dense pass stages
    tot = plan._totals_dev.tolist()
    lo, RB = plan.bin_range(2 * int((2 * window + 1e-4) // quant), True)
    assert lo <= tot[6] and tot[7] < lo + RB, (tot[6], tot[7], lo, RB)


def _fused_vs(want_fn, xyz, offset, ds, window, quant, parity, h, seed, table_scale=0.5, table_grad_floor=0.0):
    from stratified_transformer_b200 import index, pointops
    xd, od = torch.from_numpy(xyz).cuda(), torch.from_numpy(offset).cuda()
    pi = index.build_stratified_index(xd, od, window, quant, None if ds is None else torch.from_numpy(ds).cuda(), parity, fused=True)
    N = xyz.shape[0]
    L = 2 * int((2 * window + 1e-4) // quant)
    g = torch.Generator().manual_seed(seed)
    q, k, v, go = (torch.randn(N, h, 16, generator=g) for _ in range(4))
    q = q * 0.5
    tq, tk, tv = ((torch.rand(L, h, 16, 3, generator=g) - 0.5) * 2 * table_scale for _ in range(3))
    leaves = [t.cuda().requires_grad_(True) for t in (q, k, v, tq, tk, tv)]
    out = pointops.window_attention_plan(*leaves, pi.plan)
    out.backward(go.cuda())
    got = dict(out=out, gq=leaves[0].grad, gk=leaves[1].grad, gv=leaves[2].grad, gtq=leaves[3].grad, gtk=leaves[4].grad,
               gtv=leaves[5].grad)
    want = want_fn(pi, q, k, v, tq, tk, tv, go)
    msgs = [_check(val, want[name], 2e-4 if name.startswith("gt") else 1e-4, f"{name} (parity {parity}, h {h})",
                   table_grad_floor if name.startswith("gt") else 0.0) for name, val in got.items()]
    assert not any(msgs), "; ".join(m for m in msgs if m)
    return pi


def _oracle(pi, q, k, v, tq, tk, tv, go):
    from oracle import attention_oracle as ao
    return ao.layer_fwd_bwd(q.double(), k.double(), v.double(), pi.index_0_offsets.cpu().long(), pi.index_1.cpu().long(), tq.double(),
                            tk.double(), tv.double(), pi.rel_idx.cpu().long(), go.double())


@pytest.mark.parametrize("parity", [0, 1])
@pytest.mark.parametrize("case", [(2500, 0.32, 0.02, 3, False), (1500, 0.64, 0.04, 2, False), (2000, 0.32, 0.02, 1, True)])
def test_fused_matches_oracle_small(case, parity):
    n_pts, window, quant, h, lattice = case
    xyz, offset, ds = _small_scene(n_pts, 11, lattice)
    pi = _fused_vs(_oracle, xyz, offset, ds, window, quant, parity, h, 5)
    if window > 0.5:
        assert pi.plan.totals[3] > 64, "case must exercise windows larger than one block"


def test_fused_dense_only_matches_oracle():
    """dense-window pairs only (downsample_idx=None): the dense pass finalises by itself"""
    xyz, offset, _ = _small_scene(2000, 21)
    _fused_vs(_oracle, xyz, offset, None, 0.32, 0.02, 0, 2, 9)


def _ref_kernels(pi, q, k, v, tq, tk, tv, go):
    """the reference's own CUDA kernels (+ our segment softmax, whose parity is pinned separately) on the CSR of the same index"""
    from oracle import ref_cuda
    from stratified_transformer_b200 import pointops2_cuda as ext
    off, i1, rel = pi.index_0_offsets, pi.index_1, pi.rel_idx.contiguous()
    N, h, _ = q.shape
    M = i1.numel()
    qd, kd, vd, tqd, tkd, tvd, god = (t.cuda().contiguous() for t in (q, k, v, tq, tk, tv, go))
    s = ref_cuda.step1_fwd(qd, kd, off, i1) + ref_cuda.rpe_fwd(qd, kd, off, i1, tqd, tkd, rel)
    p = torch.empty_like(s)
    ext.segment_softmax_forward_cuda(N, M, h, s, None, off, p)
    out = ref_cuda.step2_rpv_fwd(p, vd, off, i1, tvd, rel)
    gp, gv, gtv = ref_cuda.step2_rpv_bwd(god, p, vd, off, i1, tvd, rel)
    gs = torch.empty_like(p)
    ext.segment_softmax_backward_cuda(N, M, h, p, gp, off, gs)
    gq1, gk1 = ref_cuda.step1_bwd(gs, qd, kd, off, i1)
    gq2, gk2, gtq, gtk = ref_cuda.rpe_bwd(gs, qd, kd, off, i1, tqd, tkd, rel)
    return dict(out=out, gq=gq1 + gq2, gk=gk1 + gk2, gv=gv, gtq=gtq, gtk=gtk, gtv=gtv)


@pytest.fixture(scope="module")
def hierarchy():
    """one 80k-point scene and its three TransitionDown levels (n -> int(n/4)+1 by FPS), as bench.py builds them"""
    from stratified_transformer_b200 import pointops
    from stratified_transformer_b200.synthetic import make_batch
    xyz, _, offset = make_batch(1, 80000, seed0=5)
    xd, od = torch.from_numpy(xyz).cuda(), torch.from_numpy(offset).cuda()
    levels = [(xd, od)]
    for _ in range(3):
        counts = torch.diff(od, prepend=od.new_zeros(1))
        new_off = torch.cumsum((counts.double() * 0.25).long() + 1, 0).int()
        sub = pointops.furthestsampling(xd, od, new_off)
        xd, od = xd[sub.long()].contiguous(), new_off
        levels.append((xd, od))
    return levels


@pytest.mark.parametrize("level", [0, 1, 2, 3])
@pytest.mark.parametrize("parity", [0, 1])
def test_fused_matches_reference_kernels_full_size(hierarchy, level, parity):
    from oracle import ref_cuda
    if not ref_cuda.available():
        pytest.skip("oracle/_ref/libpointops2_ref.so not built")
    from stratified_transformer_b200 import index, pointops
    xd, od = hierarchy[level]
    window, quant, h = 0.16 * 2 ** level, 0.01 * 2 ** level, 3 * 2 ** level
    ds = pointops.furthestsampling(xd, od, index.fps_new_offset(od, 8))
    # table gradients are sums of 1e6-1e7 fp32 terms here, which the reference accumulates with float atomics in a
    # non-deterministic order: besides the elementwise bound they get 1e-5 of the tensor's largest magnitude
    _fused_vs(_ref_kernels, xd.cpu().numpy(), od.cpu().numpy(), ds.cpu().numpy(), window, quant, parity, h, 7 + level, table_scale=0.1,
              table_grad_floor=1e-5)


def test_module_plan_path_matches_per_op_path():
    from stratified_transformer_b200 import index
    from stratified_transformer_b200.window_attention import WindowAttention
    xyz, offset, ds = _small_scene(3000, 31)
    xd, od = torch.from_numpy(xyz).cuda(), torch.from_numpy(offset).cuda()
    pi = index.build_stratified_index(xd, od, 0.32, 0.02, torch.from_numpy(ds).cuda(), 1, fused=True)
    torch.manual_seed(0)
    attn = WindowAttention(48, 0.32, 3, 0.02, rel_query=True, rel_key=True, rel_value=True).cuda()
    for t in (attn.relative_pos_query_table, attn.relative_pos_key_table, attn.relative_pos_value_table):
        torch.nn.init.uniform_(t, -0.3, 0.3)
    feats = torch.randn(xyz.shape[0], 48, device="cuda")
    res = {}
    for mode in ("plan", "per_op"):
        attn.per_op = mode == "per_op"
        attn.zero_grad(set_to_none=True)
        f = feats.clone().requires_grad_(True)
        y = attn(f, xd, pi)
        y.square().sum().backward()
        res[mode] = dict(y=y.detach(), gf=f.grad, **{n: p.grad.clone() for n, p in attn.named_parameters()})
    for name in res["plan"]:
        _close(res["plan"][name], res["per_op"][name], 5e-4, name)


@pytest.mark.parametrize("level", [0, 1, 2, 3])
@pytest.mark.parametrize("parity", [0, 1])
def test_per_op_benched_path_matches_reference_kernels_full_size(hierarchy, level, parity):
    """The code path bench.py times by default (window_logits / segment_softmax / window_aggregate and their `_ws` backward
    with packed bins, window row order, transposed CSR and both length orders) on builder-produced indices of one 80k-point
    scene at the four layer shapes, every output and gradient against the reference's own kernels."""
    from oracle import ref_cuda
    if not ref_cuda.available():
        pytest.skip("oracle/_ref/libpointops2_ref.so not built")
    from stratified_transformer_b200 import index, pointops
    xd, od = hierarchy[level]
    window, quant, h = 0.16 * 2 ** level, 0.01 * 2 ** level, 3 * 2 ** level
    ds = pointops.furthestsampling(xd, od, index.fps_new_offset(od, 8))
    pi = index.build_stratified_index(xd, od, window, quant, ds, parity)
    N, L = xd.shape[0], 64
    g = torch.Generator().manual_seed(17 + level)
    q, k, v, go = (torch.randn(N, h, 16, generator=g) for _ in range(4))
    q = q * 0.5
    tq, tk, tv = ((torch.rand(L, h, 16, 3, generator=g) - 0.5) * 0.2 for _ in range(3))
    leaves = [t.cuda().requires_grad_(True) for t in (q, k, v, tq, tk, tv)]
    s = pointops.window_logits(leaves[0], leaves[1], leaves[3], leaves[4], pi)
    p = pointops.segment_softmax(s, pi.index_0_offsets)
    out = pointops.window_aggregate(p, leaves[2], leaves[5], pi)
    out.backward(go.cuda())
    got = dict(out=out, gq=leaves[0].grad, gk=leaves[1].grad, gv=leaves[2].grad, gtq=leaves[3].grad, gtk=leaves[4].grad, gtv=leaves[5].grad)
    want = _ref_kernels(pi, q, k, v, tq, tk, tv, go)
    msgs = [_check(val, want[name], 2e-4 if name.startswith("gt") else 1e-4, f"{name} (level {level}, parity {parity})",
                   1e-5 if name.startswith("gt") else 0.0) for name, val in got.items()]
    assert not any(msgs), "; ".join(m for m in msgs if m)


@pytest.mark.parametrize("parity", [0, 1])
def test_swin_invariant_and_window_membership_full_size(hierarchy, parity):
    """Dense-only pairs at full size (80k points): the reference's own invariant M == (counts**2).sum()
    (model/swin3d_transformer.py:259,280), and a brute-force check that every emitted pair shares a window id computed in
    fp64 — except where the fp32 voxel arithmetic of torch_cluster legitimately differs from fp64 at a window boundary."""
    from stratified_transformer_b200 import index
    xd, od = hierarchy[0]
    w = 0.16
    pi = index.build_stratified_index(xd, od, w, 0.01, None, parity, want_index_0=True)
    wo = pi.win_offsets.long()
    counts = wo[1:] - wo[:-1]
    assert pi.M == int((counts ** 2).sum())
    x = xd.double().cpu()
    shift = 0.5 * w if parity else 0.0
    cell = torch.floor((x + shift - x.min(0).values) / w).long()
    i0, i1 = pi.index_0.long().cpu(), pi.index_1.long().cpu()
    differ = (cell[i0] != cell[i1]).any(1)
    # pairs whose fp64 window ids differ must have an endpoint within a few fp32 ulps of a window boundary
    frac = ((x + shift - x.min(0).values) / w)
    near = ((frac - frac.round()).abs() < 1e-5).any(1)
    assert bool((near[i0[differ]] | near[i1[differ]]).all())
    assert int(differ.sum()) < 1e-3 * pi.M
