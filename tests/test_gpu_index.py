"""Pair-index construction on the GPU (-m gpu): bit-exact against the reference-generated golden
fixtures (tests/golden/index_*.npz) and against the CPU oracle at larger sizes."""
import os

import numpy as np
import pytest
import torch

from oracle import fps_oracle, index_oracle as io

pytestmark = pytest.mark.gpu
CASES = ["s3dis_small", "s3dis_lattice", "scannet_small"]


def build(xyz, offset, w, quant, ds_idx, parity, want_index_0=False):
    from stratified_transformer_b200 import index
    return index.build_stratified_index(torch.from_numpy(xyz).cuda(), torch.from_numpy(offset).cuda(), w, quant,
                                        None if ds_idx is None else torch.from_numpy(ds_idx).cuda(), parity,
                                        want_index_0=want_index_0)


@pytest.mark.parametrize("name", CASES)
def test_pairs_match_reference_golden(golden_dir, name):
    from stratified_transformer_b200 import index, pointops
    g = np.load(os.path.join(golden_dir, f"index_{name}.npz"))
    xyz, offset = g["xyz"], g["offset"]
    w, quant, ds = float(g["window_size"]), float(g["quant_size"]), int(g["downsample_scale"])
    off_d = torch.from_numpy(offset).cuda()
    new_offset = index.fps_new_offset(off_d, ds)
    assert np.array_equal(new_offset.cpu().numpy(), g["new_offset"])
    ds_idx = pointops.furthestsampling(torch.from_numpy(xyz).cuda(), off_d, new_offset).cpu().numpy()
    assert np.array_equal(ds_idx, g["downsample_idx"])
    for parity in (0, 1):
        index.set_torch_semantics("cpu")        # the goldens come from CPU torch (no GPU where they were generated)
        try:
            r = build(xyz, offset, w, quant, ds_idx, parity, want_index_0=True)
            torch.cuda.synchronize()
        finally:
            index.set_torch_semantics("cuda")
        offs = r.index_0_offsets.cpu().numpy()
        assert np.array_equal(offs, g[f"p{parity}_offsets"])
        assert r.n_max == int(g[f"p{parity}_n_max"]) and r.M == offs[-1]
        i1 = r.index_1.cpu().numpy()
        # multiset per query == the reference's (canonicalised because its sorts are unstable)
        assert np.array_equal(io.canonicalize(offs, i1), g[f"p{parity}_index_1"])
        # exact emitted order == the oracle's canonical order (dense ascending, then sparse ascending)
        o = io.build_layer_index(xyz, offset, w, ds, ds_idx, parity)
        assert np.array_equal(i1, o["index_1"])
        assert np.array_equal(r.index_0.cpu().numpy(), o["index_0"])
        rel = r.rel_idx.cpu().numpy()
        assert np.array_equal(rel, io.rel_pos_index_stratified(xyz, o["index_0"], o["index_1"], w, quant, device="cpu"))
        # and the rel-pos index agrees with the reference's torch evaluation once both are canonicalised
        seg = np.repeat(np.arange(offs.shape[0] - 1), np.diff(offs))
        order = np.lexsort((i1, seg))
        want = g[f"p{parity}_rel_idx"].astype(np.int32)
        # row-aligned: after the (query, key) sort the golden's rows line up with ours (duplicate (query, key) pairs carry
        # the same rel-pos index), so the comparison is exact equality, not a multiset
        assert np.array_equal(rel[order], want)


def test_pairs_vs_oracle_two_scenes_20k():
    from stratified_transformer_b200.synthetic import make_batch
    xyz, _, offset = make_batch(2, 20000, seed0=3, n_raw=600_000)
    w, quant, ds = 0.32, 0.02, 8      # S3DIS layer-1 geometry
    new_offset = io.fps_new_offset(offset, ds)
    ds_idx = fps_oracle.furthestsampling(xyz, offset, new_offset)
    for parity in (0, 1):
        r = build(xyz, offset, w, quant, ds_idx, parity)
        o = io.build_layer_index(xyz, offset, w, ds, ds_idx, parity)
        assert np.array_equal(r.index_0_offsets.cpu().numpy(), o["offsets"])
        assert np.array_equal(r.index_1.cpu().numpy(), o["index_1"])
        assert r.n_max == o["n_max"]
        assert np.array_equal(r.rel_idx.cpu().numpy(), io.rel_pos_index_stratified(xyz, o["index_0"], o["index_1"], w, quant))


def test_dense_only_pairs_swin_invariant(golden_dir):
    """Swin builds dense window pairs only; the reference asserts M == (counts**2).sum()
    (model/swin3d_transformer.py:258-259)."""
    g = np.load(os.path.join(golden_dir, "index_s3dis_small.npz"))
    xyz, offset = g["xyz"], g["offset"]
    for parity in (0, 1):
        r = build(xyz, offset, 0.16, None, None, parity)
        batch = io.batch_from_offset(offset)
        w = np.full(3, 0.16, np.float32)
        if parity == 0:
            _, _, counts = io.grid_sample(xyz, batch, w, None)
        else:
            _, _, counts = io.grid_sample((xyz + np.float32(0.5) * w).astype(np.float32), batch, w, xyz.min(0))
        assert r.M == int((counts ** 2).sum())
        assert r.rel_idx is None


def test_rel_pos_index_standalone_and_swin(golden_dir):
    from stratified_transformer_b200 import index
    g = np.load(os.path.join(golden_dir, "relidx_swin.npz"))
    xyz, i0, i1 = g["xyz"], g["index_0"].astype(np.int64), g["index_1"].astype(np.int64)
    N = xyz.shape[0]
    offsets = np.concatenate([[0], np.cumsum(np.bincount(i0, minlength=N))]).astype(np.int32)
    xd, od, i1d = torch.from_numpy(xyz).cuda(), torch.from_numpy(offsets).cuda(), torch.from_numpy(i1).cuda().int()
    for shift, tag in ((0.0, "noshift"), (0.08, "shift")):
        got = index.rel_pos_index_swin(xd, od, i1d, 0.16, 0.01, shift).cpu().numpy()
        assert np.array_equal(got, g[f"rel_idx_{tag}"].astype(np.int32))
    got = index.rel_pos_index_stratified(xd, od, i1d, 4.0, 0.25).cpu().numpy()   # any geometry: arithmetic check only
    assert np.array_equal(got, io.rel_pos_index_stratified(xyz, i0, i1, 4.0, 0.25))


def test_window_attention_module_vs_oracle(golden_dir):
    """The WindowAttention mirror (reference call signature AND PairIndex form) against a CPU
    composition of the oracle ops around the same Linear layers."""
    from oracle import attention_oracle as ao
    from stratified_transformer_b200 import index
    from stratified_transformer_b200.window_attention import WindowAttention
    g = np.load(os.path.join(golden_dir, "index_s3dis_small.npz"))
    xyz, offset = g["xyz"], g["offset"]
    ds_idx = g["downsample_idx"]
    torch.manual_seed(0)
    C, h = 48, 3
    mod = WindowAttention(C, 0.16, h, 0.01, rel_query=True, rel_key=True, rel_value=True).cuda()
    feats = torch.randn(xyz.shape[0], C)
    pi = build(xyz, offset, 0.16, 0.01, ds_idx, 1, want_index_0=True)
    xd = torch.from_numpy(xyz).cuda()
    fd = feats.cuda().requires_grad_(True)
    out = mod(fd, xd, pi)
    # reference-style call with int64 tensors gives the same result
    out2 = mod(fd, xd, pi.index_0.long(), pi.index_1.long(), pi.index_0_offsets.long(), torch.tensor(pi.n_max).cuda())
    assert torch.allclose(out, out2, atol=1e-6)
    gout = torch.randn_like(out)
    out.backward(gout)
    # CPU: same Linear weights, oracle ops in fp64 with autograd through layer_autograd's formulation
    W = {k: v.detach().double().cpu() for k, v in mod.state_dict().items()}
    f = feats.double().requires_grad_(True)
    qkv = (f @ W["qkv.weight"].T + W["qkv.bias"]).reshape(-1, 3, h, C // h).permute(1, 0, 2, 3)
    q, k, v = qkv[0] * mod.scale, qkv[1], qkv[2]
    i0 = pi.index_0.long().cpu(); i1 = pi.index_1.long().cpu(); r = pi.rel_idx.long().cpu()
    tq, tk, tv = (W[f"relative_pos_{n}_table"] for n in ("query", "key", "value"))
    eq = ao.table_sum(tq, r); ek = ao.table_sum(tk, r); ev = ao.table_sum(tv, r)
    s = (q[i0] * k[i1]).sum(-1) + (q[i0] * eq + k[i1] * ek).sum(-1)
    p = ao.softmax_fwd(s.detach(), i0, f.shape[0])
    ex = torch.exp(s - s.detach().new_full((f.shape[0], h), -1e30).scatter_reduce(0, i0[:, None].expand(-1, h), s.detach(), "amax")[i0])
    p = ex / torch.zeros(f.shape[0], h, dtype=torch.float64).index_add(0, i0, ex)[i0]
    x = torch.zeros(f.shape[0], h, C // h, dtype=torch.float64).index_add(0, i0, p.unsqueeze(-1) * (v[i1] + ev))
    ref = x.reshape(-1, C) @ W["proj.weight"].T + W["proj.bias"]
    ref.backward(gout.double().cpu())
    assert torch.allclose(out.detach().double().cpu(), ref.detach(), rtol=2e-4, atol=2e-5)
    assert torch.allclose(fd.grad.double().cpu(), f.grad, rtol=2e-3, atol=2e-5)


def test_swin_window_attention_module(golden_dir):
    """3DSwin variant: dense-only pairs from the builder, Swin rel-pos index, module forward/backward against the
    oracle composition."""
    from oracle import attention_oracle as ao
    from stratified_transformer_b200.window_attention import SwinWindowAttention
    g = np.load(os.path.join(golden_dir, "index_s3dis_small.npz"))
    xyz, offset = g["xyz"], g["offset"]
    torch.manual_seed(1)
    C, h, w, quant = 96, 6, 0.16, 0.01
    mod = SwinWindowAttention(C, w, h, quant, rel_query=True, rel_key=True, rel_value=True).cuda()
    assert mod.relative_pos_query_table.shape[0] == 31
    feats = torch.randn(xyz.shape[0], C)
    pi = build(xyz, offset, w, None, None, 1, want_index_0=True)      # shifted windows
    shift = 0.5 * np.float32(w)
    fd = feats.cuda().requires_grad_(True)
    out = mod(fd, torch.from_numpy(xyz).cuda(), pi.index_0.long(), pi.index_0_offsets.long(), pi.n_max, pi.index_1.long(),
              torch.full((3,), float(shift)).cuda())
    gout = torch.randn_like(out)
    out.backward(gout)
    i0 = pi.index_0.long().cpu(); i1 = pi.index_1.long().cpu()
    r = torch.from_numpy(io.rel_pos_index_swin(xyz, i0.numpy(), i1.numpy(), w, quant, float(shift))).long()
    assert int(r.min()) >= 0 and int(r.max()) <= 30
    W = {k: v.detach().double().cpu() for k, v in mod.state_dict().items()}
    f = feats.double().requires_grad_(True)
    qkv = (f @ W["qkv.weight"].T + W["qkv.bias"]).reshape(-1, 3, h, C // h).permute(1, 0, 2, 3)
    q, k, v = qkv[0] * mod.scale, qkv[1], qkv[2]
    tq, tk, tv = (W[f"relative_pos_{n}_table"] for n in ("query", "key", "value"))
    s = (q[i0] * k[i1]).sum(-1) + (q[i0] * ao.table_sum(tq, r) + k[i1] * ao.table_sum(tk, r)).sum(-1)
    mx = s.detach().new_full((f.shape[0], h), -1e30).scatter_reduce(0, i0[:, None].expand(-1, h), s.detach(), "amax")
    ex = torch.exp(s - mx[i0])
    p = ex / torch.zeros(f.shape[0], h, dtype=torch.float64).index_add(0, i0, ex)[i0]
    x = torch.zeros(f.shape[0], h, C // h, dtype=torch.float64).index_add(0, i0, p.unsqueeze(-1) * (v[i1] + ao.table_sum(tv, r)))
    ref = x.reshape(-1, C) @ W["proj.weight"].T + W["proj.bias"]
    ref.backward(gout.double().cpu())
    assert torch.allclose(out.detach().double().cpu(), ref.detach(), rtol=2e-4, atol=2e-5)
    assert torch.allclose(fd.grad.double().cpu(), f.grad, rtol=2e-3, atol=2e-5)


@pytest.mark.parametrize("name,C,h", [("s3dis_small", 48, 3), ("s3dis_lattice", 96, 6), ("scannet_small", 96, 6)])
def test_fused_window_forward_vs_per_pair_path(golden_dir, name, C, h):
    """Fused (tensor-core, per-window) forward == the per-pair path on builder-produced pair lists, including the
    lattice scene (windows with differing key lists -> per-pair fallback rows) and windows above the key limit."""
    from stratified_transformer_b200 import pointops
    g = np.load(os.path.join(golden_dir, f"index_{name}.npz"))
    xyz, offset = g["xyz"], g["offset"]
    w, quant = float(g["window_size"]), float(g["quant_size"])
    torch.manual_seed(2)
    N, d = xyz.shape[0], C // h
    L = 2 * int((2 * w + 1e-4) // quant)
    for parity in (0, 1):
        pi = build(xyz, offset, w, quant, g["downsample_idx"], parity)
        flags, rows = pi.fused_plan()
        q, k, v = (torch.randn(N, h, d, device="cuda").requires_grad_(True) for _ in range(3))
        tq, tk, tv = (torch.randn(L, h, d, 3, device="cuda").mul_(0.5).requires_grad_(True) for _ in range(3))
        gout = torch.randn(N, h, d, device="cuda")
        out = pointops.window_attention_fused(q, k, v, tq, tk, tv, pi)
        out.backward(gout)
        got = [out.detach()] + [t.grad.clone() for t in (q, k, v, tq, tk, tv)]
        for t in (q, k, v, tq, tk, tv):
            t.grad = None
        s = pointops.window_logits(q, k, tq, tk, pi)
        p = pointops.segment_softmax(s, pi.index_0_offsets)
        ref = pointops.window_aggregate(p, v, tv, pi)
        ref.backward(gout)
        want = [ref.detach()] + [t.grad for t in (q, k, v, tq, tk, tv)]
        for a, b, nm in zip(got, want, ("out", "gq", "gk", "gv", "gtq", "gtk", "gtv")):
            err = (a - b).abs().max().item()
            assert err <= 2e-4 * max(1.0, b.abs().max().item()), (name, parity, nm, err)
        frac = float(flags.float().mean())
        print(f"{name} parity {parity}: {int(flags.sum())}/{flags.numel()} windows fused ({frac:.2%}), {rows.numel()} fallback rows")
        # shifted windows (parity 1) straddle the boundaries of the shifted 2x windows, so many of them have
        # per-query sparse key lists and take the per-pair fallback; unshifted ones are (almost) all fused
        assert frac > (0.9 if parity == 0 else 0.1)
        if parity == 1:
            assert rows.numel() > 0


def test_geometry_prefetcher_matches_serial_builder(golden_dir):
    """Split-phase construction on a side stream (what bench.py and a training loop use) gives the same pair lists,
    transposed CSR and packed bins as the serial builder."""
    from stratified_transformer_b200 import index
    g = np.load(os.path.join(golden_dir, "index_s3dis_small.npz"))
    xyz = torch.from_numpy(g["xyz"]).cuda()
    off = torch.from_numpy(g["offset"]).cuda()
    w, quant, ds = float(g["window_size"]), float(g["quant_size"]), int(g["downsample_scale"])
    L = 2 * int((2 * w + 1e-4) // quant)
    serial = index.build_layer_index(xyz, off, w, quant, ds)
    pf = index.GeometryPrefetcher([(w, quant, ds, L)])
    for _ in range(3):      # several rounds: buffers are recycled across submissions
        pf.submit([xyz], [off], [g["offset"].tolist()])
        pf.complete()
        (li,) = pf.take()
        torch.cuda.synchronize()
        assert torch.equal(li.downsample_idx, serial.downsample_idx)
        for p in (0, 1):
            a, b = li.for_block(p), serial.for_block(p)
            assert a.M == b.M and a.n_max == b.n_max and a.n_win == b.n_win
            for name in ("index_0_offsets", "index_1", "rel_idx", "row_order", "win_offsets"):
                assert torch.equal(getattr(a, name), getattr(b, name)), name
            ta, tb = a.tcsr, b.tcsr
            assert torch.equal(ta.t_offsets, tb.t_offsets) and torch.equal(ta.t_pair, tb.t_pair) and torch.equal(ta.t_index0, tb.t_index0)


def test_window_attention_under_autocast(golden_dir):
    """AMP recipe of the reference (use_amp): Linear layers in bf16, pair ops pinned to fp32 by the wrappers."""
    from stratified_transformer_b200.window_attention import WindowAttention
    g = np.load(os.path.join(golden_dir, "index_s3dis_small.npz"))
    xyz, offset = g["xyz"], g["offset"]
    torch.manual_seed(0)
    mod = WindowAttention(48, 0.16, 3, 0.01, rel_query=True, rel_key=True, rel_value=True).cuda()
    pi = build(xyz, offset, 0.16, 0.01, g["downsample_idx"], 0)
    feats = torch.randn(xyz.shape[0], 48, device="cuda", requires_grad=True)
    xd = torch.from_numpy(xyz).cuda()
    ref = mod(feats, xd, pi)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = mod(feats, xd, pi)
    out.float().sum().backward()
    assert out.dtype == torch.bfloat16 and feats.grad is not None and torch.isfinite(feats.grad).all()
    assert (out.float() - ref).abs().max() < 0.05 * ref.abs().max()
    assert mod.relative_pos_query_table.grad.dtype == torch.float32


def test_length_order_is_a_stable_sort_by_pair_count(golden_dir):
    """stb200_length_order: a permutation of the rows, non-decreasing pair count, ties in base (window) order."""
    g = np.load(os.path.join(golden_dir, "index_s3dis_small.npz"))
    pi = build(g["xyz"], g["offset"], float(g["window_size"]), float(g["quant_size"]), g["downsample_idx"], 1)
    off = pi.index_0_offsets.cpu().numpy().astype(np.int64)
    lens = off[1:] - off[:-1]
    base = pi.row_order.cpu().numpy()
    want = base[np.argsort(lens[base], kind="stable")]
    q_order, k_order = pi.len_orders
    assert np.array_equal(q_order.cpu().numpy(), want)
    t_off = pi.tcsr.t_offsets.cpu().numpy().astype(np.int64)
    t_lens = t_off[1:] - t_off[:-1]
    assert np.array_equal(k_order.cpu().numpy(), base[np.argsort(t_lens[base], kind="stable")])


def test_rel_pos_index_matches_the_reference_statements_on_cuda():
    """The reference as it really runs: its three torch statements (model/stratified_transformer.py:186-188) evaluated on
    CUDA tensors on this box, against the builder's rel-pos index (default "cuda" arithmetic) — exact, on 1.5 M pairs; and
    the numpy oracle in the same mode.  CPU torch differs from both in a few pairs (scalar division = reciprocal multiply on
    CUDA), which is why the CPU-generated goldens are compared in the "cpu" mode."""
    from stratified_transformer_b200.synthetic import make_batch
    xyz, _, offset = make_batch(2, 20000, seed0=3, n_raw=600_000)
    for w, quant in ((0.16, 0.01), (0.32, 0.02), (0.2, 0.01)):
        ds_idx = fps_oracle.furthestsampling(xyz, offset, io.fps_new_offset(offset, 8))
        r = build(xyz, offset, w, quant, ds_idx, 1, want_index_0=True)
        xd = torch.from_numpy(xyz).cuda()
        i0, i1 = r.index_0.long(), r.index_1.long()
        rel = xd[i0] - xd[i1]
        rel = torch.round(rel * 100000) / 100000
        want = ((rel + 2 * w - 0.0001) // quant).int()
        assert torch.equal(r.rel_idx, want)
        assert np.array_equal(want.cpu().numpy(), io.rel_pos_index_stratified(xyz, i0.cpu().numpy(), i1.cpu().numpy(), w, quant, device="cuda"))
        cpu = ((torch.round((xd.cpu()[i0.cpu()] - xd.cpu()[i1.cpu()]) * 100000) / 100000 + 2 * w - 0.0001) // quant).int()
        assert int((cpu != want.cpu()).sum()) < 1e-3 * want.numel()
