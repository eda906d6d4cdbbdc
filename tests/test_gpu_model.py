"""Whole-network wiring (-m gpu): the reference's own `Stratified` (model/stratified_transformer.py:399-505; text
extracted into the git-ignored oracle/_ref/ref_model_native.py, running on the reference's own autograd functions and kernels)
against `stratified_transformer_b200.model.Stratified` with the same state dict: logits, offset regression, gradients.
KPConvLayer / FastBatchNorm1d are this package's restatement on BOTH sides (torch_points3d is absent: parity unpinned), so
the stem's convolution itself is not what is being compared - everything else is."""
import importlib.util
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _native_model():
    path = os.path.join(ROOT, "oracle", "_ref", "ref_model_native.py")
    if not (os.path.exists(path) and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libpointops2_ref.so"))):
        pytest.skip("oracle/_ref/ref_model_native.py / libpointops2_ref.so not built (needs /root/reference at build time)")
    spec = importlib.util.spec_from_file_location("ref_model_native", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


CFG = dict(downsample_scale=8, depths=[2, 2, 2, 2], channels=[48, 96, 192, 384], num_heads=[3, 6, 12, 24],
           window_size=[0.16, 0.32, 0.64, 1.28], up_k=3, grid_sizes=[0.04, 0.08, 0.16, 0.32], quant_sizes=[0.01, 0.02, 0.04, 0.08],
           rel_query=True, rel_key=True, rel_value=True, drop_path_rate=0.0, num_layers=4, concat_xyz=True, num_classes=13,
           ratio=0.25, k=16, prev_grid_size=0.04, sigma=1.0, stem_transformer=False)


def _inputs():
    from stratified_transformer_b200 import prestep
    from stratified_transformer_b200.synthetic import make_batch
    xyz, rgb, offset = make_batch(2, 6000, seed0=13, n_raw=150000)
    xd, od = torch.from_numpy(xyz).cuda(), torch.from_numpy(offset).cuda()
    feat = torch.cat([torch.from_numpy(rgb).cuda().float(), xd], 1)
    batch = prestep.batch_from_offset(od)
    nbr = prestep.ball_query(2.5 * 0.04, 34, xd, xd, mode="partial_dense", batch_x=batch, batch_y=batch)[0]
    return feat, xd, od, batch, nbr


def test_whole_model_matches_reference_wiring():
    from stratified_transformer_b200.model import Stratified
    ref = _native_model()
    torch.manual_seed(4)
    theirs = ref.Stratified(**CFG).cuda()
    for name, p in theirs.named_parameters():
        if "relative_pos" in name:
            torch.nn.init.uniform_(p, -0.2, 0.2)
    mine = Stratified(**CFG).cuda()
    missing = mine.load_state_dict(theirs.state_dict())     # strict: the same parameter and buffer names everywhere
    assert not missing.missing_keys and not missing.unexpected_keys
    feat, xd, od, batch, nbr = _inputs()

    def run(model):
        model.train()
        model.zero_grad(set_to_none=True)
        out, shift = model(feat, xd, od, batch, nbr)
        (out.square().mean() + shift.square().mean()).backward()
        return out.detach(), shift.detach(), {n: p.grad.detach() for n, p in model.named_parameters() if p.grad is not None}

    w_out, w_shift, w_g = run(theirs)
    g_out, g_shift, g_g = run(mine)
    assert g_out.shape == (xd.shape[0], 13) and g_shift.shape == (xd.shape[0], 3)

    def close(a, b, name, tol):
        scale = max(1.0, float(b.abs().max()))
        err = float((a - b).abs().max())
        assert err <= tol * scale, f"{name}: max err {err:.3e} at scale {scale:.2f}"
    close(g_out, w_out, "logits", 1e-3)
    close(g_shift, w_shift, "shift", 1e-3)
    assert set(g_g) == set(w_g)
    worst = max((float((g_g[n] - w_g[n]).abs().max()) / max(1e-6, float(w_g[n].abs().max())), n) for n in w_g)
    assert worst[0] < 2e-2, f"gradient of {worst[1]}: relative max err {worst[0]:.3e}"


@pytest.mark.parametrize("cin,cout", [(6, 48), (12, 12)])
def test_kpconv_kernel_matches_the_torch_restatement(cin, cout):
    """KPConvLayer.forward (neighbourhood kernel + one GEMM) against KPConvLayer.forward_torch, forward and gradients"""
    from stratified_transformer_b200 import prestep
    from stratified_transformer_b200.model import KPConvLayer
    from stratified_transformer_b200.synthetic import make_batch
    xyz, _, offset = make_batch(2, 4000, seed0=3, n_raw=100000)
    xd, od = torch.from_numpy(xyz).cuda(), torch.from_numpy(offset).cuda()
    batch = prestep.batch_from_offset(od)
    nbr = prestep.ball_query(0.1, 34, xd, xd, batch_x=batch, batch_y=batch)[0]
    torch.manual_seed(0)
    layer = KPConvLayer(cin, cout, point_influence=0.04).cuda()
    x1 = torch.randn(xd.shape[0], cin, device="cuda", requires_grad=True)
    x2 = x1.detach().clone().requires_grad_(True)
    g = torch.randn(xd.shape[0], cout, device="cuda")
    y1 = layer(xd, xd, nbr, x1)
    y1.backward(g)
    gw1 = layer.weight.grad.clone()
    layer.zero_grad()
    y2 = layer.forward_torch(xd, xd, nbr, x2)
    y2.backward(g)
    for a, b, name in ((y1, y2, "out"), (x1.grad, x2.grad, "grad_feats"), (gw1, layer.weight.grad, "grad_weight")):
        scale = max(1.0, float(b.abs().max()))
        assert float((a - b).abs().max()) <= 2e-4 * scale, name


def test_prefetched_geometry_gives_the_same_network_output():
    """model.GeometryChain (sampling + pair lists + neighbour lists of a batch on a side stream) against the inline construction"""
    from stratified_transformer_b200.model import GeometryChain, Stratified
    torch.manual_seed(5)
    model = Stratified(**CFG).cuda()
    feat, xd, od, batch, nbr = _inputs()
    with torch.no_grad():
        want_out, want_shift = model(feat, xd, od, batch, nbr)
        chain = GeometryChain(model)
        for _ in range(2):                       # twice: the second round reuses the chain's stream and buffers
            chain.submit(xd, od, 2.5 * 0.04)
            geo = chain.take()
            assert torch.equal(geo["batch"], batch) and torch.equal(geo["neighbor_idx"], nbr)
            out, shift = model(feat, xd, od, geo["batch"], geo["neighbor_idx"], geometry=geo)
            torch.cuda.synchronize()
            assert torch.equal(out, want_out) and torch.equal(shift, want_shift)
