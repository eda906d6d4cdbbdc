"""Layer-level drop-in proof (-m gpu): the REFERENCE's own `BasicLayer` / `SwinTransformerBlock` / `TransitionDown` / `Upsample`
(model/stratified_transformer.py:67-112, 219-342; text extracted at build time into the git-ignored oracle/_ref/ref_layers.py with
only its imports redirected) run over this package's operators, next to the mirrors in stratified_transformer_b200/layers.py:
same state dict, same inputs, all outputs, all parameter and input gradients.  The reference side derives its pair lists with
its own Python (grid_sample / get_indice_pairs / sort / bincount, lines 267-317), the mirror with the device builder.
`voxel_grid` under the reference side is the oracle's restatement (parity unpinned)."""
import importlib.util
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GEN = os.path.join(ROOT, "oracle", "_ref", "ref_layers.py")


def _ref():
    if not os.path.exists(GEN):
        pytest.skip("oracle/_ref/ref_layers.py not generated (needs /root/reference at build time)")
    spec = importlib.util.spec_from_file_location("ref_layers", GEN)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _native():
    """the WHOLE reference layer: its Python, its autograd functions (functions/pointops.py), its own kernels (libpointops2_ref.so)"""
    path = os.path.join(ROOT, "oracle", "_ref", "ref_layers_native.py")
    if not (os.path.exists(path) and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libpointops2_ref.so"))):
        pytest.skip("oracle/_ref/ref_layers_native.py / libpointops2_ref.so not built (needs /root/reference at build time)")
    spec = importlib.util.spec_from_file_location("ref_layers_native", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _close(got, want, name, tol=2e-4):
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    scale = max(1.0, float(want.abs().max()))
    err = float((got - want).abs().max())
    assert err <= tol * scale, f"{name}: max err {err:.3e} at scale {scale:.2f}"


def _scene(n_scenes, n_pts, seed):
    from stratified_transformer_b200.synthetic import make_batch
    xyz, _, offset = make_batch(n_scenes, n_pts, seed0=seed, n_raw=80000)
    return torch.from_numpy(xyz).cuda(), torch.from_numpy(offset).cuda()


@pytest.mark.parametrize("depth,fused", [(2, False), (3, True)])
def test_reference_basic_layer_matches_mirror(depth, fused):
    from stratified_transformer_b200 import layers
    ref = _ref()
    xyz, offset = _scene(2, 2500, 5)
    C, h, window, quant, ds = 48, 3, 0.32, 0.02, 8
    torch.manual_seed(1)
    kw = dict(rel_query=True, rel_key=True, rel_value=True, drop_path=0.0, ratio=0.25, k=16, out_channels=96)
    theirs = ref.BasicLayer(ds, depth, C, h, window, 0.04, quant, downsample=ref.TransitionDown, **kw).cuda()
    for name, p in theirs.named_parameters():
        if "relative_pos" in name:
            torch.nn.init.uniform_(p, -0.3, 0.3)
    mine = layers.BasicLayer(ds, depth, C, h, window, 0.04, quant, downsample=layers.TransitionDown, **kw).cuda()
    mine.load_state_dict(theirs.state_dict())          # strict: same names and shapes
    mine.fused = fused
    feats = torch.randn(xyz.shape[0], C, device="cuda")

    def run(layer):
        layer.zero_grad(set_to_none=True)
        f = feats.clone().requires_grad_(True)
        out = layer(f, xyz, offset)
        (out[0].square().sum() + out[3].square().sum()).backward()
        return out, dict(gf=f.grad, **{n: p.grad for n, p in layer.named_parameters()})

    (wf, wx, wo, wfd, wxd, wod), wg = run(theirs)
    (gf, gx, go, gfd, gxd, god), gg = run(mine)
    assert torch.equal(gx, wx) and torch.equal(go.long(), wo.long())
    assert torch.equal(gxd, wxd) and torch.equal(god.long().cpu(), wod.long().cpu())      # FPS + offsets of TransitionDown: exact
    _close(gf, wf, "feats")
    _close(gfd, wfd, "feats_down")
    assert set(gg) == set(wg)
    for name in wg:
        _close(gg[name], wg[name], f"grad {name}", tol=5e-4)


def test_reference_upsample_and_transition_down_match_mirror():
    from stratified_transformer_b200 import layers
    ref = _ref()
    xyz, offset = _scene(3, 1500, 9)
    torch.manual_seed(2)
    td_ref = ref.TransitionDown(48, 96, 0.25, 16).cuda()
    td = layers.TransitionDown(48, 96, 0.25, 16).cuda()
    td.load_state_dict(td_ref.state_dict())
    feats = torch.randn(xyz.shape[0], 48, device="cuda")
    wf, wx, wo = td_ref(feats, xyz, offset)
    gf, gx, go = td(feats, xyz, offset)
    assert torch.equal(gx, wx) and torch.equal(go.cpu().long(), wo.cpu().long())
    _close(gf, wf, "TransitionDown feats", tol=1e-5)
    up_ref = ref.Upsample(3, 96, 48).cuda()
    up = layers.Upsample(3, 96, 48).cuda()
    up.load_state_dict(up_ref.state_dict())
    w = up_ref(wf, wx, xyz, wo, offset, support_feats=feats)
    g = up(gf, gx, xyz, go, offset, support_feats=feats)
    _close(g[0], w[0], "Upsample feats", tol=1e-5)
    assert torch.equal(g[1], w[1]) and torch.equal(g[2], w[2])


def test_transition_down_offsets_carry_fractions_like_the_reference():
    """stratified_transformer.py:98-102: only the first scene's count is truncated before the running sum."""
    from stratified_transformer_b200.layers import transition_down_offsets
    off = [1001, 2003, 3006, 4010]
    count = int(off[0] * 0.25) + 1
    want = [count]
    for i in range(1, len(off)):
        count += ((off[i] - off[i - 1]) * 0.25) + 1
        want.append(count)
    assert transition_down_offsets(torch.tensor(off), 0.25) == [int(c) for c in want] == [251, 502, 754, 1006]


def test_native_stand_ins_match_the_oracle():
    """the torch voxel_grid / scatter_softmax stand-ins of the native reference module against the oracle's restatements"""
    from oracle import attention_oracle as ao, index_oracle as io
    nat = _native()
    xyz, offset = _scene(2, 4000, 3)
    batch = torch.from_numpy(io.batch_from_offset(offset.cpu().numpy())).cuda()
    for w in (0.16, 0.32):
        size = torch.tensor([w] * 3, device="cuda")
        for pos, start in ((xyz, None), (xyz + 0.5 * size, xyz.min(0)[0])):
            got = nat.voxel_grid(pos, batch, size, start)
            want = io.voxel_grid(pos.cpu().numpy(), batch.cpu().numpy(), size.cpu().numpy(), None if start is None else start.cpu().numpy())
            assert np.array_equal(got.cpu().numpy(), want)
    i0 = torch.sort(torch.randint(0, 300, (5000,), device="cuda"))[0]
    s = torch.randn(5000, 3, device="cuda") * 3
    _close(nat.scatter_softmax(s, i0, dim=0), ao.softmax_fwd(s.double().cpu(), i0.cpu(), 300), "scatter_softmax stand-in", tol=1e-6)


def test_native_reference_layer_matches_mirror():
    """The mirror layer (device builder + libstb200) against the reference layer running entirely on its OWN code and kernels."""
    from stratified_transformer_b200 import layers
    nat = _native()
    xyz, offset = _scene(2, 2500, 7)
    C, h, window, quant, ds, depth = 48, 3, 0.32, 0.02, 8, 2
    torch.manual_seed(3)
    kw = dict(rel_query=True, rel_key=True, rel_value=True, drop_path=0.0, ratio=0.25, k=16, out_channels=96)
    theirs = nat.BasicLayer(ds, depth, C, h, window, 0.04, quant, downsample=nat.TransitionDown, **kw).cuda()
    for name, p in theirs.named_parameters():
        if "relative_pos" in name:
            torch.nn.init.uniform_(p, -0.3, 0.3)
    mine = layers.BasicLayer(ds, depth, C, h, window, 0.04, quant, downsample=layers.TransitionDown, **kw).cuda()
    mine.load_state_dict(theirs.state_dict())
    feats = torch.randn(xyz.shape[0], C, device="cuda")

    def run(layer):
        layer.zero_grad(set_to_none=True)
        f = feats.clone().requires_grad_(True)
        out = layer(f, xyz, offset)
        (out[0].square().sum() + out[3].square().sum()).backward()
        return out, dict(gf=f.grad, **{n: p.grad for n, p in layer.named_parameters()})

    (wf, _, _, wfd, wxd, wod), wg = run(theirs)
    (gf, _, _, gfd, gxd, god), gg = run(mine)
    assert torch.equal(gxd, wxd) and torch.equal(god.long().cpu(), wod.long().cpu())
    _close(gf, wf, "feats")
    _close(gfd, wfd, "feats_down")
    for name in wg:
        _close(gg[name], wg[name], f"grad {name}", tol=5e-4)
