"""kNN / grouping / interpolation (-m gpu): bit-exact indices and squared distances against the C
oracle (oracle/fps_oracle.c: knn_oracle) and the reference kernel itself, including tie-saturated lattice scenes."""
import numpy as np
import pytest
import torch

from oracle import fps_oracle, ref_cuda

pytestmark = pytest.mark.gpu


def pts(n, seed, lattice=None):
    rng = np.random.default_rng(seed)
    p = rng.uniform(0, 4, (n, 3))
    if lattice:
        p = np.round(p / lattice) * lattice
    return p.astype(np.float32)


@pytest.mark.parametrize("sizes,frac,k,lattice", [
    ([900], 1.0, 16, None), ([700, 1300, 40], 0.25, 16, None), ([2000, 1500], 0.3, 3, None),
    ([600, 800], 0.5, 16, 0.25), ([5, 9], 1.0, 16, None), ([3000], 0.1, 100, 0.5),
])
def test_knn_exact(sizes, frac, k, lattice):
    from stratified_transformer_b200 import pointops
    xyz = np.concatenate([pts(n, 20 + i, lattice) for i, n in enumerate(sizes)])
    offset = np.cumsum(sizes).astype(np.int32)
    starts = np.concatenate([[0], offset[:-1]])
    rng = np.random.default_rng(0)
    new_parts, new_sizes = [], []
    for s, e in zip(starts, offset):
        cnt = max(1, int((e - s) * frac))
        sel = np.sort(rng.choice(np.arange(s, e), cnt, replace=False))
        new_parts.append(xyz[sel]); new_sizes.append(cnt)
    new_xyz = np.concatenate(new_parts)
    new_offset = np.cumsum(new_sizes).astype(np.int32)
    want_idx, want_d2 = fps_oracle.knnquery(k, xyz, new_xyz, offset, new_offset)
    xd, nd = torch.from_numpy(xyz).cuda(), torch.from_numpy(new_xyz).cuda()
    od, nod = torch.from_numpy(offset).cuda(), torch.from_numpy(new_offset).cuda()
    idx, dist = pointops.knnquery(k, xd, nd, od, nod)
    assert np.array_equal(idx.cpu().numpy(), want_idx)
    assert np.array_equal(dist.cpu().numpy(), np.sqrt(want_d2))
    if ref_cuda.available():
        ridx, rd2 = ref_cuda.knnquery(k, xd, nd, od, nod)
        assert np.array_equal(ridx.cpu().numpy(), want_idx), "oracle disagrees with the reference kernel"
        assert np.array_equal(rd2.cpu().numpy(), want_d2)


def test_queryandgroup_and_interpolation():
    from stratified_transformer_b200 import pointops
    xyz = pts(1500, 3); sup = xyz[::4].copy()
    off = torch.tensor([1500], dtype=torch.int32).cuda(); soff = torch.tensor([sup.shape[0]], dtype=torch.int32).cuda()
    xd, sd = torch.from_numpy(xyz).cuda(), torch.from_numpy(sup).cuda()
    feat = torch.randn(1500, 8).cuda()
    g = pointops.queryandgroup(16, xd, sd, feat, None, off, soff, use_xyz=True)
    idx, _ = fps_oracle.knnquery(16, xyz, sup, off.cpu().numpy(), soff.cpu().numpy())
    want = np.concatenate([xyz[idx] - sup[:, None, :], feat.cpu().numpy()[idx]], -1)
    assert np.allclose(g.cpu().numpy(), want, atol=1e-6)
    sfeat = torch.randn(sup.shape[0], 8).cuda()
    out = pointops.interpolation(sd, xd, sfeat, soff, off)          # support -> dense (Upsample)
    idx3, d2 = fps_oracle.knnquery(3, sup, xyz, soff.cpu().numpy(), off.cpu().numpy())
    wgt = 1.0 / (np.sqrt(d2) + 1e-8); wgt /= wgt.sum(1, keepdims=True)
    want = (sfeat.cpu().numpy()[idx3] * wgt[..., None]).sum(1)
    assert np.allclose(out.cpu().numpy(), want, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("n_scenes,n_pts,k,direction", [(2, 30000, 16, "down"), (1, 80000, 16, "down"), (2, 30000, 3, "up")])
def test_knn_grid_search_matches_reference_kernel_on_scene_geometry(n_scenes, n_pts, k, direction):
    """Room-like synthetic scenes (surfaces, empty space): the grid-pruned search + heap completion against the reference's own
    kernel, both directions the model uses (TransitionDown: queries are a subset of the support; Upsample: the support is the
    subset and most queries are not in it)."""
    if not ref_cuda.available():
        pytest.skip("oracle/_ref not built")
    from stratified_transformer_b200 import pointops
    from stratified_transformer_b200.synthetic import make_batch
    xyz, _, offset = make_batch(n_scenes, n_pts, seed0=11)
    starts = np.concatenate([[0], offset[:-1]])
    sel = np.concatenate([np.arange(s, e)[::4] for s, e in zip(starts, offset)])
    sub = xyz[sel].copy()
    sub_off = np.cumsum([len(np.arange(s, e)[::4]) for s, e in zip(starts, offset)]).astype(np.int32)
    if direction == "down":
        sup, sup_off, qry, qry_off = xyz, offset, sub, sub_off
    else:
        sup, sup_off, qry, qry_off = sub, sub_off, xyz, offset
    sd, qd = torch.from_numpy(sup).cuda(), torch.from_numpy(qry).cuda()
    so, qo = torch.from_numpy(sup_off).cuda(), torch.from_numpy(qry_off).cuda()
    idx, dist = pointops.knnquery(k, sd, qd, so, qo)
    ridx, rd2 = ref_cuda.knnquery(k, sd, qd, so, qo)
    assert torch.equal(idx, ridx)
    assert torch.equal(dist, torch.sqrt(rd2))
