"""Host pre-step (/root/reference/train.py:319-325): oracle self-checks on the CPU, GPU == oracle exactly."""
import numpy as np
import pytest
import torch

from oracle import prestep_oracle as po


def _scenes(sizes, seed, lattice=False):
    rng = np.random.default_rng(seed)
    pts = []
    for n in sizes:
        if lattice:   # voxel-centre lattice: many exactly equal distances (ties broken by index)
            g = rng.integers(0, 12, (n, 3)).astype(np.float32) * np.float32(0.04)
            pts.append(g)
        else:
            pts.append((rng.random((n, 3)) * [2.0, 1.5, 1.0]).astype(np.float32))
    return np.concatenate(pts), np.cumsum(sizes).astype(np.int32)


def test_batch_vector_matches_the_reference_statements():
    offset = torch.tensor([5, 5, 12, 40], dtype=torch.int32)     # includes an empty scene
    counts = torch.diff(offset, prepend=offset.new_zeros(1))
    want = torch.repeat_interleave(torch.arange(len(offset)), counts.long())    # == the reference's list concatenation
    assert np.array_equal(po.batch_vector(offset.numpy()), want.numpy())


def test_ball_query_oracle_properties():
    xyz, off = _scenes([300, 200], 1)
    batch = po.batch_vector(off)
    idx, d2 = po.ball_query_partial_dense(0.2, 16, xyz, xyz, batch, batch)
    assert (idx[:, 0] == np.arange(len(xyz))).all() and (d2[:, 0] == 0).all()     # a point finds itself first
    valid = idx >= 0
    assert (batch[idx[valid]] == np.repeat(batch, 16).reshape(-1, 16)[valid]).all()      # never across scenes
    assert (d2[valid] < np.float32(0.2) ** 2).all() and (d2[~valid] == -1).all()
    dd = np.where(valid, d2, np.float32(1e30))
    assert (dd[:, 1:] >= dd[:, :-1]).all()                                            # ascending distances, padding last
    # complete when fewer than max_num are in range
    i = int(np.argmin(valid.sum(1)))
    full = np.nonzero((((xyz - xyz[i]) ** 2).sum(1) < 0.04) & (batch == batch[i]))[0]
    if valid[i].sum() < 16:
        assert set(idx[i][valid[i]]) == set(full)


@pytest.mark.gpu
def test_batch_from_offset_gpu():
    from stratified_transformer_b200 import prestep
    for off in ([7], [5, 5, 12, 40], list(np.cumsum(np.random.default_rng(0).integers(1, 3000, 37)))):
        got = prestep.batch_from_offset(torch.tensor(off, dtype=torch.int64).cuda())
        assert got.dtype == torch.int64
        assert np.array_equal(got.cpu().numpy(), po.batch_vector(np.array(off)))


@pytest.mark.gpu
@pytest.mark.parametrize("sizes,radius,k,lattice", [([1500, 900, 1200], 0.1, 34, False), ([2000], 0.25, 34, False),
                                                    ([1800, 1700], 0.1, 34, True), ([700, 1, 400], 0.3, 8, False),
                                                    ([1000, 1000], 0.1, 64, True)])
def test_ball_query_gpu_matches_oracle(sizes, radius, k, lattice):
    from stratified_transformer_b200 import prestep
    xyz, off = _scenes(sizes, len(sizes) + k, lattice)
    batch = po.batch_vector(off)
    want_idx, want_d2 = po.ball_query_partial_dense(radius, k, xyz, xyz, batch, batch)
    x = torch.from_numpy(xyz).cuda()
    b = prestep.batch_from_offset(torch.from_numpy(off).cuda())
    idx, d2 = prestep.ball_query(radius, k, x, x, mode="partial_dense", batch_x=b, batch_y=b)
    assert np.array_equal(idx.cpu().numpy(), want_idx)
    assert np.array_equal(d2.cpu().numpy(), want_d2)


@pytest.mark.gpu
def test_ball_query_gpu_separate_queries_and_no_batch():
    from stratified_transformer_b200 import prestep
    rng = np.random.default_rng(5)
    x = (rng.random((3000, 3)) * 2).astype(np.float32)
    y = (rng.random((500, 3)) * 2.4 - 0.2).astype(np.float32)        # some queries outside the support box
    want_idx, want_d2 = po.ball_query_partial_dense(0.15, 20, x, y)
    idx, d2 = prestep.ball_query(0.15, 20, torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda())
    assert np.array_equal(idx.cpu().numpy(), want_idx) and np.array_equal(d2.cpu().numpy(), want_d2)
    with pytest.raises(ValueError):
        prestep.ball_query(0.1, 8, torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), mode="dense")
