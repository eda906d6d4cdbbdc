"""Oracle (index construction) against the golden vectors produced by the reference's own Python
(tests/golden/make_golden.py) and against CPU torch's `//`, `%`, `round`."""
import os

import numpy as np
import pytest
import torch

from oracle import fps_oracle, index_oracle as io

CASES = ["s3dis_small", "s3dis_lattice", "scannet_small"]


def test_floor_div_matches_torch():
    rng = np.random.default_rng(0)
    a = np.concatenate([rng.uniform(-1, 12, 2_000_000), np.round(rng.uniform(-1, 12, 200_000) / 0.04) * 0.04,
                        [0.0, -0.0, 0.16, 0.32, -0.16]]).astype(np.float32)
    for b in (0.16, 0.32, 0.1, 0.01, 0.02, 0.08, 1.28):
        want = (torch.from_numpy(a) // b).numpy()
        got = io.floor_div_f32(a, np.float32(b))
        assert np.array_equal(want.view(np.int32), got.view(np.int32)), b


def test_remainder_matches_torch():
    rng = np.random.default_rng(1)
    a = rng.uniform(-3, 12, 1_000_000).astype(np.float32)
    for b in (0.16, 0.32, 0.1):
        want = (torch.from_numpy(a) % b).numpy()
        assert np.array_equal(want.view(np.int32), io.remainder_f32(a, np.float32(b)).view(np.int32))


@pytest.mark.parametrize("name", CASES)
def test_index_matches_reference_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, f"index_{name}.npz"))
    xyz, offset = g["xyz"], g["offset"]
    w, quant, ds = float(g["window_size"]), float(g["quant_size"]), int(g["downsample_scale"])
    new_offset = io.fps_new_offset(offset, ds)
    assert np.array_equal(new_offset, g["new_offset"])
    ds_idx = fps_oracle.furthestsampling(xyz, offset, new_offset)
    assert np.array_equal(ds_idx, g["downsample_idx"])
    for parity in (0, 1):
        r = io.build_layer_index(xyz, offset, w, ds, ds_idx, parity)
        assert np.array_equal(r["offsets"], g[f"p{parity}_offsets"])
        assert r["n_max"] == int(g[f"p{parity}_n_max"])
        i1c = io.canonicalize(r["offsets"], r["index_1"])
        assert np.array_equal(i1c, g[f"p{parity}_index_1"])
        # the goldens were produced by CPU torch (this container has no GPU): the oracle's "cpu" arithmetic
        rel = io.rel_pos_index_stratified(xyz, r["index_0"], i1c, w, quant, device="cpu")
        assert np.array_equal(rel, g[f"p{parity}_rel_idx"].astype(np.int32))


def test_lattice_case_has_duplicate_keys(golden_dir):
    """The reference emits the same (query, key) twice when its two window-id roundings disagree
    (SURVEY B.4); the lattice fixture must exercise that, kernels must tolerate it."""
    g = np.load(os.path.join(golden_dir, "index_s3dis_lattice.npz"))
    dup = 0
    for parity in (0, 1):
        off, i1 = g[f"p{parity}_offsets"].astype(np.int64), g[f"p{parity}_index_1"].astype(np.int64)
        seg = np.repeat(np.arange(off.shape[0] - 1), np.diff(off))
        key = seg * (off.shape[0]) + i1
        dup += key.shape[0] - np.unique(key).shape[0]
    assert dup > 0


def test_swin_rel_idx_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "relidx_swin.npz"))
    for shift, tag in ((0.0, "noshift"), (0.08, "shift")):
        got = io.rel_pos_index_swin(g["xyz"], g["index_0"], g["index_1"], 0.16, 0.01, shift)
        assert np.array_equal(got, g[f"rel_idx_{tag}"].astype(np.int32))


def test_rel_idx_matches_torch_random():
    rng = np.random.default_rng(5)
    xyz = rng.uniform(0, 8, (50_000, 3)).astype(np.float32)
    i0 = rng.integers(0, 50_000, 400_000)
    i1 = np.clip(i0 + rng.integers(-3, 4, 400_000), 0, 49_999)
    xyz[i1] = xyz[i0] + rng.uniform(-0.3199, 0.3199, (400_000, 3)).astype(np.float32)
    x = torch.from_numpy(xyz)
    rel = x[torch.from_numpy(i0)] - x[torch.from_numpy(i1)]
    rel = torch.round(rel * 100000) / 100000
    want = ((rel + 2 * 0.16 - 0.0001) // 0.01).int().numpy()
    assert np.array_equal(want, io.rel_pos_index_stratified(xyz, i0, i1, 0.16, 0.01, device="cpu"))
    # the CUDA form (reciprocal multiply) differs in a small fraction of the pairs and only by one bin
    cuda = io.rel_pos_index_stratified(xyz, i0, i1, 0.16, 0.01, device="cuda")
    diff = cuda != want
    assert 0 < diff.mean() < 1e-3 and np.abs(cuda - want).max() == 1


def test_fps_oracle_properties():
    rng = np.random.default_rng(2)
    xyz = rng.uniform(0, 5, (3000, 3)).astype(np.float32)
    offset = np.array([1000, 1700, 3000], np.int32)
    new_offset = io.fps_new_offset(offset, 8)
    idx = fps_oracle.furthestsampling(xyz, offset, new_offset)
    starts = np.concatenate([[0], offset[:-1]])
    nstarts = np.concatenate([[0], new_offset[:-1]])
    for s in range(3):
        sel = idx[nstarts[s]:new_offset[s]]
        assert sel[0] == starts[s]
        assert ((sel >= starts[s]) & (sel < offset[s])).all()
        assert np.unique(sel).shape[0] == sel.shape[0]
        # greedy property: every pick maximises the distance to the already picked set
        pts = xyz[starts[s]:offset[s]].astype(np.float64)
        mind = np.full(pts.shape[0], np.inf)
        for j in range(1, min(40, sel.shape[0])):
            mind = np.minimum(mind, ((pts - pts[sel[j - 1] - starts[s]]) ** 2).sum(1))
            assert mind[sel[j] - starts[s]] >= mind.max() * (1 - 1e-5)


def test_fps_block_size():
    assert fps_oracle.block_size(80000) == 1024
    assert fps_oracle.block_size(1000) == 512
    assert fps_oracle.block_size(1024) == 1024
    assert fps_oracle.block_size(5) == 4


@pytest.mark.parametrize("seed,lattice", [(0, False), (1, True), (2, False)])
def test_index_oracle_equals_definition_brute_force(seed, lattice):
    """Independent of the reference's tensor code: per query, keys = points of its small window (dense), then FPS-sampled
    points of its large window whose floor-div window coordinate differs (sparse) — straight from SURVEY B.4."""
    rng = np.random.default_rng(seed)
    sizes = [230, 170]
    xyz = rng.uniform(0, 1.3, (sum(sizes), 3))
    if lattice:
        xyz = np.round(xyz / 0.04) * 0.04
    xyz = xyz.astype(np.float32)
    offset = np.cumsum(sizes).astype(np.int32)
    w, ds = 0.16, 4
    batch = io.batch_from_offset(offset)
    ds_idx = fps_oracle.furthestsampling(xyz, offset, io.fps_new_offset(offset, ds))
    sampled = np.zeros(xyz.shape[0], bool)
    sampled[ds_idx] = True
    w3 = np.full(3, w, np.float32)
    for parity in (0, 1):
        got = io.build_layer_index(xyz, offset, w, ds, ds_idx, parity)
        if parity == 0:
            small = io.voxel_grid(xyz, batch, w3, None)
            large = io.voxel_grid(xyz, batch, (np.float32(2) * w3).astype(np.float32), None)
            wc = io.floor_div_f32((xyz - xyz.min(0)).astype(np.float32), w3)
        else:
            mn = xyz.min(0)
            half, full = (np.float32(0.5) * w3).astype(np.float32), w3
            small = io.voxel_grid((xyz + half).astype(np.float32), batch, w3, mn)
            large = io.voxel_grid((xyz + full).astype(np.float32), batch, (np.float32(2) * w3).astype(np.float32), mn)
            wc = io.floor_div_f32(((xyz + half).astype(np.float32) - mn).astype(np.float32), w3)
        keys = []
        for a in range(xyz.shape[0]):
            dense = np.nonzero(small == small[a])[0]
            sparse = np.nonzero((large == large[a]) & sampled & (wc != wc[a]).any(1))[0]
            keys.append(np.concatenate([dense, sparse]))
        want_off = np.concatenate([[0], np.cumsum([len(k) for k in keys])])
        assert np.array_equal(got["offsets"], want_off)
        assert np.array_equal(got["index_1"], np.concatenate(keys))
