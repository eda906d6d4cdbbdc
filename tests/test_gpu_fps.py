"""Exact FPS (-m gpu): the cluster kernel must return bit-identical indices to the C oracle
(oracle/fps_oracle.c) and to the reference kernel itself (oracle/_ref), including tie-heavy
lattice-aligned scenes and every code path (single CTA, cluster, several points per thread)."""
import numpy as np
import pytest
import torch

from oracle import fps_oracle, index_oracle as io, ref_cuda

pytestmark = pytest.mark.gpu


def ours(xyz, offset, new_offset):
    from stratified_transformer_b200 import pointops
    idx = pointops.furthestsampling(torch.from_numpy(xyz).cuda(), torch.from_numpy(offset).cuda(),
                                    torch.from_numpy(new_offset).cuda())
    torch.cuda.synchronize()
    return idx.cpu().numpy()


def scene(n, seed, lattice=False, extent=(6.0, 5.0, 3.0)):
    rng = np.random.default_rng(seed)
    pts = rng.uniform(0, 1, (n, 3)) * np.asarray(extent)
    if lattice:
        pts = np.round(pts / 0.04) * 0.04
    return pts.astype(np.float32)


@pytest.mark.parametrize("sizes,ds,lattice", [
    ([700], 8, False),                 # B = 512 < 1024, single CTA
    ([1000, 333, 1024], 4, False),     # ragged batch, mixed sizes
    ([5000, 4100], 8, False),          # cluster path
    ([3000, 2500], 8, True),           # exact distance ties everywhere
    ([20001, 19999, 20000], 8, False), # S3DIS layer-1 sizes
    ([1], 8, False), ([2, 5], 1, False),
])
def test_fps_exact_vs_oracle_and_reference(sizes, ds, lattice):
    xyz = np.concatenate([scene(n, 10 + i, lattice) for i, n in enumerate(sizes)])
    offset = np.cumsum(sizes).astype(np.int32)
    new_offset = io.fps_new_offset(offset, ds)
    want = fps_oracle.furthestsampling(xyz, offset, new_offset)
    got = ours(xyz, offset, new_offset)
    assert np.array_equal(got, want)
    if ref_cuda.available():
        ref = ref_cuda.furthestsampling(torch.from_numpy(xyz).cuda(), torch.from_numpy(offset).cuda(),
                                        torch.from_numpy(new_offset).cuda()).cpu().numpy()
        assert np.array_equal(ref, want), "oracle disagrees with the reference kernel"


@pytest.mark.parametrize("n,spacing", [(700, 0.5), (1251, 0.25), (4000, 0.25), (5001, 0.5), (20001, 0.25),
                                       (52000, 0.125)])   # > 49152 points: 128-thread CTAs, 40 points per thread, packed pairs
def test_fps_exact_with_massive_ties(n, spacing):
    """Coarse lattice: most distances tie exactly, so every level of the (distance, rank) reduction and the in-thread
    scan order are exercised for each launch configuration (single CTA, small / large clusters)."""
    rng = np.random.default_rng(n)
    xyz = (np.round(rng.uniform(0, 6, (n, 3)) / spacing) * spacing).astype(np.float32)
    offset = np.array([n], np.int32)
    new_offset = io.fps_new_offset(offset, 8)
    want = fps_oracle.furthestsampling(xyz, offset, new_offset)
    assert np.array_equal(ours(xyz, offset, new_offset), want)
    two = np.concatenate([xyz, xyz[::-1].copy()])
    offset2 = np.array([n, 2 * n], np.int32)
    new_offset2 = io.fps_new_offset(offset2, 4)
    assert np.array_equal(ours(two, offset2, new_offset2), fps_oracle.furthestsampling(two, offset2, new_offset2))


def test_fps_full_size_scene():
    """BASELINE cfg2 scene size: 80k points, 10 001 samples; oracle takes ~2 s."""
    from stratified_transformer_b200.synthetic import make_scene
    xyz, _ = make_scene(0, 80000)
    offset = np.array([80000], np.int32)
    new_offset = io.fps_new_offset(offset, 8)
    want = fps_oracle.furthestsampling(xyz, offset, new_offset)
    got = ours(xyz, offset, new_offset)
    assert np.array_equal(got, want)


def test_fps_streaming_fallback_for_very_large_scene():
    """A 200k-point scene exceeds the register-resident cluster kernel (16 CTAs x 256 threads x 40 points): the
    streaming kernel takes over (scratch allocated by the wrapper) and must give the same exact result."""
    rng = np.random.default_rng(5)
    n = 200_000
    xyz = rng.uniform(0, 10, (n, 3)).astype(np.float32)
    offset = np.array([n], np.int32)
    new_offset = io.fps_new_offset(offset, 64)
    want = fps_oracle.furthestsampling(xyz, offset, new_offset)
    assert np.array_equal(ours(xyz, offset, new_offset), want)


@pytest.mark.parametrize("b,n_big", [(20, 30000), (38, 9000)])
def test_fps_large_batches_fall_back_to_streaming(b, n_big):
    """Batches whose cluster configuration exceeds the register kernel's budget (b = 20 scenes with a 30k-point one: cluster 2
    x 512 threads x 30 points per thread; b >= 38: one CTA per scene, 1024 threads, more than 8 points per thread) used to
    fail with STB200_ERR_ARG; they must take the streaming kernel and stay exact."""
    sizes = [n_big] + [600 + 37 * i for i in range(b - 1)]
    xyz = np.concatenate([scene(n, 100 + i) for i, n in enumerate(sizes)])
    offset = np.cumsum(sizes).astype(np.int32)
    new_offset = io.fps_new_offset(offset, 64)
    want = fps_oracle.furthestsampling(xyz, offset, new_offset)
    assert np.array_equal(ours(xyz, offset, new_offset), want)


def test_shorter_fps_run_is_a_prefix_of_the_longer_one():
    """index.fps_prefix: the picks for n // 8 + 1 samples are the first picks of the run for n // 4 + 1 samples, per scene
    (what layers.BasicLayer uses to run FPS once per layer instead of twice)."""
    from stratified_transformer_b200 import index, pointops
    from stratified_transformer_b200.synthetic import make_batch
    xyz, _, offset = make_batch(3, 6000, seed0=2, n_raw=100000)
    xd, od = torch.from_numpy(xyz).cuda(), torch.from_numpy(offset).cuda()
    long_off = index.fps_new_offset(od, 4)
    short_off = index.fps_new_offset(od, 8)
    long_idx = pointops.furthestsampling(xd, od, long_off)
    want = pointops.furthestsampling(xd, od, short_off)
    assert torch.equal(index.fps_prefix(long_idx, long_off, short_off), want)


def test_pruned_variant_is_bit_identical(monkeypatch):
    """STB200_FPS_PRUNE=1: exact bounding-box pruning of the distance update (csrc/fps.cu, fps_pruned_kernel) - same picks as
    the plain kernel on room-like scenes, a tie-saturated lattice and uniform noise, 1 and 4 scenes."""
    from stratified_transformer_b200 import pointops
    from stratified_transformer_b200.synthetic import make_batch
    rng = np.random.default_rng(0)
    cases = [make_batch(1, 80000, seed0=21)[::2], make_batch(4, 50000, seed0=22)[::2],
             ((rng.integers(0, 40, (60000, 3)) * 0.05).astype(np.float32), np.array([60000], np.int32)),
             (rng.random((70000, 3)).astype(np.float32), np.array([30000, 70000], np.int32))]
    for xyz, offset in cases:
        xd, od = torch.from_numpy(np.ascontiguousarray(xyz)).cuda(), torch.from_numpy(offset).cuda()
        new_off = torch.cumsum(torch.diff(od, prepend=od.new_zeros(1)) // 8 + 1, 0).int()
        monkeypatch.setenv("STB200_FPS_PRUNE", "0")
        want = pointops.furthestsampling(xd, od, new_off)
        monkeypatch.setenv("STB200_FPS_PRUNE", "1")
        got = pointops.furthestsampling(xd, od, new_off)
        assert torch.equal(got, want)
