"""Synthetic S3DIS / ScanNet-shape scenes (there is no dataset access in this environment).

Mirrors what the reference's data pipeline hands to the model — util/data_util.py:179-202
(`data_prepare_v101`: voxelize, crop `voxel_max` nearest points, shuffle, subtract min) and
util/voxelize.py:80-93 (one *raw* point kept per occupied voxel, coordinates stay continuous) —
on a procedurally generated room: 6 bounding planes plus axis-aligned clutter boxes standing
on the floor, sampled uniformly by area with 5 mm noise.  Point order is random, as in the
reference (util/data_util.py:192-195): that is what makes the k/v gathers of the per-op path
spatially incoherent.
"""
from __future__ import annotations

import numpy as np


def _box_faces(lo, hi):
    """6 faces of an axis-aligned box as (origin, edge_u, edge_v)."""
    lo, hi = np.asarray(lo, np.float64), np.asarray(hi, np.float64)
    ext = hi - lo
    faces = []
    for ax in range(3):
        u, v = (ax + 1) % 3, (ax + 2) % 3
        eu, ev = np.zeros(3), np.zeros(3)
        eu[u], ev[v] = ext[u], ext[v]
        for side in (lo, hi):
            o = lo.copy()
            o[ax] = side[ax]
            faces.append((o, eu, ev))
    return faces


def make_scene(seed: int, n_points: int = 80000, voxel: float = 0.04, n_raw: int = 1_500_000,
               n_boxes: int = 25, lattice: bool = False):
    """Return (xyz float32 [n_points,3], rgb float32 [n_points,3]).

    lattice=True snaps coordinates to the voxel lattice: an adversarial case for the two
    window-id roundings of the reference, used by exactness tests only.
    """
    rng = np.random.default_rng(1000 + seed)
    room = np.array([rng.uniform(6, 10), rng.uniform(5, 8), rng.uniform(2.8, 3.2)])
    faces = _box_faces(np.zeros(3), room)
    for _ in range(n_boxes):
        e = rng.uniform(0.3, 1.5, 3)
        lo = np.array([rng.uniform(0, room[0] - e[0]), rng.uniform(0, room[1] - e[1]), 0.0])
        faces += _box_faces(lo, lo + e)
    area = np.array([np.linalg.norm(np.cross(eu, ev)) for _, eu, ev in faces])
    which = rng.choice(len(faces), size=n_raw, p=area / area.sum())
    o = np.stack([f[0] for f in faces])[which]
    eu = np.stack([f[1] for f in faces])[which]
    ev = np.stack([f[2] for f in faces])[which]
    pts = o + rng.random((n_raw, 1)) * eu + rng.random((n_raw, 1)) * ev
    pts += rng.normal(0.0, 0.005, pts.shape)
    # voxelize (mode 0): one raw point per occupied voxel
    key = np.floor((pts - pts.min(0)) / voxel).astype(np.int64)
    key = (key[:, 0] * 1_000_003 + key[:, 1]) * 1_000_003 + key[:, 2]
    pick = rng.permutation(n_raw)
    _, first = np.unique(key[pick], return_index=True)
    pts = pts[pick[first]]
    if pts.shape[0] < n_points:
        raise ValueError(f"scene {seed}: only {pts.shape[0]} occupied voxels, need {n_points}")
    centre = pts[rng.integers(pts.shape[0])]
    near = np.argsort(((pts - centre) ** 2).sum(1))[:n_points]
    pts = pts[near][rng.permutation(n_points)]
    pts = pts - pts.min(0)
    if lattice:
        pts = np.round(pts / voxel) * voxel
    return pts.astype(np.float32), rng.random((n_points, 3), dtype=np.float32)


def make_batch(n_scenes: int, n_points: int = 80000, voxel: float = 0.04, seed0: int = 0, **kw):
    """-> xyz [N,3] f32, rgb [N,3] f32, offset int32 [b] cumulative (the collate_fn output shape,
    util/data_util.py:61-79)."""
    xs, cs = zip(*(make_scene(seed0 + s, n_points, voxel, **kw) for s in range(n_scenes)))
    offset = np.cumsum([x.shape[0] for x in xs]).astype(np.int32)
    return np.concatenate(xs), np.concatenate(cs), offset
