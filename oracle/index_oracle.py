"""CPU oracle: pair-index construction of the window-attention hot path (TEST INFRASTRUCTURE).

NumPy restatement, fp32 arithmetic reproduced operation by operation.  Reference it follows
(paths under /root/reference):
  * voxel_grid          THIRD PARTY, not vendored: torch_geometric==1.7.0 (requirements.txt:13)
                        `voxel_grid` -> torch_cluster `grid_cluster` (version not pinned anywhere).
                        Restated from the published algorithm; call sites
                        model/stratified_transformer.py:50, model/swin3d_transformer.py:17.
                        PARITY UNPINNED for this one function (no reference test touches it).
  * grid_sample         model/stratified_transformer.py:44-65
  * get_indice_pairs    model/stratified_transformer.py:10-42
  * sort / CSR          model/stratified_transformer.py:311-317
  * new_offset          model/stratified_transformer.py:282-288
  * rel-pos index       model/stratified_transformer.py:186-188 (Stratified),
                        model/swin3d_transformer.py:129-130,151-154 (Swin)
  * torch `//`, `%`     c10/util/generic_math.h div_floor_floating (fmod based)

Pinning: grid_sample / get_indice_pairs / sort are pinned against the reference's own Python
source executed on CPU (tests/golden/make_golden.py extracts the function text from
/root/reference at generation time, strips `.cuda()`, supplies `voxel_grid` below) -> fixtures
tests/golden/index_*.npz.  floor_div_f32 / rel_pos_index are pinned against CPU torch's own
`//`, `round`, `%` in tests/test_oracle_index.py.

Canonical order: the reference sorts with unstable sorts, so within a
query segment its key order is unspecified.  The oracle (and the CUDA builder) emit, per query:
dense keys ascending by point id, then sparse keys ascending by point id.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


# ----------------------------------------------------------------------------- fp32 helpers
def floor_div_f32(a, b):
    """torch `a // b` for float32 (div_floor_floating): fmod-based, with the 0.5 fix-up."""
    a = np.asarray(a, dtype=f32)
    b = np.asarray(b, dtype=f32)
    with np.errstate(all="ignore"):
        mod = np.fmod(a, b).astype(f32)
        div = ((a - mod).astype(f32) / b).astype(f32)
        adj = (mod != 0) & ((b < 0) != (mod < 0))
        div = np.where(adj, (div - f32(1)).astype(f32), div)
        fl = np.floor(div).astype(f32)
        fl = np.where((div - fl).astype(f32) > f32(0.5), (fl + f32(1)).astype(f32), fl)
        zero = np.copysign(f32(0), (a / b).astype(f32)).astype(f32)
        out = np.where(div != 0, fl, zero)
        out = np.where(b == 0, (a / b).astype(f32), out)
    return out.astype(f32)


def remainder_f32(a, b):
    """torch `a % b` for float32."""
    a = np.asarray(a, dtype=f32)
    b = np.asarray(b, dtype=f32)
    mod = np.fmod(a, b).astype(f32)
    adj = (mod != 0) & ((b < 0) != (mod < 0))
    return np.where(adj, (mod + b).astype(f32), mod).astype(f32)


# ----------------------------------------------------------------------------- voxel_grid
def voxel_grid(pos, batch, size, start=None):
    """torch_geometric.nn.voxel_grid(pos, batch, size, start=start) -> int64 cluster id per point.

    pos4 = [x,y,z,float(batch)], size4 = [w,w,w,1], start4 = min or [start,0], end4 = max;
    c = sum_d trunc((pos4_d - start4_d) / size4_d) * stride_d, x fastest, batch slowest.
    """
    pos = np.asarray(pos, dtype=f32)
    pos4 = np.concatenate([pos, np.asarray(batch).astype(f32)[:, None]], 1)
    size4 = np.concatenate([np.asarray(size, dtype=f32).reshape(3), np.ones(1, f32)])
    if start is None:
        start4 = pos4.min(0)
    else:
        start4 = np.concatenate([np.asarray(start, dtype=f32).reshape(3), np.zeros(1, f32)])
    end4 = pos4.max(0)
    c = np.zeros(pos.shape[0], np.int64)
    k = np.int64(1)
    for d in range(4):
        c += ((pos4[:, d] - start4[d]).astype(f32) / size4[d]).astype(f32).astype(np.int64) * k
        k = k * (np.int64(((end4[d] - start4[d]).astype(f32) / size4[d]).astype(f32)) + 1)
    return c


def grid_sample(pos, batch, size, start):
    """-> (cluster rank [N] int64, p2v_map [n,k] int64 zero padded, counts [n] int64).
    Row w lists the points of window w ascending by point id (stable argsort = canonical)."""
    cluster = voxel_grid(pos, batch, size, start)
    _, inv, counts = np.unique(cluster, return_inverse=True, return_counts=True)
    inv = inv.reshape(-1).astype(np.int64)
    n, k = counts.shape[0], int(counts.max())
    order = np.argsort(inv, kind="stable")
    p2v = np.zeros((n, k), np.int64)
    col = np.arange(inv.shape[0]) - np.repeat(np.cumsum(counts) - counts, counts)
    p2v[inv[order], col] = order
    return inv, p2v, counts.astype(np.int64)


def batch_from_offset(offset):
    """offset: cumulative int counts [b] -> batch id per point (stratified_transformer.py:273-275)."""
    offset = np.asarray(offset, np.int64)
    counts = np.diff(np.concatenate([[0], offset]))
    return np.repeat(np.arange(offset.shape[0], dtype=np.int64), counts)


def fps_new_offset(offset, downsample_scale):
    """per scene n_i // ds + 1, cumulative (stratified_transformer.py:282-288)."""
    offset = np.asarray(offset, np.int64)
    counts = np.diff(np.concatenate([[0], offset]))
    return np.cumsum(counts // downsample_scale + 1).astype(np.int32)


# ----------------------------------------------------------------------------- pairs
def _window_pairs(p2v, counts, key_mask=None):
    """All (a, b) with a, b in the same row of p2v (row-major); b restricted to key_mask[b]."""
    n, k = p2v.shape
    i0_parts, i1_parts = [], []
    for w in range(n):
        row = p2v[w, : counts[w]]
        keys = row if key_mask is None else row[key_mask[row]]
        if keys.size == 0:
            continue
        i0_parts.append(np.repeat(row, keys.size))
        i1_parts.append(np.tile(keys, row.size))
    if not i0_parts:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    return np.concatenate(i0_parts), np.concatenate(i1_parts)


def get_indice_pairs(p2v, counts, new_p2v, new_counts, downsample_idx, xyz, window_size, parity):
    """Dense pairs inside each small window, then sparse pairs: b sampled, same large window,
    window_coord(a) != window_coord(b) on any axis (window_coord uses torch `//`)."""
    xyz = np.asarray(xyz, f32)
    w = np.asarray(window_size, f32).reshape(3)
    d0, d1 = _window_pairs(p2v, counts)
    ds_mask = np.zeros(xyz.shape[0], bool)
    ds_mask[np.asarray(downsample_idx, np.int64)] = True
    s0, s1 = _window_pairs(new_p2v, new_counts, ds_mask)
    xyz_min = xyz.min(0)
    if parity % 2 == 0:
        wc = floor_div_f32((xyz - xyz_min).astype(f32), w)
    else:
        half = (f32(0.5) * w).astype(f32)
        wc = floor_div_f32(((xyz + half).astype(f32) - xyz_min).astype(f32), w)
    keep = (wc[s0] != wc[s1]).any(-1)
    return np.concatenate([d0, s0[keep]]), np.concatenate([d1, s1[keep]])


def csr_from_pairs(index_0, index_1, N):
    """Stable sort by query id -> (index_0 sorted, index_1, offsets [N+1] int64, n_max)."""
    perm = np.argsort(index_0, kind="stable")
    i0, i1 = index_0[perm], index_1[perm]
    counts = np.bincount(i0, minlength=N)
    offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    return i0, i1, offsets, int(counts.max()) if counts.size else 0


def canonicalize(offsets, index_1):
    """Sort keys inside each segment (used to compare against an unstable-sort producer when the
    dense/sparse split is not available)."""
    offsets = np.asarray(offsets, np.int64)
    seg = np.repeat(np.arange(offsets.shape[0] - 1), np.diff(offsets))
    order = np.lexsort((index_1, seg))
    return np.asarray(index_1)[order]


def build_layer_index(xyz, offset, window_size, downsample_scale, downsample_idx, parity):
    """BasicLayer.forward prologue + one block's index (stratified_transformer.py:267-317) for one
    parity.  `downsample_idx` comes from FPS (fps_oracle).  window_size is a python float."""
    xyz = np.asarray(xyz, f32)
    N = xyz.shape[0]
    batch = batch_from_offset(offset)
    w = np.full(3, window_size, f32)                 # torch.tensor([w]*3).type_as(xyz)
    w2 = (f32(2) * w).astype(f32)                    # 2 * tensor
    if parity % 2 == 0:
        _, p2v, cnt = grid_sample(xyz, batch, w, None)
        _, np2v, ncnt = grid_sample(xyz, batch, w2, None)
    else:
        mn = xyz.min(0)
        _, p2v, cnt = grid_sample((xyz + (f32(0.5) * w).astype(f32)).astype(f32), batch, w, mn)
        _, np2v, ncnt = grid_sample((xyz + (f32(0.5) * w2).astype(f32)).astype(f32), batch, w2, mn)
    i0, i1 = get_indice_pairs(p2v, cnt, np2v, ncnt, downsample_idx, xyz, w, parity)
    i0, i1, offsets, n_max = csr_from_pairs(i0, i1, N)
    return dict(index_0=i0, index_1=i1, offsets=offsets, n_max=n_max)


# ----------------------------------------------------------------------------- rel-pos index
TORCH_DEVICE = "cuda"   # which torch device's arithmetic rel_pos_index_stratified restates by default (see its docstring)


def rel_pos_index_stratified(xyz, index_0, index_1, window_size, quant_size, device=None):
    """idx = ((round((xyz[i0]-xyz[i1])*1e5)/1e5) + 2w - 1e-4) // quant, all fp32, -> int32 [M,3].

    `/ 100000` is the one operation of the path that torch evaluates differently on the two devices: IEEE division on CPU
    tensors, multiplication by the fp32 reciprocal on CUDA tensors (ATen's scalar-divisor fast path; measured on the B200
    box with tools/dbg_torch_cuda_div.py).  device="cuda" (default: the reference only runs on GPUs) / "cpu" (pinned
    bit for bit against CPU torch in tests/test_oracle_index.py and used for the CPU-generated goldens)."""
    device = device or TORCH_DEVICE
    xyz = np.asarray(xyz, f32)
    r = (xyz[index_0] - xyz[index_1]).astype(f32)
    r = np.rint((r * f32(100000)).astype(f32)).astype(f32)
    r = (r * (f32(1) / f32(100000))).astype(f32) if device == "cuda" else (r / f32(100000)).astype(f32)
    t = ((r + f32(2 * window_size)).astype(f32) - f32(0.0001)).astype(f32)
    return floor_div_f32(t, f32(quant_size)).astype(np.int32)


def rel_pos_index_swin(xyz, index_0, index_1, window_size, quant_size, shift_size):
    """xq = ((xyz - min + shift) % w) // quant ; idx = xq[i0] - xq[i1] + int(w/quant) - 1."""
    xyz = np.asarray(xyz, f32)
    qgl = int(window_size / quant_size)
    t = ((xyz - xyz.min(0)).astype(f32) + f32(shift_size)).astype(f32)
    xq = floor_div_f32(remainder_f32(t, f32(window_size)), f32(quant_size))
    rel = (xq[index_0] - xq[index_1]).astype(f32)
    return (rel + f32(qgl - 1)).astype(f32).astype(np.int32)
