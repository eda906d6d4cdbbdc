"""TEST INFRASTRUCTURE (oracle): CPU restatement of the reference's host pre-step, /root/reference/train.py:319-325.
Only tests/ may import this.

batch_vector: the reference's own statements, verbatim semantics (Python list per scene, concatenated, int64).
ball_query_partial_dense: torch_points_kernels 0.6.x `ball_query(..., mode="partial_dense")` (third party, not vendored:
PARITY UNPINNED).  Its published CPU algorithm is a nanoflann radius search per query (strict d^2 < r^2, squared L2 accumulated
x, y, z in float) whose first `max_num` results are kept, -1 padded; with sort=False the order of the results is the kd-tree
traversal order, which the library does not specify.  This restatement fixes the order as (d^2, index) ascending and keeps
the closest `max_num`; as a set it equals the library's result whenever at most `max_num` points are in range."""
import numpy as np


def batch_vector(offset):
    offset = np.asarray(offset, dtype=np.int64)
    counts = offset.copy()
    counts[1:] = offset[1:] - offset[:-1]                     # train.py:319-320
    return np.concatenate([np.full(int(o), ii, dtype=np.int64) for ii, o in enumerate(counts)]) if len(counts) else np.zeros(0, np.int64)


def ball_query_partial_dense(radius, max_num, x, y, batch_x=None, batch_y=None):
    x = np.asarray(x, dtype=np.float32)
    y = np.asarray(y, dtype=np.float32)
    bx = np.zeros(len(x), np.int64) if batch_x is None else np.asarray(batch_x, np.int64)
    by = np.zeros(len(y), np.int64) if batch_y is None else np.asarray(batch_y, np.int64)
    r2 = np.float32(radius) * np.float32(radius)
    idx = np.full((len(y), max_num), -1, dtype=np.int64)
    dist2 = np.full((len(y), max_num), -1.0, dtype=np.float32)
    for s in np.unique(by):
        sx = np.nonzero(bx == s)[0]
        sy = np.nonzero(by == s)[0]
        if len(sx) == 0:
            continue
        for c0 in range(0, len(sy), 512):
            qs = sy[c0:c0 + 512]
            e = x[sx][None, :, :] - y[qs][:, None, :]          # fp32 subtractions
            d = (e[..., 0] * e[..., 0] + e[..., 1] * e[..., 1]) + e[..., 2] * e[..., 2]   # fp32, left to right, no fma
            for row, q in enumerate(qs):
                cand = np.nonzero(d[row] < r2)[0]
                order = np.lexsort((sx[cand], d[row][cand]))[:max_num]
                idx[q, :len(order)] = sx[cand][order]
                dist2[q, :len(order)] = d[row][cand][order]
    return idx, dist2
