"""CPU oracle of the fused-kernel work plan (TEST INFRASTRUCTURE).

NumPy restatement of what `stb200_fused_plan_*` (stratified_transformer_b200/csrc/fused_plan.cu) builds on the
device: the window partition of `grid_sample` (/root/reference/model/stratified_transformer.py:44-65), the pair
structure of `get_indice_pairs` (:10-42) and the relative position index (:186-188) re-expressed as

  * dense tiles : for every small window W (points in ascending id = canonical order) the [n x n] matrix of packed
                  rel-pos bins  r0 | r1<<8 | r2<<16  of (query i, key j), row major, windows back to back
  * sparse tiles: for every large window V the [n_V x n_s(V)] matrix (queries = points of V, keys = its FPS-sampled
                  points, both ascending id) with bit 31 set where the reference drops the pair
                  (window_coord(query) == window_coord(key), :28-35)
  * items       : blocks of at most BQ query rows x BK key rows (stb200::fw::Item), grouped by key-chunk ordinal

The union of the valid tile entries is exactly the pair multiset of the reference (checked against
oracle.index_oracle.build_layer_index in tests/test_fused_emu.py).
"""
from __future__ import annotations

import numpy as np

from . import index_oracle as io

f32 = np.float32
F_PACKED, F_FIRST, F_FINAL, F_KEY_ATOMIC = 1, 2, 4, 8
REL_INVALID = np.uint32(0x80000000)
SEG = 128          # windows per greedy-packing segment (device: one thread per segment)
MAXORD = 8


def _partition(xyz, batch, w, parity, large):
    W = np.full(3, w, f32)
    size = (f32(2) * W).astype(f32) if large else W
    if parity % 2 == 0:
        c = io.voxel_grid(xyz, batch, size, None)
    else:
        shift = (f32(0.5) * size).astype(f32)
        c = io.voxel_grid((xyz + shift).astype(f32), batch, size, xyz.min(0))
    order = np.argsort(c, kind="stable").astype(np.int32)
    sc = c[order]
    flag = np.ones(sc.shape[0], bool)
    flag[1:] = sc[1:] != sc[:-1]
    wstart = np.concatenate([np.nonzero(flag)[0], [sc.shape[0]]]).astype(np.int32)
    pos_win = (np.cumsum(flag) - 1).astype(np.int32)
    return order, wstart, pos_win


def pack_rel(rel):
    rel = rel.astype(np.int64)
    return (rel[..., 0] | (rel[..., 1] << 8) | (rel[..., 2] << 16)).astype(np.uint32)


def build(xyz, offset, window_size, quant_size, parity, downsample_idx, BQ=64, BK=64, BQS=48, BKS=32, swin_shift=None):
    """-> dict(dense=..., sparse=... or None).  swin_shift: use the 3DSwin rel-pos index (per-point quantised)."""
    xyz = np.asarray(xyz, f32)
    N = xyz.shape[0]
    batch = io.batch_from_offset(offset)
    order_s, wstart_s, pos_win = _partition(xyz, batch, window_size, parity, False)
    n_win = wstart_s.shape[0] - 1
    sizes = np.diff(wstart_s).astype(np.int64)
    tile_base = np.concatenate([[0], np.cumsum(sizes * sizes)]).astype(np.int64)
    drel = np.zeros(int(tile_base[-1]), np.uint32)
    for wdx in range(n_win):
        ids = order_s[wstart_s[wdx]:wstart_s[wdx + 1]].astype(np.int64)
        n = ids.shape[0]
        i0 = np.repeat(ids, n)
        i1 = np.tile(ids, n)
        if swin_shift is None:
            rel = io.rel_pos_index_stratified(xyz, i0, i1, window_size, quant_size)
        else:
            rel = io.rel_pos_index_swin(xyz, i0, i1, window_size, quant_size, swin_shift)
        drel[tile_base[wdx]:tile_base[wdx + 1]] = pack_rel(rel)
    # ---- dense items: greedy packing of consecutive small windows, per segment of SEG windows
    items = [[] for _ in range(MAXORD)]
    for s0 in range(0, n_win, SEG):
        cur_start, cur_rows = -1, 0

        def flush():
            nonlocal cur_start, cur_rows
            if cur_rows:
                items[0].append((cur_start, cur_rows, cur_start, cur_rows, 0, 0, F_PACKED | F_FIRST, 0))
            cur_start, cur_rows = -1, 0
        for wdx in range(s0, min(s0 + SEG, n_win)):
            n = int(sizes[wdx])
            ws = int(wstart_s[wdx])
            if n > BK:
                flush()
                nc = (n + BK - 1) // BK
                for kc in range(nc):
                    for qc in range(nc):
                        qa, ka = qc * BQ, kc * BK
                        items[kc].append((ws + qa, min(BQ, n - qa), ws + ka, min(BK, n - ka),
                                          int(tile_base[wdx]) + qa * n + ka, n,
                                          (F_FIRST if kc == 0 else 0) | F_KEY_ATOMIC | (0x100 if kc == nc - 1 else 0), 0))
                continue
            if cur_rows + n > BQ:
                flush()
            if cur_rows == 0:
                cur_start = ws
            cur_rows += n
        flush()
    has_sparse = downsample_idx is not None and len(downsample_idx) > 0
    dense_items = []
    dense_counts = []
    for kc in range(MAXORD):
        arr = np.array(items[kc], np.int32).reshape(-1, 8)
        # 0x100 marks "last key chunk of its window"; packed items are always last
        last = ((arr[:, 6] & 0x100) != 0) | ((arr[:, 6] & F_PACKED) != 0)
        arr[:, 6] &= 0xff
        if not has_sparse:
            arr[last, 6] |= F_FINAL
        dense_items.append(arr)
        dense_counts.append(arr.shape[0])
    dense = dict(items=np.concatenate(dense_items), counts=np.array(dense_counts, np.int32), q_order=order_s, k_order=order_s,
                 rel=drel, pos_win=pos_win, wstart=wstart_s, tile_base=tile_base[:-1].astype(np.int32), n_win=n_win,
                 max_win=int(sizes.max()), BQ=BQ, BK=BK)
    if not has_sparse:
        return dict(dense=dense, sparse=None, N=N)

    # ---- sparse part
    order_l, wstart_l, _ = _partition(xyz, batch, window_size, parity, True)
    n_win_l = wstart_l.shape[0] - 1
    mask = np.zeros(N, bool)
    mask[np.asarray(downsample_idx, np.int64)] = True
    sflag = mask[order_l]
    spos = np.concatenate([[0], np.cumsum(sflag)]).astype(np.int64)
    samp = order_l[sflag]
    sstart = spos[wstart_l]                                     # [n_win_l+1] sample range of every large window
    Wv = np.full(3, window_size, f32)
    mn = xyz.min(0)
    if parity % 2 == 0:
        wc = io.floor_div_f32((xyz - mn).astype(f32), Wv)
    else:
        wc = io.floor_div_f32(((xyz + (f32(0.5) * Wv).astype(f32)).astype(f32) - mn).astype(f32), Wv)
    nv = np.diff(wstart_l).astype(np.int64)
    ns = np.diff(sstart).astype(np.int64)
    sbase = np.concatenate([[0], np.cumsum(nv * ns)]).astype(np.int64)
    srel = np.zeros(int(sbase[-1]), np.uint32)
    sitems = [[] for _ in range(MAXORD)]
    for v in range(n_win_l):
        qs = order_l[wstart_l[v]:wstart_l[v + 1]].astype(np.int64)
        ks = samp[sstart[v]:sstart[v + 1]].astype(np.int64)
        n_v, n_s = qs.shape[0], ks.shape[0]
        if n_s:
            i0 = np.repeat(qs, n_s)
            i1 = np.tile(ks, n_v)
            word = pack_rel(io.rel_pos_index_stratified(xyz, i0, i1, window_size, quant_size))
            same = (wc[i0] == wc[i1]).all(-1)
            word[same] |= REL_INVALID
            srel[sbase[v]:sbase[v + 1]] = word
        nkc = max(1, (n_s + BKS - 1) // BKS)
        if nkc > MAXORD:
            raise ValueError(f"large window with {n_s} sampled keys needs more than {MAXORD} key chunks of {BKS}")
        nqc = (n_v + BQS - 1) // BQS
        for kc in range(nkc):
            for qc in range(nqc):
                qa, ka = qc * BQS, kc * BKS
                sitems[kc].append((int(wstart_l[v]) + qa, min(BQS, n_v - qa), int(sstart[v]) + ka, max(0, min(BKS, n_s - ka)),
                                   int(sbase[v]) + qa * n_s + ka, n_s, F_KEY_ATOMIC | (F_FINAL if kc == nkc - 1 else 0), 0))
    arrs = [np.array(x, np.int32).reshape(-1, 8) for x in sitems]
    sparse = dict(items=np.concatenate(arrs), counts=np.array([a.shape[0] for a in arrs], np.int32), q_order=order_l,
                  k_order=samp.astype(np.int32), rel=srel, n_win=n_win_l, max_ns=int(ns.max()) if ns.size else 0,
                  BQ=BQS, BK=BKS)
    return dict(dense=dense, sparse=sparse, N=N)


def pairs_from_plan(plan):
    """(index_0, index_1, rel [M,3]) of every valid tile entry: dense entries first, then sparse ones (any order)."""
    d = plan["dense"]
    i0p, i1p, rp = [], [], []
    ws = d["wstart"]
    base = np.concatenate([d["tile_base"].astype(np.int64), [d["rel"].shape[0]]])
    for wdx in range(d["n_win"]):
        ids = d["q_order"][ws[wdx]:ws[wdx + 1]].astype(np.int64)
        n = ids.shape[0]
        i0p.append(np.repeat(ids, n)); i1p.append(np.tile(ids, n)); rp.append(d["rel"][base[wdx]:base[wdx] + n * n])
    s = plan["sparse"]
    if s is not None:
        seen = set()
        for it in s["items"]:
            q_pos, nq, k_pos, nk, rel_off, pitch = (int(x) for x in it[:6])
            for r in range(nq):
                for j in range(nk):
                    w = s["rel"][rel_off + r * pitch + j]
                    if w & REL_INVALID:
                        continue
                    key = (q_pos + r, k_pos + j)
                    assert key not in seen
                    seen.add(key)
                    i0p.append(np.array([s["q_order"][q_pos + r]], np.int64)); i1p.append(np.array([s["k_order"][k_pos + j]], np.int64))
                    rp.append(np.array([w], np.uint32))
    i0, i1, r = np.concatenate(i0p), np.concatenate(i1p), np.concatenate(rp)
    rel = np.stack([r & 0xff, (r >> 8) & 0xff, (r >> 16) & 0xff], 1).astype(np.int32)
    return i0, i1, rel
