#!/usr/bin/env python
"""Development tool: per-pass times of the fused window-attention kernels on one synthetic S3DIS-shape batch, per layer of the
S3DIS schedule (library profiler: CUDA events around every launch).  `--level L` restricts to one layer (for ncu)."""
import argparse, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stratified_transformer_b200 import _cabi, index, pointops
from stratified_transformer_b200.synthetic import make_batch

ap = argparse.ArgumentParser()
ap.add_argument("--scenes", type=int, default=2)
ap.add_argument("--points", type=int, default=80000)
ap.add_argument("--level", type=int, default=-1)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--perop", action="store_true", help="also time the per-op path on the same index")
a = ap.parse_args()
xyz, _, offset = make_batch(a.scenes, a.points, seed0=0)
xd, od = torch.from_numpy(xyz).cuda(), torch.from_numpy(offset).cuda()
res = {}
for lvl in range(4):
    if lvl > 0:
        counts = torch.diff(od, prepend=od.new_zeros(1))
        new_off = torch.cumsum((counts.double() * 0.25).long() + 1, 0).int()
        sub = pointops.furthestsampling(xd, od, new_off)
        xd, od = xd[sub.long()].contiguous(), new_off
    if a.level >= 0 and lvl != a.level:
        continue
    w, qz, h, L = 0.16 * 2 ** lvl, 0.01 * 2 ** lvl, 3 * 2 ** lvl, 64
    li = index.build_layer_index(xd, od, w, qz, 8, fused=True, csr=a.perop)
    N = xd.shape[0]
    g = torch.Generator(device="cuda").manual_seed(lvl)
    q, k, v, go = (torch.randn(N, h, 16, device="cuda", generator=g) for _ in range(4))
    tq, tk, tv = (torch.randn(L, h, 16, 3, device="cuda", generator=g) * 0.02 for _ in range(3))
    for parity in (0, 1):
        pi = li.for_block(parity)
        plan = pi.plan
        leaves = [t.clone().requires_grad_(True) for t in (q, k, v, tq, tk, tv)]
        for rep in range(a.reps + 1):
            if rep == 1:
                torch.cuda.synchronize(); _cabi.profile_dump(); _cabi.profile_enable(True)
            out = pointops.window_attention_plan(*leaves, plan)
            out.backward(go)
            if a.perop:
                s = pointops.window_logits(leaves[0], leaves[1], leaves[3], leaves[4], pi)
                p = pointops.segment_softmax(s, pi.index_0_offsets)
                o2 = pointops.window_aggregate(p, leaves[2], leaves[5], pi)
                o2.backward(go)
        torch.cuda.synchronize(); _cabi.profile_enable(False)
        prof = _cabi.profile_dump()
        t = plan.totals
        res[f"L{lvl}p{parity}"] = dict(N=N, h=h, dense_words=t[0], sparse_words=t[1], max_win=t[3], max_ns=t[4],
                                        dense_items=t[8:12], sparse_items=t[16:20],
                                        ms={kk: round(vv["ms"] / a.reps, 4) for kk, vv in sorted(prof.items())})
print(json.dumps(res, indent=1))
