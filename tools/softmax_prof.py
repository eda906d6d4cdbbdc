"""Segment-softmax kernels alone on a layer-0-like CSR (N rows of ~33 pairs, h heads): time per launch and GB/s; run under
ncu for the details (`ncu --set full -k regex:softmax python tools/softmax_prof.py --reps 1`)."""
import argparse, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stratified_transformer_b200 import pointops2_cuda as ext

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=640000)
ap.add_argument("--h", type=int, default=3)
ap.add_argument("--kappa", type=int, default=33)
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
rng = np.random.default_rng(0)
lens = np.clip(rng.normal(a.kappa, a.kappa / 3, a.rows).astype(np.int64), 1, 4 * a.kappa)
off = torch.from_numpy(np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)).cuda()
M = int(off[-1])
x = torch.randn(M, a.h, device="cuda")
g = torch.randn(M, a.h, device="cuda")
p = torch.empty_like(x)
gs = torch.empty_like(x)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, fn, nbytes in (("fwd", lambda: ext.segment_softmax_forward_cuda(a.rows, M, a.h, x, None, off, p), 8.0 * M * a.h),
                         ("bwd", lambda: ext.segment_softmax_backward_cuda(a.rows, M, a.h, p, g, off, gs), 12.0 * M * a.h)):
    fn()
    ts = []
    for _ in range(a.reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = float(np.median(ts))
    print(f"{name}: N={a.rows} M={M} h={a.h} variant={os.environ.get('STB200_SOFTMAX_VARIANT', '0')}: {t:.3f} ms = {nbytes / t / 1e6:.0f} GB/s")
