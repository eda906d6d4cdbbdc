"""Per-operator timing on a B200: new kernels vs the reference's own kernels (oracle/_ref) on the same
S3DIS-shape index (one or more synthetic scenes, stratified pairs).  Development tool (uses oracle/ to
build the index and to run the reference kernels; not part of the product path).

  python tools/op_bench.py --scenes 1 --points 80000 [--layer 0] [--ref 1]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import fps_oracle, index_oracle as io, ref_cuda  # noqa: E402
from stratified_transformer_b200 import pointops2_cuda as ext  # noqa: E402
from stratified_transformer_b200.synthetic import make_batch  # noqa: E402


def timed(fn, reps=10, warm=3, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if flush is not None:
            flush.add_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenes", type=int, default=1)
    ap.add_argument("--points", type=int, default=80000)
    ap.add_argument("--heads", type=int, default=3)
    ap.add_argument("--window", type=float, default=0.16)
    ap.add_argument("--quant", type=float, default=0.01)
    ap.add_argument("--ds", type=int, default=8)
    ap.add_argument("--ref", type=int, default=1)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    h, d = a.heads, 16
    t0 = time.time()
    xyz, _, offset = make_batch(a.scenes, a.points)
    new_offset = io.fps_new_offset(offset, a.ds)
    ds_idx = fps_oracle.furthestsampling(xyz, offset, new_offset)
    r = io.build_layer_index(xyz, offset, a.window, a.ds, ds_idx, 0)
    rel = io.rel_pos_index_stratified(xyz, r["index_0"], r["index_1"], a.window, a.quant)
    N, M = xyz.shape[0], r["index_1"].shape[0]
    L = 2 * int((2 * a.window + 1e-4) // a.quant)
    print(f"index built on CPU in {time.time()-t0:.1f}s: N={N} M={M} kappa={M/N:.1f} n_max={r['n_max']} L={L} rel range {rel.min()}..{rel.max()}", flush=True)
    dev = torch.device("cuda")
    g = torch.Generator(device="cuda").manual_seed(1)
    q, k, v, go = (torch.randn(N, h, d, device=dev, generator=g) for _ in range(4))
    tq, tk, tv = (torch.randn(L, h, d, 3, device=dev, generator=g) * 0.02 for _ in range(3))
    off = torch.from_numpy(r["offsets"]).to(dev).int()
    i1 = torch.from_numpy(r["index_1"]).to(dev).int()
    rel_d = torch.from_numpy(rel).to(dev).int().contiguous()
    attn = torch.empty(M, h, device=dev); bias = torch.empty(M, h, device=dev); p = torch.empty(M, h, device=dev)
    out = torch.empty(N, h, d, device=dev)
    gp = torch.empty(M, h, device=dev); gs = torch.empty(M, h, device=dev)
    gq = torch.empty(N, h, d, device=dev); gk = torch.zeros(N, h, d, device=dev); gv = torch.zeros(N, h, d, device=dev)
    gq2 = torch.empty(N, h, d, device=dev); gk2 = torch.zeros(N, h, d, device=dev)
    gtq, gtk, gtv = (torch.zeros(L, h, d, 3, device=dev) for _ in range(3))
    flush = torch.empty(64 * 1024 * 1024, device=dev)  # 256 MB > L2
    C = h * d
    t_build = timed(lambda: ext.build_transposed_csr(off, i1), flush=flush)
    tc = ext.build_transposed_csr(off, i1)
    ops = {
        "step1_fwd": lambda: ext.attention_step1_forward_cuda_v2(N, M, h, C, 0, q, k, off, i1, attn),
        "rpe_fwd": lambda: ext.dot_prod_with_idx_forward_cuda_v3(N, M, h, d, 0, q, off, k, i1, tq, tk, rel_d, bias),
        "softmax_fwd": lambda: ext.segment_softmax_forward_cuda(N, M, h, attn, bias, off, p),
        "step2_fwd": lambda: ext.attention_step2_with_rel_pos_value_forward_cuda_v2(N, M, h, d, 0, p, v, off, i1, tv, rel_d, out),
        "step2_bwd": lambda: ext.attention_step2_with_rel_pos_value_backward_cuda_v2(N, M, h, d, 0, go, off, i1, p, v, tv, rel_d, gp, gv, gtv, tc),
        "softmax_bwd": lambda: ext.segment_softmax_backward_cuda(N, M, h, p, gp, off, gs),
        "rpe_bwd": lambda: ext.dot_prod_with_idx_backward_cuda_v3(N, M, h, d, 0, gs, q, off, k, i1, tq, tk, rel_d, gq2, gk2, gtq, gtk, tc),
        "step1_bwd": lambda: ext.attention_step1_backward_cuda_v2(N, M, h, C, 0, gs, off, i1, q, k, gq, gk, tc),
    }
    i0 = torch.repeat_interleave(torch.arange(N, device=dev, dtype=torch.int32), (off[1:] - off[:-1]).long())
    ops["v1_step1_bwd(red.v4 scatter)"] = lambda: ext.attention_step1_backward_cuda(N, M, h, C, gs, i0, i1, q, k, gq, gk)
    ops["v1_step2_fwd(red.v4 scatter)"] = lambda: ext.attention_step2_forward_cuda(N, M, h, C, p, v, i0, i1, out)
    res = {"N": N, "M": M, "h": h, "transpose_csr_ms": t_build}
    for name, fn in ops.items():
        res[name + "_ms"] = timed(fn, flush=flush)
    res["total_ms"] = sum(v for kk, v in res.items() if kk.endswith("_ms") and kk != "transpose_csr_ms" and not kk.startswith("v1_"))
    kap = M / N
    fwd_b = 4 * (6 * C + 9 * kap + 5 * kap * h + 4) * N
    bwd_b = 4 * (11 * C + 9 * kap + 7 * kap * h + 3) * N
    res["algorithmic_GB"] = (fwd_b + bwd_b) / 1e9
    res["achieved_GBps"] = res["algorithmic_GB"] / (res["total_ms"] * 1e-3)
    print(json.dumps(res, indent=1), flush=True)
    if a.ref and ref_cuda.available():
        n_max = r["n_max"]
        rops = {
            "step1_fwd": lambda: ref_cuda.step1_fwd(q, k, off, i1),
            "rpe_fwd": lambda: ref_cuda.rpe_fwd(q, k, off, i1, tq, tk, rel_d),
            "step2_fwd": lambda: ref_cuda.step2_rpv_fwd(p, v, off, i1, tv, rel_d),
            "step2_bwd": lambda: ref_cuda.step2_rpv_bwd(go, p, v, off, i1, tv, rel_d),
            "rpe_bwd": lambda: ref_cuda.rpe_bwd(gs, q, k, off, i1, tq, tk, rel_d),
            "step1_bwd": lambda: ref_cuda.step1_bwd(gs, q, k, off, i1),
        }
        ref = {}
        for name, fn in rops.items():
            # wrappers allocate+zero outputs and synchronise: time with host clock around a synced call, minus an empty-call estimate
            fn(); torch.cuda.synchronize()
            ts = []
            for _ in range(5):
                t = time.perf_counter(); fn(); ts.append((time.perf_counter() - t) * 1e3)
            ref[name + "_ms"] = float(np.median(ts))
        ref["total_ms"] = sum(ref.values())
        print("reference kernels (host-timed incl. output zero-fill):", json.dumps(ref, indent=1), flush=True)
        res["reference"] = ref
    if a.out:
        json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
