#!/usr/bin/env python
"""bench.py — points/sec fwd+bwd of the Stratified-Transformer window-attention hot path on S3DIS-shape scenes.

Workload (BASELINE.json configs[1], restricted to the hot path of SURVEY §8): per GPU 8 synthetic 80k-point
scenes; the S3DIS layer schedule of config/s3dis/s3dis_stratified_transformer.yaml (channels 48/96/192/384,
heads 3/6/12/24, depths 2/2/6/2, window 0.16*2^l, quant 0.01*2^l, downsample_scale 8, point hierarchy
n -> int(n*0.25)+1 per TransitionDown).  One step = for every layer: FPS of the stratified keys + pair-index
construction for both block parities (+ rel-pos index), then for every block: attention_step1 +
dot_prod_with_idx (one fused pass) -> segment softmax -> attention_step2_with_rel_pos_value forward and the full backward
(grads of q, k, v and the three tables).  The dense GEMMs around the path (qkv/proj Linear, MLP, KPConv) and
TransitionDown/Upsample are out of scope; the point hierarchy is precomputed data.

  value : device-resident inputs, extension-level calls, CUDA-event timed, max over ranks.
  e2e   : the same path through the public module API (WindowAttention + autograd), inputs start in pinned HOST
          memory every step (coordinates of every layer + 6-channel point features), loss read back to the host.
  --impl reference : the reference's pure-PyTorch gather/scatter_add formulation on the host CPU cores
          (oracle/, all threads) on a bounded sample of the same workload.

Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "points/sec fwd+bwd (S3DIS 80k-pt scenes), window-attention hot path"
LAYERS = [  # channels, heads, depth, window, quant  (train.py:110-113 with the s3dis yaml)
    dict(C=48, h=3, depth=2, window=0.16, quant=0.01),
    dict(C=96, h=6, depth=2, window=0.32, quant=0.02),
    dict(C=192, h=12, depth=6, window=0.64, quant=0.04),
    dict(C=384, h=24, depth=2, window=1.28, quant=0.08),
]
DS_SCALE = 8
HEAD_DIM = 16


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scenes", type=int, default=8)
    ap.add_argument("--points", type=int, default=80000)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="build the geometry serially inside the step instead of prefetching it on a side stream")
    ap.add_argument("--fused", action="store_true", help="(round-1 experiment) per-window mma.sync forward kernel in the per-op device arm")
    ap.add_argument("--path", default="fused", choices=["fused", "perop"],
                    help="fused: window-centric fused kernels on the work plan (default); perop: the per-pair entry points on the CSR pair list")
    ap.add_argument("--no-profile", action="store_true", help="do not bracket kernels with CUDA events in the timed region")
    ap.add_argument("--cpu-sample-points", type=int, default=0, help="points of the CPU sample scene (0 = auto)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_hot_path(n_points, steps, warmup, threads=None):
    """The reference formulation on the host: numpy/torch index construction + gather/scatter_add attention with
    autograd backward (oracle/), one scene of n_points, full layer schedule.  Returns (points/s, seconds/step)."""
    from oracle import attention_oracle as ao, fps_oracle, index_oracle as io
    from stratified_transformer_b200.synthetic import make_scene
    torch.set_num_threads(threads or os.cpu_count())
    xyz0, _ = make_scene(0, n_points)
    g = torch.Generator().manual_seed(0)
    # point hierarchy (precomputed data, as for the GPU arm)
    xyzs = [xyz0]
    for lvl in range(1, len(LAYERS)):
        prev = xyzs[-1]
        off = np.array([prev.shape[0]], np.int32)
        new_off = np.array([int(prev.shape[0] * 0.25) + 1], np.int32)
        xyzs.append(np.ascontiguousarray(prev[fps_oracle.furthestsampling(prev, off, new_off)]))
    feats = []
    for lvl, cfg in enumerate(LAYERS):
        n, h = xyzs[lvl].shape[0], cfg["h"]
        L = 2 * int((2 * cfg["window"] + 1e-4) // cfg["quant"])
        feats.append(dict(q=torch.randn(n, h, HEAD_DIM, generator=g), k=torch.randn(n, h, HEAD_DIM, generator=g),
                          v=torch.randn(n, h, HEAD_DIM, generator=g), g=torch.randn(n, h, HEAD_DIM, generator=g),
                          t=[torch.randn(L, h, HEAD_DIM, 3, generator=g) * 0.02 for _ in range(3)]))

    def step():
        for lvl, cfg in enumerate(LAYERS):
            xyz = xyzs[lvl]
            off = np.array([xyz.shape[0]], np.int32)
            ds_idx = fps_oracle.furthestsampling(xyz, off, io.fps_new_offset(off, DS_SCALE))
            idx = []
            for parity in (0, 1):
                r = io.build_layer_index(xyz, off, cfg["window"], DS_SCALE, ds_idx, parity)
                rel = io.rel_pos_index_stratified(xyz, r["index_0"], r["index_1"], cfg["window"], cfg["quant"])
                idx.append((torch.from_numpy(r["offsets"]), torch.from_numpy(r["index_1"]), torch.from_numpy(rel)))
            f = feats[lvl]
            for blk in range(cfg["depth"]):
                offs, i1, rel = idx[blk % 2]
                ao.layer_autograd(f["q"], f["k"], f["v"], offs, i1, f["t"][0], f["t"][1], f["t"][2], rel, f["g"])

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return n_points / dt, dt


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    total = a.steps + a.warmup
    n = a.cpu_sample_points or int(max(4000, min(30000, 400000 / max(total, 1))))
    cores = os.cpu_count()
    pps, dt = cpu_hot_path(n, a.steps, a.warmup)
    sample = f"1 synthetic S3DIS-shape scene cropped to {n} points, full 4-layer / 12-block schedule, fp32, torch CPU {cores} threads"
    line = {
        "metric": METRIC, "value": pps, "unit": "points/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "impl": "reference",
        "config": workload_config(a, note="reference arm: CPU port of the reference's gather/scatter_add formulation"),
        "cpu_baseline": {"value": pps, "unit": "points/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": pps, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(a, note=None):
    cfg = {
        "workload": f"s3dis_stratified_attention_path: {a.scenes}x{a.points}-pt synthetic scenes per GPU, layers C/h/depth "
                    "48/3/2,96/6/2,192/12/6,384/24/2, window 0.16*2^l, quant 0.01*2^l, stratified keys ds=8; per step: FPS + "
                    "pair index (both parities) per layer, then step1+rpe+softmax+step2 fwd and bwd per block",
        "scenes_per_gpu": a.scenes, "points_per_scene": a.points, "head_dim": HEAD_DIM,
        "l2_policy": "inputs larger than L2 (q/k/v + pair arrays of layer 0 alone exceed 2 GB)",
        "parallelism": f"dp{a.gpus} by scene",
        "geometry": "serial inside the step" if getattr(a, "no_overlap", False) else
                    "FPS + pair lists of batch t+1 computed on a side stream during the attention of batch t (one full geometry per step)",
    }
    if note:
        cfg["note"] = note
    return cfg


# ------------------------------------------------------------------------------------------------ GPU arm
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.proc.wait()
        sm, mx, reasons = [], [], set()
        for ln in open(self.path):
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 8:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_inputs(a, rank, dev):
    """Scenes, point hierarchy (precomputed with the library's own FPS) and per-layer operands, on the host and device."""
    from stratified_transformer_b200 import pointops
    from stratified_transformer_b200.synthetic import make_batch
    xyz0, rgb, offset0 = make_batch(a.scenes, a.points, seed0=100 * rank)
    levels = []
    xyz_d = torch.from_numpy(xyz0).to(dev)
    off_d = torch.from_numpy(offset0).to(dev)
    gen = torch.Generator(device=dev).manual_seed(1 + rank)
    sub_idx = None
    for lvl, cfg in enumerate(LAYERS):
        if lvl > 0:
            counts = torch.diff(off_d, prepend=off_d.new_zeros(1))
            new_off = torch.cumsum((counts.double() * 0.25).long() + 1, 0).int()
            sub_idx = pointops.furthestsampling(xyz_d, off_d, new_off)
            xyz_d = xyz_d[sub_idx.long()].contiguous()
            off_d = new_off
        n, h = xyz_d.shape[0], cfg["h"]
        L = 2 * int((2 * cfg["window"] + 1e-4) // cfg["quant"])
        lv = dict(cfg=cfg, xyz=xyz_d, offset=off_d, L=L, sub_idx=sub_idx,
                  q=torch.randn(n, h, HEAD_DIM, device=dev, generator=gen) * (HEAD_DIM ** -0.5 * 4),
                  k=torch.randn(n, h, HEAD_DIM, device=dev, generator=gen),
                  v=torch.randn(n, h, HEAD_DIM, device=dev, generator=gen),
                  g=torch.randn(n, h, HEAD_DIM, device=dev, generator=gen),
                  tables=[[torch.nn.init.trunc_normal_(torch.empty(L, h, HEAD_DIM, 3, device=dev), std=0.02) for _ in range(3)]
                          for _ in range(cfg["depth"])])
        levels.append(lv)
    return levels, torch.from_numpy(rgb)


def device_step(levels, grads_out, geo=None, no_fused=True):
    """One pass of the hot path with device-resident operands through the extension-level API (fused entry points:
    logits = q.k + rel-pos bias in one pass, segment softmax, aggregation; and their single-pass gradients)."""
    import ctypes
    from stratified_transformer_b200 import _cabi, index, pointops2_cuda as ext
    stream = torch.cuda.current_stream().cuda_stream
    for lvl, lv in enumerate(levels):
        cfg = lv["cfg"]
        h, L = cfg["h"], lv["L"]
        li = geo[lvl] if geo is not None else index.build_layer_index(lv["xyz"], lv["offset"], cfg["window"], cfg["quant"], DS_SCALE)
        q, k, v, g = lv["q"], lv["k"], lv["v"], lv["g"]
        N = q.shape[0]
        dev = q.device
        for blk in range(cfg["depth"]):
            pi = li.for_block(blk)
            M, off = pi.M, pi.index_0_offsets
            tq, tk, tv = lv["tables"][blk]
            ix = pi.c_struct(L, backward=True)
            s = torch.empty(M, h, device=dev); p = torch.empty(M, h, device=dev)
            out = torch.empty(N, h, HEAD_DIM, device=dev)
            plan = None if no_fused else pi.fused_plan()
            rows = None
            if plan is not None:      # per-window tensor-core kernel for the windows with one shared key list
                flags, rows = plan
                _cabi.call("stb200_window_attention_forward_fused", ctypes.byref(ix), pi.n_win, pi.win_offsets.data_ptr(),
                           flags.data_ptr(), h, HEAD_DIM, L, q.data_ptr(), k.data_ptr(), v.data_ptr(), tq.data_ptr(), tk.data_ptr(),
                           tv.data_ptr(), out.data_ptr(), p.data_ptr(), stream)
            if plan is None or rows.numel() > 0:   # per-pair kernels on the remaining rows
                ixf = pi.c_struct(L, backward=True)
                if rows is not None:
                    ixf.row_order, ixf.N = rows.data_ptr(), rows.numel()
                _cabi.call("stb200_window_logits_forward", ctypes.byref(ixf), h, HEAD_DIM, L, q.data_ptr(), k.data_ptr(),
                           tq.data_ptr(), tk.data_ptr(), s.data_ptr(), stream)
                _cabi.call("stb200_segment_softmax_forward_rows", ixf.N, None if rows is None else rows.data_ptr(), h,
                           s.data_ptr(), None, off.data_ptr(), p.data_ptr(), stream)
                _cabi.call("stb200_window_aggregate_forward", ctypes.byref(ixf), h, HEAD_DIM, L, p.data_ptr(), v.data_ptr(),
                           tv.data_ptr(), out.data_ptr(), stream)
            # backward
            gp = s                                      # reuse the M-sized buffer
            gv = torch.empty_like(v); gtv = torch.zeros_like(tv)
            _cabi.call("stb200_window_aggregate_backward", ctypes.byref(ix), h, HEAD_DIM, L, g.data_ptr(), p.data_ptr(),
                       v.data_ptr(), tv.data_ptr(), gp.data_ptr(), gv.data_ptr(), gtv.data_ptr(), stream)
            gs = torch.empty(M, h, device=dev)
            ext.segment_softmax_backward_cuda(N, M, h, p, gp, off, gs)
            gq = torch.empty_like(q); gk = torch.empty_like(k); gtq = torch.zeros_like(tq); gtk = torch.zeros_like(tk)
            ws = torch.empty(M * h + 64, device=dev)   # scratch: grad rows in transposed order (see include/stb200.h)
            wsb = 0 if os.environ.get("STB200_NO_PERMUTE_WS") else ws.numel() * 4
            _cabi.call("stb200_window_logits_backward_ws", ctypes.byref(ix), h, HEAD_DIM, L, gs.data_ptr(), q.data_ptr(),
                       k.data_ptr(), tq.data_ptr(), tk.data_ptr(), gq.data_ptr(), gk.data_ptr(), gtq.data_ptr(),
                       gtk.data_ptr(), ws.data_ptr(), wsb, stream)
            grads_out.append((gtq, gtk, gtv))
    return grads_out


def device_step_fused(levels, grads_out, geo):
    """One pass of the hot path on the window-centric fused kernels: per block one forward (dense pass + sparse pass) that
    keeps only the output and the row log-sum-exp, and one backward producing the six gradients.  No [M,h] tensor."""
    from stratified_transformer_b200 import _cabi
    stream = torch.cuda.current_stream().cuda_stream
    for lvl, lv in enumerate(levels):
        cfg = lv["cfg"]
        h, L = cfg["h"], lv["L"]
        li = geo[lvl]
        q, k, v, g = lv["q"], lv["k"], lv["v"], lv["g"]
        N = q.shape[0]
        dev = q.device
        for blk in range(cfg["depth"]):
            plan = li.for_block(blk).plan
            passes, n_passes = plan.passes(L)
            tq, tk, tv = lv["tables"][blk]
            out = torch.empty(N, h, HEAD_DIM, device=dev)
            lse = torch.empty(N, h, device=dev); lsum = torch.empty(N, h, device=dev)
            _cabi.call("stb200_fused_attention_forward", passes, n_passes, N, h, L, q.data_ptr(), k.data_ptr(), v.data_ptr(),
                       tq.data_ptr(), tk.data_ptr(), tv.data_ptr(), out.data_ptr(), lse.data_ptr(), lsum.data_ptr(), stream)
            gq = torch.empty_like(q)
            alloc = torch.zeros_like if plan.needs_zeroed_key_grads else torch.empty_like
            gk, gv = alloc(k), alloc(v)
            gtq, gtk, gtv = torch.zeros_like(tq), torch.zeros_like(tk), torch.zeros_like(tv)
            _cabi.call("stb200_fused_attention_backward", passes, n_passes, N, h, L, g.data_ptr(), out.data_ptr(), lse.data_ptr(),
                       q.data_ptr(), k.data_ptr(), v.data_ptr(), tq.data_ptr(), tk.data_ptr(), tv.data_ptr(), gq.data_ptr(),
                       gk.data_ptr(), gv.data_ptr(), gtq.data_ptr(), gtk.data_ptr(), gtv.data_ptr(), stream)
            grads_out.append((gtq, gtk, gtv))
    return grads_out


class HotPathModel(torch.nn.Module):
    """e2e harness: the package's WindowAttention modules in the S3DIS schedule.  The blocks outside the hot path are
    replaced by stand-ins that keep shapes right (a Linear stem 6 -> 48, and `Linear(C_l -> C_{l+1})` on the points the
    precomputed hierarchy keeps instead of TransitionDown); residual connection around each attention block."""

    def __init__(self, fused=True):
        super().__init__()
        self.fused = fused
        from stratified_transformer_b200.window_attention import WindowAttention
        self.stem = torch.nn.Linear(6, LAYERS[0]["C"])
        self.down = torch.nn.ModuleList([torch.nn.Linear(LAYERS[i]["C"], LAYERS[i + 1]["C"]) for i in range(len(LAYERS) - 1)])
        self.blocks = torch.nn.ModuleList([
            torch.nn.ModuleList([WindowAttention(c["C"], c["window"], c["h"], c["quant"], rel_query=True, rel_key=True,
                                                 rel_value=True) for _ in range(c["depth"])]) for c in LAYERS])

    def forward(self, feat6, xyzs, offsets, sub_idx, geo=None):
        from stratified_transformer_b200 import index
        feats = self.stem(feat6)
        for lvl, cfg in enumerate(LAYERS):
            if lvl > 0:
                feats = self.down[lvl - 1](feats[sub_idx[lvl].long()])
            li = geo[lvl] if geo is not None else index.build_layer_index(xyzs[lvl], offsets[lvl], cfg["window"], cfg["quant"], DS_SCALE,
                                                                          fused=self.fused, csr=not self.fused)
            for blk, attn in enumerate(self.blocks[lvl]):
                feats = feats + attn(feats, xyzs[lvl], li.for_block(blk))
        return feats.float().pow(2).mean()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference_arm(a)
        return
    from stratified_transformer_b200 import _cabi
    _cabi.load()   # fail loudly if the CUDA library is missing: there is no fallback
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference for the CPU port)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"   # keep stdout to the single JSON line (NCCL prints its version there)
        dist.init_process_group("nccl", device_id=dev)

    if os.environ.get("STB200_MAIN_PRIORITY"):   # development knob: run the attention stream at another priority
        torch.cuda.set_stream(torch.cuda.Stream(device=dev, priority=int(os.environ["STB200_MAIN_PRIORITY"])))
    levels, rgb = build_inputs(a, rank, dev)
    n_points = levels[0]["xyz"].shape[0]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    from stratified_transformer_b200 import index as st_index
    geo_cfgs = [(lv["cfg"]["window"], lv["cfg"]["quant"], DS_SCALE, lv["L"]) for lv in levels]
    xyzs_d = [lv["xyz"] for lv in levels]
    offs_d = [lv["offset"] for lv in levels]
    offs_h = [lv["offset"].cpu().tolist() for lv in levels]
    pf = None
    use_fused = a.path == "fused"
    if not a.no_overlap:
        # geometry (FPS + pair lists) of the NEXT batch runs on a side stream under the attention of the current one;
        # every step still computes one complete geometry from the coordinates
        pf = st_index.GeometryPrefetcher(geo_cfgs, dev, fused=use_fused, csr=not use_fused)
        pf.submit(xyzs_d, offs_d, offs_h)

    def one_step():
        geo = None
        if pf is not None:
            geo = pf.take()
            pf.submit(xyzs_d, offs_d, offs_h)
        if use_fused:
            if geo is None:
                geo = [st_index.build_layer_index(lv["xyz"], lv["offset"], lv["cfg"]["window"], lv["cfg"]["quant"], DS_SCALE,
                                                  fused=True, csr=False) for lv in levels]
            grads = device_step_fused(levels, [], geo)
        else:
            grads = device_step(levels, [], geo, not a.fused)
        if pf is not None:
            pf.complete()
        if dist is not None:   # training only: data-parallel gradient all-reduce of the attention parameters
            flat = torch.cat([t.reshape(-1) for trip in grads for t in trip])
            dist.all_reduce(flat)
        return grads

    for _ in range(max(a.warmup, 3)):
        one_step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    _cabi.profile_dump()
    _cabi.profile_enable(not a.no_profile)
    launches0 = _cabi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    marks = []
    for _ in range(a.steps):
        one_step()
        if os.environ.get("STB200_BENCH_STEP_TIMES"):   # development aid: per-step device times on stderr
            marks.append(torch.cuda.Event(enable_timing=True)); marks[-1].record()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / a.steps
    if marks and rank == 0:
        ts = [e0.elapsed_time(marks[0])] + [marks[i - 1].elapsed_time(marks[i]) for i in range(1, len(marks))]
        print("step ms:", " ".join(f"{t:.1f}" for t in ts), file=sys.stderr)
    _cabi.profile_enable(False)
    prof = _cabi.profile_dump()
    launches = (_cabi.launch_count() - launches0) // a.steps
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = n_points * world / (ms * 1e-3)

    # ---- e2e: module API, host inputs ----
    e2e = None
    if not a.no_e2e:
        torch.manual_seed(0)
        model = HotPathModel(use_fused).to(dev)
        if dist is not None:
            model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local])
        host = dict(feat6=torch.cat([rgb, levels[0]["xyz"].cpu()], 1).pin_memory(),
                    xyz=[lv["xyz"].cpu().pin_memory() for lv in levels],
                    off=[lv["offset"].cpu().pin_memory() for lv in levels],
                    sub=[None] + [lv["sub_idx"].cpu().pin_memory() for lv in levels[1:]])
        h2d = host["feat6"].numel() * 4 + sum(x.numel() * 4 for x in host["xyz"]) + sum(o.numel() * 4 for o in host["off"]) + \
            sum(s.numel() * 4 for s in host["sub"] if s is not None)

        main = torch.cuda.current_stream()
        pf2 = None if a.no_overlap else st_index.GeometryPrefetcher(geo_cfgs, dev, fused=use_fused, csr=not use_fused)

        def upload(stream):
            with torch.cuda.stream(stream):
                d = dict(feat6=host["feat6"].to(dev, non_blocking=True),
                         xyz=[x.to(dev, non_blocking=True) for x in host["xyz"]],
                         off=[o.to(dev, non_blocking=True) for o in host["off"]],
                         sub=[None if s is None else s.to(dev, non_blocking=True) for s in host["sub"]])
            return d

        state = {}
        if pf2 is not None:
            state["next"] = upload(pf2.side)
            pf2.submit(state["next"]["xyz"], state["next"]["off"], offs_h)

        def e2e_step():
            if pf2 is not None:   # inputs + geometry of this batch were prefetched during the previous step
                cur, geo = state["next"], pf2.take()
                for t in [cur["feat6"]] + cur["xyz"] + cur["off"] + [s for s in cur["sub"] if s is not None]:
                    t.record_stream(main)
                state["next"] = upload(pf2.side)
                pf2.submit(state["next"]["xyz"], state["next"]["off"], offs_h)
            else:
                cur, geo = upload(main), None
            model.zero_grad(set_to_none=True)
            # the reference trains under AMP (config use_amp: True, train.py:336): Linear layers in reduced precision,
            # the pair ops pinned to fp32 (custom_fwd cast_inputs, like the reference's .float() call sites)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                loss = model(cur["feat6"], cur["xyz"], cur["off"], cur["sub"], geo)
            loss.backward()
            if pf2 is not None:
                pf2.complete()
            return float(loss.item())   # D2H read of the step result

        for _ in range(2):
            e2e_step()
        barrier()
        n_e2e = max(3, min(a.steps, 10))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / n_e2e
        tt = torch.tensor([dt], device=dev)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": n_points * world / float(tt.item()), "unit": "points/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": 4, "ms_per_step": float(tt.item()) * 1e3, "steps": n_e2e,
               "api": "WindowAttention modules (autograd, bf16 autocast around the Linear layers as in the reference's AMP recipe; "
                      "pair ops fp32) + index builder, pinned host inputs"}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (live CUDA events from the timed region) ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    roofline = None
    kern_ms = {k: v["ms"] for k, v in prof.items()}
    if kern_ms:
        total_k = sum(kern_ms.values())
        attn = {k: v for k, v in prof.items() if v["bytes"] > 0}
        top = max(attn, key=lambda k: attn[k]["ms"])
        tv = prof[top]
        ach = tv["bytes"] / (tv["ms"] * 1e-3) / 1e9
        traffic = None
        try:   # DRAM bytes per launch of this kernel from the committed ncu --set full capture (profiles/)
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json"))).get(top, {}).get("dram_bytes_per_launch")
        except OSError:
            pass
        roofline = {"bound": "hbm", "kernel": top, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback 6650 GB/s",
                    "traffic": traffic, "algorithmic_bytes_per_launch": tv["bytes"] / tv["launches"],
                    "launches": tv["launches"], "avg_ms": tv["ms"] / tv["launches"],
                    "share_of_kernel_time": tv["ms"] / total_k,
                    "per_kernel_ms_per_step": {k: round(v / a.steps, 4) for k, v in sorted(kern_ms.items(), key=lambda kv: -kv[1])}}

    cpu_baseline = None
    if not a.no_cpu_baseline and world == 1:
        n = a.cpu_sample_points or 24000
        pps, dt = cpu_hot_path(n, 1, 1)
        cpu_baseline = {"value": pps, "unit": "points/s", "cores": os.cpu_count(), "kind": "port",
                        "sample": f"1 scene cropped to {n} points, full 4-layer/12-block schedule, 1 warm-up + 1 timed pass ({dt:.1f} s)"}

    line = {
        "metric": METRIC, "value": value, "unit": "points/s", "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(a), "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        "roofline": roofline, "cpu_baseline": cpu_baseline,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
