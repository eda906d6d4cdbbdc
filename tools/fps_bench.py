"""FPS micro-benchmark: time stb200_furthestsampling for the four S3DIS layer sizes (8 scenes)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stratified_transformer_b200 import pointops

def run(b, n, ds, reps=3):
    g = torch.Generator(device="cuda").manual_seed(0)
    xyz = torch.rand(b * n, 3, device="cuda", generator=g) * torch.tensor([8.0, 6.0, 3.0], device="cuda")
    off = (torch.arange(1, b + 1, device="cuda") * n).int()
    new_off = (torch.arange(1, b + 1, device="cuda") * (n // ds + 1)).int()
    pointops.furthestsampling(xyz, off, new_off); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); pointops.furthestsampling(xyz, off, new_off); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    m = n // ds + 1
    t = min(ts)
    print(f"b={b} n={n} m={m}: {t:8.3f} ms  {1e3*t/m:7.3f} us/iter  env cluster={os.environ.get('STB200_FPS_CLUSTER','-')} threads={os.environ.get('STB200_FPS_THREADS','-')}", flush=True)

if __name__ == "__main__":
    for n in (80000, 20001, 5001, 1251):
        run(8, n, 8)
