"""Where one drop-in BasicLayer (layer 0 of cfg2 + TransitionDown, 1 scene of 80k points) spends its forward + backward:
library kernels by CUDA events (stb200 profiler) and the wall clock of the whole call."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from stratified_transformer_b200 import _cabi, layers
from stratified_transformer_b200.synthetic import make_batch

dev = torch.device("cuda")
xyz, _, offset = make_batch(1, 80000, seed0=7)
xd, od = torch.from_numpy(xyz).to(dev), torch.from_numpy(offset).to(dev)
cfg = bench.LAYERS[0]
layer = layers.BasicLayer(bench.DS_SCALE, cfg["depth"], cfg["C"], cfg["h"], cfg["window"], 0.04, cfg["quant"], rel_query=True, rel_key=True,
                          rel_value=True, downsample=layers.TransitionDown, ratio=0.25, k=16, out_channels=96).to(dev)
feats = torch.randn(xd.shape[0], cfg["C"], device=dev)


def step():
    layer.zero_grad(set_to_none=True)
    f = feats.clone().requires_grad_(True)
    out = layer(f, xd, od)
    (out[0].square().mean() + out[3].square().mean()).backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
_cabi.profile_dump()
_cabi.profile_enable(True)
t0 = time.perf_counter()
n = 5
for _ in range(n):
    step()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / n * 1e3
_cabi.profile_enable(False)
prof = _cabi.profile_dump()
print(f"wall {wall:.2f} ms per fwd+bwd")
tot = 0.0
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
    print(f"  {k:40s} {v['ms'] / n:8.3f} ms  ({v['launches'] // n} launches)")
    tot += v["ms"] / n
print(f"  library kernels total {tot:.2f} ms")
