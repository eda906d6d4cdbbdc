import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import test_gpu_fused as T
from stratified_transformer_b200 import index, pointops
tag = sys.argv[1]
xyz, offset, ds = T._small_scene(2500, 11, False)
window, quant, parity, h = 0.32, 0.02, 0, 3
xd, od = torch.from_numpy(xyz).cuda(), torch.from_numpy(offset).cuda()
pi = index.build_stratified_index(xd, od, window, quant, torch.from_numpy(ds).cuda(), parity, fused=True)
N = xyz.shape[0]; L = 64
g = torch.Generator().manual_seed(5)
q, k, v, go = (torch.randn(N, h, 16, generator=g) for _ in range(4)); q = q * 0.5
tq, tk, tv = ((torch.rand(L, h, 16, 3, generator=g) - 0.5) for _ in range(3))
leaves = [t.cuda().requires_grad_(True) for t in (q, k, v, tq, tk, tv)]
out = pointops.window_attention_plan(*leaves, pi.plan); out.backward(go.cuda())
torch.save({n: t.grad.cpu() for n, t in zip("q k v tq tk tv".split(), leaves)}, f"gpurun_out/dbg_{tag}.pt")
print(tag, "done", pi.plan.totals[:24])
