import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import fps_oracle, index_oracle as io
from stratified_transformer_b200 import index
from stratified_transformer_b200.synthetic import make_batch
xyz, _, offset = make_batch(2, 3000, seed0=41, n_raw=80000)
ds = fps_oracle.furthestsampling(xyz, offset, io.fps_new_offset(offset, 8))
xd, od = torch.from_numpy(xyz).cuda(), torch.from_numpy(offset).cuda()
window, quant = 0.32, 0.02
pi = index.build_stratified_index(xd, od, window, quant, torch.from_numpy(ds).cuda(), 0, want_index_0=True)
i0, i1 = pi.index_0.long(), pi.index_1.long()
def three(x, a, b):
    r = x[a] - x[b]
    r = torch.round(r * 100000) / 100000
    return ((r + 2 * window - 0.0001) // quant)
gpu = three(xd, i0, i1).int()
cpu = three(xd.cpu(), i0.cpu(), i1.cpu()).int()
mine = pi.rel_idx
print("M", pi.M, "gpu!=cpu", int((gpu.cpu() != cpu).sum()), "mine!=cpu", int((mine.cpu() != cpu).sum()), "mine!=gpu", int((mine != gpu).sum()))
bad = (gpu.cpu() != cpu).any(1).nonzero().flatten()[:5]
for m in bad.tolist():
    a, b = int(i0[m]), int(i1[m])
    r = (xd[a] - xd[b]).cpu()
    print(m, r.tolist(), "gpu", gpu[m].tolist(), "cpu", cpu[m].tolist())
    r2 = torch.round(r * 100000) / 100000
    t = r2 + 2 * window - 0.0001
    print("   t", [f"{float(x):.9f}" for x in t], "t/q", [f"{float(x)/quant:.6f}" for x in t], "gpu //:", (t.cuda() // quant).tolist(), "cpu //:", (t // quant).tolist())
    rg = torch.round((xd[a] - xd[b]) * 100000) / 100000
    print("   round-div gpu", [f"{float(x):.9f}" for x in rg.cpu()], "cpu", [f"{float(x):.9f}" for x in r2])
