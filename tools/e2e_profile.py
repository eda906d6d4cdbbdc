"""Development aid: torch.profiler breakdown of one e2e step of bench.py's HotPathModel (serial geometry)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from torch.profiler import profile, ProfilerActivity

class A: scenes=8; points=80000
dev = torch.device("cuda")
levels, rgb = bench.build_inputs(A, 0, dev)
model = bench.HotPathModel().to(dev)
feat6 = torch.cat([rgb, levels[0]["xyz"].cpu()], 1).to(dev)
xyzs = [lv["xyz"] for lv in levels]; offs = [lv["offset"] for lv in levels]; subs = [None] + [lv["sub_idx"] for lv in levels[1:]]
def step():
    model.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):   # as bench.py's e2e leg
        loss = model(feat6, xyzs, offs, subs)
    loss.backward(); return float(loss.item())
for _ in range(2): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
