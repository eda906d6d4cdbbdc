"""torch-profiler kernel table of one whole-network forward + backward (model.Stratified, S3DIS configuration, 1 scene of 80k points)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from stratified_transformer_b200 import prestep
from stratified_transformer_b200.model import Stratified
from stratified_transformer_b200.synthetic import make_batch
from torch.profiler import profile, ProfilerActivity

dev = torch.device("cuda")
scenes = int(sys.argv[1]) if len(sys.argv) > 1 else 1
xyz, rgb, offset = make_batch(scenes, 80000, seed0=7)
xd, od = torch.from_numpy(xyz).to(dev), torch.from_numpy(offset).to(dev)
feat = torch.cat([torch.from_numpy(rgb).to(dev).float(), xd], 1)
batch = prestep.batch_from_offset(od)
nbr = prestep.ball_query(0.1, 34, xd, xd, batch_x=batch, batch_y=batch)[0]
L = bench.LAYERS
model = Stratified(8, [c["depth"] for c in L], [c["C"] for c in L], [c["h"] for c in L], [c["window"] for c in L], 3,
                   [0.04 * 2 ** i for i in range(4)], [c["quant"] for c in L], rel_query=True, rel_key=True, rel_value=True,
                   drop_path_rate=0.0, concat_xyz=True, stem_transformer=True).to(dev)


amp = len(sys.argv) > 2 and sys.argv[2] == "amp"


def step():
    model.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
        out, shift = model(feat, xd, od, batch, nbr)
    (out.float().square().mean() + shift.float().square().mean()).backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as tp:
    step()
    torch.cuda.synchronize()
print(tp.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70))
