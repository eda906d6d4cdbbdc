import os, sys, time, torch
sys.path.insert(0, os.getcwd())
from stratified_transformer_b200 import pointops
from stratified_transformer_b200.synthetic import make_batch
xyz, _, off = make_batch(1, 80000, seed0=7)
xd, od = torch.from_numpy(xyz).cuda(), torch.from_numpy(off).cuda()
for m in (10001, 20001):
    no = torch.tensor([m], dtype=torch.int32, device="cuda")
    pointops.furthestsampling(xd, od, no); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(3): pointops.furthestsampling(xd, od, no)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / 3
    print(f"threads={os.environ.get('STB200_FPS_THREADS','-')} cluster={os.environ.get('STB200_FPS_CLUSTER','-')} m={m}: {dt*1e3:.2f} ms {dt/m*1e6:.3f} us/iter")
