import torch
a, b = torch.load("gpurun_out/dbg_tc.pt"), torch.load("gpurun_out/dbg_fma.pt")
for n in a:
    print(n, float((a[n] - b[n]).abs().max()))
d = (a["tk"] - b["tk"]).abs()          # [L, h, 16, 3]
print("per bin max err:", [round(float(x), 2) for x in d.amax((1, 2, 3))])
print("per axis:", d.amax((0, 1, 2)).tolist(), "per head:", d.amax((0, 2, 3)).tolist(), "per channel:", [round(float(x),2) for x in d.amax((0, 1, 3))])
