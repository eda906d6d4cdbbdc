import torch
torch.manual_seed(0)
n = 20_000_000
for q in (0.01, 0.02, 0.04, 0.08, 0.005):
    t = (torch.rand(n) * (64 * q)).float()
    g = (t.cuda() // q).cpu(); c = t // q
    inv = torch.tensor(1.0, dtype=torch.float32) / torch.tensor(q, dtype=torch.float32)
    mod = torch.fmod(t, q)
    div = (t - mod) * inv
    fl = torch.floor(div); fl = torch.where((div - fl) > 0.5, fl + 1, fl)
    cand = torch.where(div != 0, fl, torch.copysign(torch.zeros_like(t), t * inv))
    print(f"q={q}: gpu!=cpu {int((g != c).sum())}  gpu!=cand(inv_b) {int((g != cand).sum())}")
x = torch.round((torch.rand(n) * 2 - 1) * 1.3 * 100000)
g = (x.cuda() / 100000).cpu(); c = x / 100000
cand = x * (torch.tensor(1.0, dtype=torch.float32) / torch.tensor(100000.0, dtype=torch.float32))
print("div by 1e5: gpu!=cpu", int((g != c).sum()), "gpu!=x*(1/1e5 in fp32)", int((g != cand).sum()), "gpu!=x*float(1e-5 from double)", int((g != x * torch.tensor(1e-5, dtype=torch.float32)).sum()))
# remainder by a python scalar (Swin: % window) and the add / sub of python scalars
w = 0.16
y = (torch.rand(n) * 10).float()
print("remainder: gpu!=cpu", int(((y.cuda() % w).cpu() != (y % w)).sum()))
print("add/sub scalars: gpu!=cpu", int((((y.cuda() + 2 * w) - 0.0001).cpu() != ((y + 2 * w) - 0.0001)).sum()))
print("mul 1e5: gpu!=cpu", int(((y.cuda() * 100000).cpu() != (y * 100000)).sum()))
# tensor // tensor (window_coord in get_indice_pairs uses a tensor divisor)
wt = torch.tensor([w, w, w])
z = (torch.rand(n // 3, 3) * 10).float()
print("tensor//tensor: gpu!=cpu", int(((z.cuda() // wt.cuda()).cpu() != (z // wt)).sum()))
