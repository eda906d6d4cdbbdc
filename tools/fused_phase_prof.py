#!/usr/bin/env python
"""Development tool: per-phase cycle counts of one CTA of the tcgen05 fused kernels (clock stamps between the barrier-separated
phases), one synthetic 80k-point scene, layer 0.  Prints cycles per item for every phase of the four launches."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stratified_transformer_b200 import _cabi, index
from stratified_transformer_b200.synthetic import make_batch
lib = _cabi.load()
lib.stb200_fused_phase_profile.argtypes = [ctypes.c_void_p]; lib.stb200_fused_phase_profile.restype = None
xyz, _, offset = make_batch(1, 80000, seed0=0)
xd, od = torch.from_numpy(xyz).cuda(), torch.from_numpy(offset).cuda()
li = index.build_layer_index(xd, od, 0.16, 0.01, 8, fused=True, csr=False)
plan = li.for_block(1).plan
N, h, L = xd.shape[0], 3, 64
g = torch.Generator(device="cuda").manual_seed(0)
q, k, v, go = (torch.randn(N, h, 16, device="cuda", generator=g) for _ in range(4))
tq, tk, tv = (torch.randn(L, h, 16, 3, device="cuda", generator=g) * 0.02 for _ in range(3))
out = torch.empty(N, h, 16, device="cuda"); lse = torch.empty(N, h, device="cuda"); lsum = torch.empty(N, h, device="cuda")
gq, gk, gv = (torch.zeros(N, h, 16, device="cuda") for _ in range(3)); gt = [torch.zeros_like(tq) for _ in range(3)]
passes, n_passes = plan.passes(L)
stream = torch.cuda.current_stream().cuda_stream
buf = torch.zeros(64, dtype=torch.int64, device="cuda")
names_f = ["describe", "stage", "issue", "wait", "copy-out", "logits+max", "rowmax+zero hist", "exp+sum", "hist", "PV+hist*T", "merge+store"]
names_b = ["describe", "stage", "issue1", "wait1", "copy1", "issue KT", "wait", "copy KT", "issue GT + E1", "wait", "copy GT", "E2", "zero hists",
           "build hists", "FMA GEMMs", "split+store rows", "lo write-back", "issue G5", "wait G5", "flush"]
for i in range(n_passes):
    one = (_cabi.FusedPass * 1)(passes[i])
    items_per_cta = -(-passes[i].n_items // (148 // h))
    for bwd in (0, 1):
        for rep in range(2):
            buf.zero_()
            lib.stb200_fused_phase_profile(buf.data_ptr() if rep else None)
            if bwd:
                _cabi.call("stb200_fused_attention_backward", one, 1, N, h, L, go.data_ptr(), out.data_ptr(), lse.data_ptr(), q.data_ptr(), k.data_ptr(),
                           v.data_ptr(), tq.data_ptr(), tk.data_ptr(), tv.data_ptr(), gq.data_ptr(), gk.data_ptr(), gv.data_ptr(), gt[0].data_ptr(),
                           gt[1].data_ptr(), gt[2].data_ptr(), stream)
            else:
                _cabi.call("stb200_fused_attention_forward", one, 1, N, h, L, q.data_ptr(), k.data_ptr(), v.data_ptr(), tq.data_ptr(), tk.data_ptr(),
                           tv.data_ptr(), out.data_ptr(), lse.data_ptr(), lsum.data_ptr(), stream)
            torch.cuda.synchronize()
        c = buf.tolist()
        names = names_b if bwd else names_f
        tot = sum(c[:len(names)])
        print(f"== pass {i} ({'dense' if passes[i].pos_win else 'sparse'}, {passes[i].n_items} items, ~{items_per_cta} per CTA) {'backward' if bwd else 'forward'}: {tot / items_per_cta:.0f} cycles per item")
        for n_, x in zip(names, c):
            print(f"   {n_:<18} {x / items_per_cta:8.0f}  {100 * x / max(tot, 1):5.1f}%")
lib.stb200_fused_phase_profile(None)
