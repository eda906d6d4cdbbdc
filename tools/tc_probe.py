#!/usr/bin/env python
"""Development tool: structured probes of stb200_tc_selftest (which rows of the accumulator are right, per variant)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stratified_transformer_b200 import _cabi

def run(mode, variant, M, N, K, A, B):
    out = torch.full((128, N), float("nan"), device="cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    Ad, Bd = A.cuda().contiguous(), B.cuda().contiguous()   # keep the device copies alive across the call
    _cabi.call("stb200_tc_selftest", mode | (variant << 8), M, N, K, Ad.data_ptr(), Bd.data_ptr(),
               out.data_ptr(), status.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return out.cpu().double(), int(status.item())

g = torch.Generator().manual_seed(0)
for mode in (0, 1, 2):
    for (M, N, K) in [(128, 16, 8), (128, 32, 8), (128, 64, 16), (128, 112, 16), (64, 32, 16)]:
        for variant in (0, 1, 2, 4, 6):
            a_shape = (K, M) if mode == 1 else (M, K)
            b_shape = (K, N) if mode in (1, 2) else (N, K)
            A = torch.randint(-8, 9, a_shape, generator=g).float(); B = torch.randint(-8, 9, b_shape, generator=g).float()
            Ad, Bd = A.double(), B.double()
            want = Ad @ Bd.T if mode == 0 else (Ad.T @ Bd if mode == 1 else Ad @ Bd)
            got, st = run(mode, variant, M, N, K, A, B)
            if M == 128:
                bad = ((got[:M] - want).abs().max(1).values > 1e-3).nonzero().flatten().tolist()
                desc = f"bad rows: {len(bad)} {bad[:8]}{'...' if len(bad) > 8 else ''}{bad[-3:] if len(bad) > 8 else ''}"
            else:
                lanes = []
                for r in range(M):
                    d = torch.nan_to_num((got - want[r]).abs().max(1).values, nan=1e9)
                    lanes.append(int(d.argmin()) if float(d.min()) < 1e-3 else -1)
                desc = f"row->lane {lanes}"
            print(f"mode {mode} M{M} N{N} K{K} variant {variant} status {st}: {desc}")
