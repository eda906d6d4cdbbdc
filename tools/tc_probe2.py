#!/usr/bin/env python
"""Development tool: which element an MN-major operand descriptor makes the tensor core read (structured probe)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stratified_transformer_b200 import _cabi
torch.set_printoptions(linewidth=250, precision=0, sci_mode=False)

def run(mode, variant, M, N, K, A, B):
    out = torch.full((128, N), float("nan"), device="cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    Ad, Bd = A.cuda().contiguous(), B.cuda().contiguous()
    _cabi.call("stb200_tc_selftest", mode | (variant << 8), M, N, K, Ad.data_ptr(), Bd.data_ptr(),
               out.data_ptr(), status.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return out.cpu(), int(status.item())

for (M, N, K) in [(128, 16, 8), (128, 32, 16)]:
    A = torch.zeros(M, K); A[torch.arange(K), torch.arange(K)] = 1          # K-major A = unit rows
    B = (torch.arange(K).view(K, 1) * N + torch.arange(N).view(1, N)).float()  # B[k][n] = k*N + n, MN-major
    D, st = run(2, 1, M, N, K, A, B)
    print(f"== mode 2 (B MN-major) M{M} N{N} K{K}: want D[m][n] = {N}*m + n for m < {K}; status {st}")
    print(D[:K, :N])
    # A MN-major probe: A[k][m] = k*M + m ; B K-major identity rows: B[n][k] = delta(n,k) -> D[m][n] = A[n][m] for n < K
    A = (torch.arange(K).view(K, 1) * M + torch.arange(M).view(1, M)).float()
    Bk = torch.zeros(N, K); Bk[torch.arange(min(N, K)), torch.arange(min(N, K))] = 1
    # mode 3 does not exist: use mode 1 with an MN-major B that equals identity: B[k][n] = delta(k,n)
    Bm = torch.zeros(K, N); Bm[torch.arange(min(N, K)), torch.arange(min(N, K))] = 1
    D, st = run(1, 1, M, N, K, A, Bm)
    print(f"== mode 1 (both MN-major) want D[m][n] = {M}*n + m for n < {K} (if B is read right)")
    print(D[:12, :N]); print(D[60:64, :N])
