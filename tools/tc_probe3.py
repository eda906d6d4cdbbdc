import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stratified_transformer_b200 import _cabi
def run(mode, variant, M, N, K, A, B):
    out = torch.full((128, N), float("nan"), device="cuda"); status = torch.zeros(1, dtype=torch.int32, device="cuda")
    Ad, Bd = A.cuda().contiguous(), B.cuda().contiguous()
    _cabi.call("stb200_tc_selftest", mode | (variant << 8), M, N, K, Ad.data_ptr(), Bd.data_ptr(), out.data_ptr(), status.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize(); return out.cpu().double(), int(status.item())
g = torch.Generator().manual_seed(0)
for (M, N, K) in [(128,16,8),(128,16,16),(128,16,24),(128,16,32),(128,16,40),(128,16,48),(128,16,64),(64,16,32),(64,16,24),(128,32,32),(128,64,32)]:
    A = torch.randn(M, K, generator=g); B = torch.randn(N, K, generator=g)
    want = A.double() @ B.double().T
    got, st = run(0, 0, M, N, K, A, B)
    if M == 128: err = (got[:M] - want).abs().max(1).values
    else:
        lanes = [32*(r//16) + r%16 for r in range(M)]; err = (got[lanes] - want).abs().max(1).values
    bad = (err > 1e-4).nonzero().flatten().tolist()
    print(f"M{M} N{N} K{K} (ro {K//4*128}) status {st}: max err {float(err.max()):.2e} bad rows {len(bad)} {bad[:10]}")
