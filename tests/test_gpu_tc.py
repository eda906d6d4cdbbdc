"""tcgen05 building blocks (-m gpu): 3xTF32 GEMMs through tensor memory against an fp64 product for the operand form the
fused kernels use (csrc/tc_umma.cuh): K-major un-swizzled chunked matrices, M = 128 and 64 (and the TMEM row -> lane map of
an M = 64 accumulator).  MN-major tf32 operands exist only with the 128B_BASE32B swizzle (CUTLASS: "for mn-major tf32
operands, SW128_32B is the only available smem layout"); the un-swizzled MN-major modes of the self-test (1, 2) return
zeros on hardware and are not used by any kernel — the fused kernels keep every operand K-major instead."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def run(mode, M, N, K, seed=0):
    from stratified_transformer_b200 import _cabi
    g = torch.Generator().manual_seed(seed)
    a_shape = (K, M) if mode == 1 else (M, K)
    b_shape = (K, N) if mode in (1, 2) else (N, K)
    A = torch.randn(*a_shape, generator=g)
    B = torch.randn(*b_shape, generator=g)
    Ad, Bd = A.double(), B.double()
    want = Ad @ Bd.T if mode == 0 else (Ad.T @ Bd if mode == 1 else Ad @ Bd)
    out = torch.full((128, N), float("nan"), device="cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    A_dev, B_dev = A.cuda().contiguous(), B.cuda().contiguous()   # keep the device copies alive across the call
    _cabi.call("stb200_tc_selftest", mode, M, N, K, A_dev.data_ptr(), B_dev.data_ptr(), out.data_ptr(), status.data_ptr(),
               torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert int(status.item()) == 0, "MMA never completed (mbarrier timeout)"
    return out.cpu().double(), want


@pytest.mark.parametrize("mode", [0])
@pytest.mark.parametrize("shape", [(128, 16, 32), (128, 16, 8), (128, 112, 16), (128, 16, 112), (128, 64, 64), (128, 192, 16), (128, 16, 64), (128, 240, 16)])
def test_gemm_3xtf32_m128(mode, shape):
    M, N, K = shape
    got, want = run(mode, M, N, K)
    err = (got[:M] - want).abs().max().item()
    scale = want.abs().max().item()
    print(f"mode {mode} {shape}: max err {err:.3e} (scale {scale:.1f})")
    assert err <= 2e-6 * max(scale, 1.0) * max(1.0, K / 16), f"mode {mode} {shape}: max err {err:.3e}"


@pytest.mark.parametrize("mode", [0])
def test_m64_accumulator_layout(mode):
    """where the 64 rows of an M = 64 accumulator live in TMEM (documented in tc_umma.cuh)"""
    M, N, K = 64, 32, 16
    got, want = run(mode, M, N, K)
    lanes = []
    for r in range(M):
        d = (got - want[r]).abs().max(1).values
        d = torch.nan_to_num(d, nan=1e9)
        lanes.append(int(d.argmin()))
        assert float(d.min()) < 1e-4, f"row {r} not found in TMEM"
    print("M=64 row -> lane:", lanes)
    assert lanes == [32 * (r // 16) + r % 16 for r in range(64)], lanes
