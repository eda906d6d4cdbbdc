"""CPU checks of the fused window-attention formulation (no GPU needed).

1. oracle/fused_plan_oracle.py (the spec of the device plan builder): the valid entries of its dense + sparse tiles are
   exactly the pair multiset and rel-pos indices of the reference's get_indice_pairs restatement (oracle/index_oracle.py,
   pinned to the reference's own Python by tests/golden).
2. stratified_transformer_b200/csrc/fused_phases.cuh — the body of the CUDA kernels — compiled for the host
   (tests/emu/fused_emu.cpp: every barrier-separated phase becomes a loop over the thread ids) reproduces the fp64
   oracle's forward output, log-sum-exp and all six gradients within the fp32 tolerance of the GPU parity tests.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import attention_oracle as ao, fps_oracle, fused_plan_oracle as fpo, index_oracle as io
from stratified_transformer_b200.synthetic import make_scene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_SRC = os.path.join(ROOT, "tests", "emu", "fused_emu.cpp")
EMU_LIB = os.path.join(ROOT, "tests", "emu", "_build", "libfw_emu.so")
CUDA_INC = os.environ.get("CUDA_HOME", "/usr/local/cuda") + "/include"


class PassParams(ctypes.Structure):
    _fields_ = [("items", ctypes.c_void_p), ("n_items", ctypes.c_int), ("q_order", ctypes.c_void_p), ("k_order", ctypes.c_void_p),
                ("rel", ctypes.c_void_p), ("pos_win", ctypes.c_void_p), ("wstart", ctypes.c_void_p), ("tile_base", ctypes.c_void_p),
                ("bin_lo", ctypes.c_int), ("RB", ctypes.c_int), ("Rpad", ctypes.c_int), ("L", ctypes.c_int), ("h", ctypes.c_int),
                ("q", ctypes.c_void_p), ("k", ctypes.c_void_p), ("v", ctypes.c_void_p),
                ("tq", ctypes.c_void_p), ("tk", ctypes.c_void_p), ("tv", ctypes.c_void_p),
                ("out", ctypes.c_void_p), ("m", ctypes.c_void_p), ("l", ctypes.c_void_p),
                ("g", ctypes.c_void_p), ("lse", ctypes.c_void_p),
                ("gq", ctypes.c_void_p), ("gk", ctypes.c_void_p), ("gv", ctypes.c_void_p),
                ("gtq", ctypes.c_void_p), ("gtk", ctypes.c_void_p), ("gtv", ctypes.c_void_p), ("dbg", ctypes.c_int)]


@pytest.fixture(scope="module")
def emu():
    if not os.path.isdir(CUDA_INC):
        pytest.skip("CUDA headers not available")
    os.makedirs(os.path.dirname(EMU_LIB), exist_ok=True)
    hdrs = [os.path.join(ROOT, "stratified_transformer_b200", "csrc", n) for n in ("fused_phases.cuh", "fused_tc.cuh", "tc_umma.cuh")]
    if not os.path.exists(EMU_LIB) or os.path.getmtime(EMU_LIB) < max([os.path.getmtime(EMU_SRC)] + [os.path.getmtime(x) for x in hdrs]):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", CUDA_INC, EMU_SRC, "-o", EMU_LIB])
    lib = ctypes.CDLL(EMU_LIB)
    lib.fw_emu_run.argtypes = [ctypes.POINTER(PassParams), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    lib.fw_emu_run_tc.argtypes = [ctypes.POINTER(PassParams), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    return lib


def ptr(a):
    return None if a is None else a.ctypes.data


def bin_range(L, dense, swin=False):
    """staged bin range [lo, lo+RB) and padded product width (multiple of 8)"""
    if swin or not dense:
        lo, RB = 0, L
    else:
        lo, RB = max(L // 4 - 2, 0), (L + 1) // 2 + 4
    return lo, RB, (3 * RB + 15) // 16 * 16


def run_pass(lib, part, ord_idx, backward, arrays, h, L, dense, swin=False, n_cta=3, tc=False):
    off = int(part["counts"][:ord_idx].sum())
    cnt = int(part["counts"][ord_idx])
    if cnt == 0:
        return
    items = np.ascontiguousarray(part["items"][off:off + cnt])
    lo, RB, Rpad = bin_range(L, dense, swin)
    P = PassParams()
    P.items, P.n_items = ptr(items), cnt
    P.q_order, P.k_order, P.rel = ptr(part["q_order"]), ptr(part["k_order"]), ptr(part["rel"])
    P.pos_win, P.wstart, P.tile_base = ptr(part.get("pos_win")), ptr(part.get("wstart")), ptr(part.get("tile_base"))
    P.bin_lo, P.RB, P.Rpad, P.L, P.h = lo, RB, Rpad, L, h
    for name in ("q", "k", "v", "tq", "tk", "tv", "out", "m", "l", "g", "lse", "gq", "gk", "gv", "gtq", "gtk", "gtv"):
        setattr(P, name, ptr(arrays.get(name)))
    fn = lib.fw_emu_run_tc if tc else lib.fw_emu_run
    assert fn(ctypes.byref(P), part["BQ"], part["BK"], int(backward), n_cta) == 0


def fused_forward_backward(lib, plan, q, k, v, tq, tk, tv, g, swin=False, tc=False):
    N, h, _ = q.shape
    L = tq.shape[0]
    A = dict(q=q, k=k, v=v, tq=tq, tk=tk, tv=tv, g=g)
    A["out"] = np.full((N, h, 16), np.nan, np.float32)
    A["m"] = np.full((N, h), np.nan, np.float32)
    A["l"] = np.full((N, h), np.nan, np.float32)
    for o in range(fpo.MAXORD):
        run_pass(lib, plan["dense"], o, False, A, h, L, True, swin, tc=tc)
    if plan["sparse"] is not None:
        for o in range(fpo.MAXORD):
            run_pass(lib, plan["sparse"], o, False, A, h, L, False, tc=tc)
    A["lse"] = A["m"]
    chunked = plan["dense"]["max_win"] > plan["dense"]["BK"]
    fill = 0.0 if chunked else np.nan     # key rows of chunked windows are accumulated into: the caller zero-fills
    A["gq"] = np.full((N, h, 16), np.nan, np.float32)
    A["gk"] = np.full((N, h, 16), fill, np.float32)
    A["gv"] = np.full((N, h, 16), fill, np.float32)
    for name, t in (("gtq", tq), ("gtk", tk), ("gtv", tv)):
        A[name] = np.zeros_like(t)
    for o in range(fpo.MAXORD):
        run_pass(lib, plan["dense"], o, True, A, h, L, True, swin, tc=tc)
    if plan["sparse"] is not None:
        for o in range(fpo.MAXORD):
            run_pass(lib, plan["sparse"], o, True, A, h, L, False, tc=tc)
    return A


def small_case(n_pts, seed, window, quant, lattice=False, scenes=2):
    xs = [make_scene(seed + s, n_pts, n_raw=60000, lattice=lattice)[0] for s in range(scenes)]
    xyz = np.concatenate(xs)
    offset = np.cumsum([x.shape[0] for x in xs]).astype(np.int32)
    ds = fps_oracle.furthestsampling(xyz, offset, io.fps_new_offset(offset, 8))
    return xyz, offset, ds


@pytest.mark.parametrize("parity", [0, 1])
@pytest.mark.parametrize("lattice", [False, True])
def test_plan_reproduces_reference_pairs(parity, lattice):
    xyz, offset, ds = small_case(900, 3, 0.16, 0.01, lattice)
    window, quant = 0.32, 0.02
    plan = fpo.build(xyz, offset, window, quant, parity, ds, BQ=16, BK=16, BQS=16, BKS=8)
    ref = io.build_layer_index(xyz, offset, window, 8, ds, parity)
    rel_ref = io.rel_pos_index_stratified(xyz, ref["index_0"], ref["index_1"], window, quant)
    i0, i1, rel = fpo.pairs_from_plan(plan)
    N = xyz.shape[0]

    def canon(a0, a1, r):
        key = np.lexsort((r[:, 2], r[:, 1], r[:, 0], a1, a0))
        return a0[key], a1[key], r[key]
    g0, g1, gr = canon(i0, i1, rel)
    w0, w1, wr = canon(ref["index_0"].astype(np.int64), ref["index_1"].astype(np.int64), rel_ref)
    assert g0.shape == w0.shape
    assert np.array_equal(g0, w0) and np.array_equal(g1, w1) and np.array_equal(gr, wr)
    assert plan["dense"]["counts"][1:].sum() > 0, "case must exercise chunked windows"


CASES = [
    # n_pts, window, quant, BQ, BK, BQS, BKS, h, lattice
    (700, 0.32, 0.02, 64, 64, 48, 32, 2, False),
    (700, 0.32, 0.02, 16, 16, 16, 8, 1, False),    # tiny blocks: chunked dense windows, several sparse key chunks
    (500, 0.32, 0.02, 32, 32, 32, 32, 3, True),   # lattice scene: duplicate keys (dense + sparse) for some queries
]


@pytest.mark.parametrize("tc", [False, True], ids=["fma", "tcgen05"])
@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("parity", [0, 1])
def test_emulated_kernels_match_oracle(emu, case, parity, tc):
    n_pts, window, quant, BQ, BK, BQS, BKS, h, lattice = case
    xyz, offset, ds = small_case(n_pts, 11, window, quant, lattice)
    plan = fpo.build(xyz, offset, window, quant, parity, ds, BQ=BQ, BK=BK, BQS=BQS, BKS=BKS)
    ref = io.build_layer_index(xyz, offset, window, 8, ds, parity)
    rel_ref = io.rel_pos_index_stratified(xyz, ref["index_0"], ref["index_1"], window, quant)
    N = xyz.shape[0]
    L = 2 * int((2 * window + 1e-4) // quant)
    g = torch.Generator().manual_seed(5)
    q, k, v, go = (torch.randn(N, h, 16, generator=g) for _ in range(4))
    q = q * 0.5
    tq, tk, tv = (torch.rand(L, h, 16, 3, generator=g) - 0.5 for _ in range(3))   # U(-.5,.5): strong table signal
    want = ao.layer_fwd_bwd(q.double(), k.double(), v.double(), torch.from_numpy(ref["offsets"]), torch.from_numpy(ref["index_1"]),
                            tq.double(), tk.double(), tv.double(), torch.from_numpy(rel_ref), go.double())
    A = fused_forward_backward(emu, plan, *(t.numpy().copy() for t in (q, k, v, tq, tk, tv, go)), tc=tc)
    # log-sum-exp of every row
    i0 = want["i0"]
    s = want["s"]
    mx = torch.full((N, h), -float("inf"), dtype=torch.float64).scatter_reduce(0, i0.unsqueeze(-1).expand(-1, h), s, "amax")
    lse = mx + torch.log(torch.zeros(N, h, dtype=torch.float64).index_add(0, i0, torch.exp(s - mx[i0])))
    got = dict(out=A["out"], lse=A["lse"], gq=A["gq"], gk=A["gk"], gv=A["gv"], gtq=A["gtq"], gtk=A["gtk"], gtv=A["gtv"])
    want = dict(want, lse=lse)
    for name, val in got.items():
        r = want[name].numpy()
        err = np.abs(val.astype(np.float64) - r)
        tol = (2e-4 if name.startswith("gt") else 1e-4) * np.maximum(1.0, np.abs(r))
        assert np.isfinite(val).all(), name
        assert (err <= tol).all(), f"{name}: max err {err.max():.3e} (|ref| max {np.abs(r).max():.2f})"


@pytest.mark.parametrize("tc", [False, True], ids=["fma", "tcgen05"])
def test_emulated_swin_dense_only(emu, tc):
    """3DSwin variant: dense windows only, tables of length 2*int(w/quant)-1, per-point quantised rel-pos index."""
    n_pts, window, quant, h = 600, 0.32, 0.02, 2
    xyz, offset, _ = small_case(n_pts, 21, window, quant)
    shift = 0.0
    plan = fpo.build(xyz, offset, window, quant, 0, None, BQ=32, BK=32, swin_shift=shift)
    assert plan["sparse"] is None
    i0, i1, rel = fpo.pairs_from_plan(plan)
    order = np.lexsort((i1, i0))
    i0, i1, rel = i0[order], i1[order], rel[order]
    N = xyz.shape[0]
    offsets = np.concatenate([[0], np.cumsum(np.bincount(i0, minlength=N))])
    L = 2 * int(window / quant) - 1
    g = torch.Generator().manual_seed(9)
    q, k, v, go = (torch.randn(N, h, 16, generator=g) for _ in range(4))
    tq, tk, tv = (torch.rand(L, h, 16, 3, generator=g) - 0.5 for _ in range(3))
    want = ao.layer_fwd_bwd(q.double(), k.double(), v.double(), torch.from_numpy(offsets), torch.from_numpy(i1), tq.double(),
                            tk.double(), tv.double(), torch.from_numpy(rel), go.double())
    A = fused_forward_backward(emu, plan, *(t.numpy().copy() for t in (q, k, v, tq, tk, tv, go)), swin=True, tc=tc)
    for name in ("out", "gq", "gk", "gv", "gtq", "gtk", "gtv"):
        r = want[name].numpy()
        err = np.abs(A[name].astype(np.float64) - r)
        tol = (2e-4 if name.startswith("gt") else 1e-4) * np.maximum(1.0, np.abs(r))
        assert (err <= tol).all(), f"{name}: max err {err.max():.3e}"
