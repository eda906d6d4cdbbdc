"""Device versions of the per-step host work of the reference's training loop (/root/reference/train.py:319-325): the per-point
scene id and the radius neighbour search that feeds the KPConv stem.  Same call forms as the statements they replace."""
from __future__ import annotations

import torch

from . import _cabi


def _stream():
    return torch.cuda.current_stream().cuda_stream


def batch_from_offset(offset: torch.Tensor, n_points: int | None = None) -> torch.Tensor:
    """`torch.cat([torch.tensor([ii] * o) for ii, o in enumerate(offset_)], 0).long()` (train.py:319-321) without the Python lists:
    offset = cumulative scene ends [b] (any integer dtype, on the GPU); returns int64 [N] on the same device."""
    if not offset.is_cuda:
        raise _cabi.Stb200Error("batch_from_offset: offset must be a CUDA tensor (there is no CPU path)")
    off = offset.to(torch.int32).contiguous()
    N = int(off[-1]) if n_points is None else int(n_points)
    batch = torch.empty(N, dtype=torch.int64, device=off.device)
    _cabi.call("stb200_batch_from_offset", N, off.numel(), off.data_ptr(), batch.data_ptr(), _stream())
    return batch


def ball_query(radius, nsample, x, y, mode="partial_dense", batch_x=None, batch_y=None, sort=True):
    """torch_points_kernels.ball_query(radius, nsample, x, y, mode="partial_dense", batch_x=..., batch_y=...): x = support points
    [Nx, 3], y = query points [Ny, 3]; returns (idx int64 [Ny, nsample] into x, -1 padded; dist2 float [Ny, nsample], -1 padded).
    The matches of a query are ordered by (distance, index) and the closest `nsample` are kept (the reference library's order
    and, beyond `nsample` matches, its choice are unspecified: include/stb200.h)."""
    if mode.lower() != "partial_dense":
        raise ValueError("only mode='partial_dense' (the one the reference uses, train.py:325) is implemented")
    if not (x.is_cuda and y.is_cuda):
        raise _cabi.Stb200Error("ball_query: x and y must be CUDA tensors (there is no CPU path)")
    if (batch_x is None) != (batch_y is None):
        raise ValueError("batch_x and batch_y: give both or neither")
    x = x.float().contiguous()
    y = y.float().contiguous()
    Nx, Ny = x.shape[0], y.shape[0]
    bx = None if batch_x is None else batch_x.to(device=x.device, dtype=torch.int64).contiguous()
    by = None if batch_y is None else batch_y.to(device=x.device, dtype=torch.int64).contiguous()
    idx = torch.empty(Ny, int(nsample), dtype=torch.int64, device=x.device)
    dist2 = torch.empty(Ny, int(nsample), dtype=torch.float32, device=x.device)
    nbytes = int(_cabi.load().stb200_ball_query_workspace_bytes(Nx))
    ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=x.device)
    _cabi.call("stb200_ball_query", Nx, Ny, float(radius), int(nsample), x.data_ptr(), y.data_ptr(),
               None if bx is None else bx.data_ptr(), None if by is None else by.data_ptr(), ws.data_ptr(), nbytes,
               idx.data_ptr(), dist2.data_ptr(), _stream())
    return idx, dist2
