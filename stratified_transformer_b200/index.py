"""Pair-index construction on the device (host-side mirror of the reference's Python).

Replaces, for the attention hot path, `grid_sample` + `get_indice_pairs` + the sort/bincount/cumsum of
`BasicLayer.forward` (/root/reference/model/stratified_transformer.py:10-65, 267-317) and the relative
position index of `WindowAttention.forward` (:186-188; Swin: model/swin3d_transformer.py:151-154).

The reference rebuilds the pair list for every block although only two distinct lists exist per layer
(even blocks: unshifted windows, odd blocks: shifted).  `LayerIndex` builds both once per layer.
"""
from __future__ import annotations

import os
import warnings
from dataclasses import dataclass, field

import torch

from . import _cabi
from . import pointops
from . import pointops2_cuda as ext


def _stream():
    return torch.cuda.current_stream().cuda_stream


def set_torch_semantics(device: str = "cuda") -> None:
    """Which torch device's arithmetic the Stratified rel-pos index reproduces bit for bit: "cuda" (default; the reference as it
    runs: `x / 100000` is a reciprocal multiply on CUDA tensors) or "cpu" (IEEE division; for fixtures made with CPU torch)."""
    if device not in ("cuda", "cpu"):
        raise ValueError("device must be 'cuda' or 'cpu'")
    _cabi.call("stb200_set_torch_semantics", 1 if device == "cuda" else 0)


FUSED_BLOCKS = dict(BQ=64, BK=64, BQS=48, BKS=32)   # block shapes the fused kernels are built for (dense square, sparse 48x32)
_PLAN_MAXORD = 8


@dataclass
class FusedPlan:
    """Work plan of the window-centric fused kernels (include/stb200.h, "Window-centric fused attention"): the pair structure of
    one block parity as dense tiles per small window + sparse tiles per large window + block items.  Built by the device
    builder from the same window partition as the CSR pair list; replaces it on the fused path."""
    N: int
    totals: list                         # host copy of the 40 plan totals
    dense_items: torch.Tensor            # [n,8] int32, grouped by key-chunk ordinal
    dense_rel: torch.Tensor
    tile_base: torch.Tensor
    pos_win: torch.Tensor
    order_s: torch.Tensor
    wstart_s: torch.Tensor
    sparse_items: torch.Tensor | None
    sparse_rel: torch.Tensor | None
    order_l: torch.Tensor | None
    samp: torch.Tensor | None
    blocks: dict
    swin: bool = False
    _passes: dict = field(default_factory=dict, repr=False)

    @property
    def max_window(self) -> int:
        return self.totals[3]

    @property
    def needs_zeroed_key_grads(self) -> bool:
        """windows larger than a block accumulate their key-side gradients from several items"""
        return self.totals[3] > self.blocks["BK"]

    def tensors(self):
        return [t for t in (self.dense_items, self.dense_rel, self.tile_base, self.pos_win, self.order_s, self.wstart_s,
                            self.sparse_items, self.sparse_rel, self.order_l, self.samp) if t is not None]

    def bin_range(self, L: int, dense: bool):
        """bins [lo, lo+RB) staged by a pass.  Dense pairs share a small window, |delta| < w: their bins lie in the middle
        half of the table (model/stratified_transformer.py:186-188 with |r| < window_size); two bins of margin each side."""
        if self.swin or not dense:
            return 0, L
        return max(L // 4 - 2, 0), (L + 1) // 2 + 4

    def passes(self, L: int):
        """(ctypes array of stb200_fused_pass in launch order, count) for table length L"""
        ent = self._passes.get(L)
        if ent is None:
            lst = []
            lo, RB = self.bin_range(L, True)
            off = 0
            for o in range(_PLAN_MAXORD):
                n = self.totals[8 + o]
                if n:
                    lst.append(_cabi.FusedPass(self.dense_items.data_ptr() + off * 32, n, self.order_s.data_ptr(), self.order_s.data_ptr(),
                                               self.dense_rel.data_ptr(), self.pos_win.data_ptr(), self.wstart_s.data_ptr(),
                                               self.tile_base.data_ptr(), lo, RB, self.blocks["BQ"], self.blocks["BK"]))
                off += n
            if self.sparse_items is not None:
                lo, RB = self.bin_range(L, False)
                off = 0
                for o in range(_PLAN_MAXORD):
                    n = self.totals[16 + o]
                    if n:
                        lst.append(_cabi.FusedPass(self.sparse_items.data_ptr() + off * 32, n, self.order_l.data_ptr(), self.samp.data_ptr(),
                                                   self.sparse_rel.data_ptr(), None, None, None, lo, RB, self.blocks["BQS"], self.blocks["BKS"]))
                    off += n
            ent = ((_cabi.FusedPass * len(lst))(*lst), len(lst))
            self._passes[L] = ent
        return ent


def _plan_count(N, has_sparse, ws, blocks):
    """enqueue the plan's counting pass on the builder workspace of a finished stb200_stratified_pairs_count"""
    lib = _cabi.load()
    nbytes = lib.stb200_fused_plan_scratch_bytes(N)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=ws.device)
    totals = torch.empty(40, dtype=torch.int32, device=ws.device)
    _cabi.call("stb200_fused_plan_count", N, ws.data_ptr(), ws.numel(), int(has_sparse), blocks["BQ"], blocks["BK"], blocks["BQS"],
               blocks["BKS"], scratch.data_ptr(), nbytes, totals.data_ptr(), _stream())
    return scratch, totals


def _plan_fill(N, xyz, window_size, quant_size, n_win, n_samp, ws, scratch, totals_dev, totals, blocks, swin_shift=None):
    """FusedPlan, or None when the window structure is outside what the fused kernels take (a window that needs more than
    _PLAN_MAXORD key chunks, i.e. more than 512 points in one small window / sampled keys in one large window, or more than
    2^31 tile words): the caller then builds the CSR pair list and the per-op kernels run - another CUDA path, not a CPU one."""
    if totals[5]:
        warnings.warn(f"fused plan not applicable to this geometry (error bits {totals[5]}: a window needs more than "
                      f"{_PLAN_MAXORD} key chunks, or more than 2^31 tile words); using the per-op pair-list kernels", stacklevel=3)
        return None
    dev = xyz.device
    has_sparse = n_samp > 0

    def i32(n):
        return torch.empty(max(int(n), 1), dtype=torch.int32, device=dev)
    n_dense_items = sum(totals[8:16])
    n_sparse_items = sum(totals[16:24])
    dense_items = torch.empty(max(n_dense_items, 1), 8, dtype=torch.int32, device=dev)
    dense_rel, tile_base, pos_win, order_s, wstart_s = i32(totals[0]), i32(n_win), i32(N), i32(N), i32(n_win + 1)
    sparse_items = sparse_rel = order_l = samp = None
    if has_sparse:
        sparse_items = torch.empty(max(n_sparse_items, 1), 8, dtype=torch.int32, device=dev)
        sparse_rel, order_l, samp = i32(totals[1]), i32(N), i32(n_samp)
    swin = swin_shift is not None
    _cabi.call("stb200_fused_plan_fill", N, xyz.data_ptr(), float(2 * window_size), float(quant_size if quant_size is not None else 1.0),
               int(has_sparse), blocks["BQ"], blocks["BK"], blocks["BQS"], blocks["BKS"], int(swin), float(window_size),
               float(swin_shift or 0.0), ws.data_ptr(), ws.numel(), scratch.data_ptr(), scratch.numel(), totals_dev.data_ptr(),
               dense_rel.data_ptr(), tile_base.data_ptr(), pos_win.data_ptr(), order_s.data_ptr(), wstart_s.data_ptr(),
               dense_items.data_ptr(), None if sparse_rel is None else sparse_rel.data_ptr(),
               None if order_l is None else order_l.data_ptr(), None if samp is None else samp.data_ptr(), int(n_samp),
               None if sparse_items is None else sparse_items.data_ptr(), _stream())
    plan = FusedPlan(N, list(totals), dense_items, dense_rel, tile_base, pos_win, order_s, wstart_s, sparse_items, sparse_rel,
                     order_l, samp, dict(blocks), swin)
    plan._totals_dev = totals_dev
    return plan



@dataclass
class PairIndex:
    """CSR pair list of one block parity.  All index tensors are int32 on the device."""
    index_0_offsets: torch.Tensor          # [N+1]
    index_1: torch.Tensor                  # [M]   keys; per query: dense ascending id, then sparse ascending id
    rel_idx: torch.Tensor | None           # [M,3] relative-position index
    n_max: int
    M: int
    index_0: torch.Tensor | None = None    # [M]   only when asked for (v1 ops, scatter_softmax callers)
    row_order: torch.Tensor | None = None  # [N]   points sorted by window (locality hint for the fused entry points)
    win_offsets: torch.Tensor | None = None  # [n_win+1] window boundaries inside row_order
    n_win: int = 0
    plan: FusedPlan | None = None           # work plan of the fused kernels (build_stratified_index(..., fused=True))
    _fused: tuple | None = field(default=None, repr=False)   # (flags u8 [n_win], fallback_rows i32 [count])
    _tcsr: ext.TransposedCSR | None = field(default=None, repr=False)
    _packed: dict = field(default_factory=dict, repr=False)   # L -> (rel_packed, t_rel_packed | None)
    _len_orders: tuple | None = field(default=None, repr=False)   # (queries by pair count, keys by incoming pair count)

    @property
    def N(self) -> int:
        return self.index_0_offsets.numel() - 1 if self.index_0_offsets is not None else self.plan.N

    @property
    def tcsr(self) -> ext.TransposedCSR:
        """pairs grouped by key; built on first use (backward only)"""
        if self._tcsr is None:
            self._tcsr = ext.build_transposed_csr(self.index_0_offsets, self.index_1)
        return self._tcsr

    def fused_plan(self):
        """(window flags, fallback row list) for the fused forward kernel: which windows have one shared key list that
        fits its tile, and the rows of all the others.  Computed once per pair list (one small host read)."""
        if self._fused is None:
            if self.row_order is None or self.win_offsets is None or self.n_win == 0:
                return None
            dev = self.index_1.device
            flags = torch.empty(self.n_win, dtype=torch.uint8, device=dev)
            rows = torch.empty(self.N, dtype=torch.int32, device=dev)
            count = torch.zeros(1, dtype=torch.int32, device=dev)
            _cabi.call("stb200_classify_windows", self.n_win, self.win_offsets.data_ptr(), self.row_order.data_ptr(),
                       self.index_0_offsets.data_ptr(), self.index_1.data_ptr(), flags.data_ptr(), rows.data_ptr(),
                       count.data_ptr(), _stream())
            self._fused = (flags, rows[:int(count.item())].contiguous())
        return self._fused

    def _pack(self, L: int, perm):
        out = torch.empty(self.M, dtype=torch.int32, device=self.index_1.device)
        _cabi.call("stb200_pack_rel", self.M, int(L), self.rel_idx.data_ptr(), None if perm is None else perm.data_ptr(),
                   out.data_ptr(), _stream())
        return out

    def _length_order(self, offsets: torch.Tensor) -> torch.Tensor:
        lib = _cabi.load()
        nbytes = lib.stb200_length_order_workspace_bytes(self.N)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=offsets.device)
        out = torch.empty(self.N, dtype=torch.int32, device=offsets.device)
        _cabi.call("stb200_length_order", self.N, offsets.data_ptr(), None if self.row_order is None else self.row_order.data_ptr(),
                   out.data_ptr(), ws.data_ptr(), nbytes, _stream())
        return out

    @property
    def len_orders(self):
        """(queries, keys) sorted by pair count: tiles of equally long rows for the table-gradient kernels (backward only)"""
        if self._len_orders is None:
            self._len_orders = (self._length_order(self.index_0_offsets), self._length_order(self.tcsr.t_offsets))
        return self._len_orders

    def c_struct(self, L: int, backward: bool = False) -> "_cabi.IndexStruct":
        """`stb200_index` for the fused entry points; packs the rel-pos bins on first use (per table length)."""
        if self.index_1 is None or self.rel_idx is None:
            raise ValueError("PairIndex: the per-op entry points need the CSR pair list with its rel-pos index (built with csr=True "
                             "and a quant_size, or rel_idx filled by WindowAttention.forward); this index only carries the fused plan")
        ent = self._packed.get(L)
        if ent is None:
            ent = [self._pack(L, None), None]
            self._packed[L] = ent
        if backward and ent[1] is None:
            ent[1] = self._pack(L, self.tcsr.t_pair)
        st = _cabi.IndexStruct(self.N, self.M, self.index_0_offsets.data_ptr(), self.index_1.data_ptr(),
                               self.rel_idx.data_ptr(), None, None, None, ent[0].data_ptr(), None,
                               None if self.row_order is None else self.row_order.data_ptr(), None, None)
        if backward:
            t = self.tcsr
            st.t_offsets, st.t_pair, st.t_index0 = t.t_offsets.data_ptr(), t.t_pair.data_ptr(), t.t_index0.data_ptr()
            st.t_rel_packed = ent[1].data_ptr()
            if self.N and not os.environ.get("STB200_NO_LEN_ORDER"):   # (knob for A/B measurements)
                lo = self.len_orders
                st.len_order, st.t_len_order = lo[0].data_ptr(), lo[1].data_ptr()
        return st


def fps_new_offset(offset: torch.Tensor, downsample_scale: int) -> torch.Tensor:
    """per scene n_i // ds + 1, cumulative (stratified_transformer.py:282-288), without the .item() loop."""
    counts = torch.diff(offset, prepend=offset.new_zeros(1))
    return torch.cumsum(torch.div(counts, downsample_scale, rounding_mode="floor") + 1, 0).to(torch.int32)


def fps_prefix(long_idx: torch.Tensor, long_offset: torch.Tensor, short_offset: torch.Tensor, total: int | None = None) -> torch.Tensor:
    """Furthest point sampling is greedy from point 0 of every scene, so the picks for a smaller sample count are the first
    picks of a longer run over the same scenes.  long_idx: picks of the longer run, long_offset / short_offset: cumulative counts
    per scene of the longer / the wanted run (short count <= long count in every scene).  Returns the shorter run's picks."""
    lo = long_offset.to(torch.int64)
    so = short_offset.to(device=lo.device, dtype=torch.int64)
    long_start = lo - torch.diff(lo, prepend=lo.new_zeros(1))
    short_cnt = torch.diff(so, prepend=so.new_zeros(1))
    short_start = so - short_cnt
    if total is None:
        total = int(so[-1])        # host sync; pass `total` (the last entry of short_offset) to stay asynchronous
    scene = torch.repeat_interleave(torch.arange(so.numel(), device=lo.device), short_cnt, output_size=total)
    pos = torch.arange(total, device=lo.device) - short_start[scene] + long_start[scene]
    return long_idx[pos].contiguous()


def build_stratified_index(xyz: torch.Tensor, offset: torch.Tensor, window_size: float, quant_size: float | None,
                           downsample_idx: torch.Tensor | None, parity: int, want_index_0: bool = False,
                           workspace: torch.Tensor | None = None, fused: bool = False, csr: bool = True,
                           swin_shift: float | None = None) -> PairIndex:
    """One block parity of `get_indice_pairs` + sort + CSR (+ rel-pos index when quant_size is given).
    downsample_idx=None gives dense window pairs only (the Swin / 3DSwin variant).
    fused=True also builds the work plan of the fused kernels (`PairIndex.plan`); csr=False then skips the M-sized CSR
    arrays, which only the per-op entry points need.  swin_shift: the plan carries the 3DSwin rel-pos index."""
    if not (xyz.is_cuda and xyz.dtype == torch.float32 and xyz.is_contiguous() and xyz.dim() == 2 and xyz.shape[1] == 3):
        raise TypeError("xyz must be a contiguous CUDA float32 [N,3] tensor")
    N, b = xyz.shape[0], offset.numel()
    dev = xyz.device
    offset = offset.to(device=dev, dtype=torch.int32).contiguous()
    lib = _cabi.load()
    nbytes = lib.stb200_pair_builder_workspace_bytes(N)
    if workspace is None or workspace.numel() < nbytes:
        workspace = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    m = 0 if downsample_idx is None else downsample_idx.numel()
    ds_ptr = None
    if m:
        downsample_idx = downsample_idx.to(torch.int32).contiguous()
        ds_ptr = downsample_idx.data_ptr()
    offsets = torch.empty(N + 1, dtype=torch.int32, device=dev)
    totals = torch.empty(4, dtype=torch.int32, device=dev)
    _cabi.call("stb200_stratified_pairs_count", N, b, xyz.data_ptr(), offset.data_ptr(), float(window_size),
               int(parity) & 1, ds_ptr, m, workspace.data_ptr(), workspace.numel(), offsets.data_ptr(),
               totals.data_ptr(), _stream())
    plan_part = _plan_count(N, m > 0, workspace, FUSED_BLOCKS) if fused else None
    if fused:
        both = torch.cat([totals, plan_part[1]]).tolist()      # the one host sync: the caller has to allocate the outputs
        (M, n_max, err, n_win), ptot = both[:4], both[4:]
    else:
        M, n_max, err, n_win = totals.tolist()
    if err:
        raise _cabi.Stb200Error("pair builder: window grid has more than 2^32 cells (window too small for the scene extent)")
    plan = None
    if fused:
        plan = _plan_fill(N, xyz, window_size, quant_size, n_win, m, workspace, plan_part[0], plan_part[1], ptot, FUSED_BLOCKS, swin_shift)
        if not csr and plan is not None:
            return PairIndex(None, None, None, int(n_max), int(M), None, plan.order_s, plan.wstart_s, int(n_win), plan)
    index_1 = torch.empty(M, dtype=torch.int32, device=dev)
    rel_idx = torch.empty(M, 3, dtype=torch.int32, device=dev) if quant_size is not None else None
    index_0 = torch.empty(M, dtype=torch.int32, device=dev) if want_index_0 else None
    row_order = torch.empty(N, dtype=torch.int32, device=dev)
    win_offsets = torch.empty(n_win + 1, dtype=torch.int32, device=dev)
    if M:
        _cabi.call("stb200_stratified_pairs_fill", N, xyz.data_ptr(), float(2 * window_size),
                   float(quant_size if quant_size is not None else 1.0), int(m > 0), workspace.data_ptr(),
                   workspace.numel(), offsets.data_ptr(), index_1.data_ptr(),
                   None if rel_idx is None else rel_idx.data_ptr(), None if index_0 is None else index_0.data_ptr(),
                   row_order.data_ptr(), win_offsets.data_ptr(), n_win, M, _stream())
    return PairIndex(offsets, index_1, rel_idx, int(n_max), int(M), index_0, row_order if M else None,
                     win_offsets if M else None, int(n_win) if M else 0, plan)


def rel_pos_index_stratified(xyz, index_0_offsets, index_1, window_size: float, quant_size: float) -> torch.Tensor:
    """relative_position_index of stratified_transformer.py:186-188 for an existing CSR pair list -> int32 [M,3]."""
    N = index_0_offsets.numel() - 1
    out = torch.empty(index_1.numel(), 3, dtype=torch.int32, device=xyz.device)
    _cabi.call("stb200_rel_pos_index_stratified", N, xyz.data_ptr(), index_0_offsets.data_ptr(), index_1.data_ptr(),
               float(2 * window_size), float(quant_size), out.data_ptr(), _stream())
    return out


def rel_pos_index_swin(xyz, index_0_offsets, index_1, window_size: float, quant_size: float, shift_size: float) -> torch.Tensor:
    """relative_position_index of swin3d_transformer.py:151-154 (+ map_func :129-130) -> int32 [M,3]."""
    N = index_0_offsets.numel() - 1
    out = torch.empty(index_1.numel(), 3, dtype=torch.int32, device=xyz.device)
    xq = torch.empty(N, 3, dtype=torch.float32, device=xyz.device)
    mm = torch.empty(6, dtype=torch.int32, device=xyz.device)
    _cabi.call("stb200_rel_pos_index_swin", N, xyz.data_ptr(), index_0_offsets.data_ptr(), index_1.data_ptr(),
               float(window_size), float(quant_size), float(shift_size), int(window_size / quant_size),
               xq.data_ptr(), mm.data_ptr(), out.data_ptr(), _stream())
    return out


@dataclass
class LayerIndex:
    """Everything `BasicLayer.forward` derives from (xyz, offset) before its block loop:
    FPS-sampled key candidates and the two pair lists (even / odd blocks)."""
    downsample_idx: torch.Tensor | None
    parity: tuple
    ready: object = None      # CUDA event recorded when the index was produced on a side stream

    def for_block(self, i: int) -> PairIndex:
        return self.parity[i % 2]


def build_layer_index(xyz: torch.Tensor, offset: torch.Tensor, window_size: float, quant_size: float,
                      downsample_scale: int | None, want_index_0: bool = False, parities=(0, 1), fused: bool = False,
                      csr: bool = True, downsample_idx: torch.Tensor | None = None) -> LayerIndex:
    """stratified_transformer.py:267-317 for one layer: FPS (n_i // ds + 1 samples per scene), then both parities.
    downsample_scale=None -> dense-only pairs (Swin).  downsample_idx: the FPS picks when the caller already has them
    (`fps_prefix`: they are a prefix of any longer FPS run over the same scenes)."""
    offset = offset.to(device=xyz.device, dtype=torch.int32)
    ds_idx = downsample_idx
    if downsample_scale is not None and ds_idx is None:
        ds_idx = pointops.furthestsampling(xyz, offset, fps_new_offset(offset, downsample_scale))
    ws = torch.empty(_cabi.load().stb200_pair_builder_workspace_bytes(xyz.shape[0]), dtype=torch.uint8, device=xyz.device)
    built = {p: build_stratified_index(xyz, offset, window_size, quant_size, ds_idx, p, want_index_0, ws, fused, csr) for p in parities}
    return LayerIndex(ds_idx, tuple(built.get(p) for p in (0, 1)))


# ------------------------------------------------------------------------------------------------
# Split-phase construction + prefetching.  The geometry of a batch (FPS, window partition, pair lists) depends only
# on coordinates, never on features, so a training loop can compute it for batch t+1 on a side stream while the
# attention of batch t runs (the reference does its geometric pre-step, tp.ball_query, in the data loop for the same
# reason: train.py:323-325).  The only host round trip of the builder (reading M to allocate the pair arrays) is
# split off: `start` enqueues everything up to the per-query counts without blocking the host, `finish` blocks on
# an event that has usually fired long ago, then enqueues the fill / transposed CSR / packing kernels.
class PendingLayerIndex:
    def __init__(self, xyz, offset, window_size, quant_size, downsample_scale, offset_host, want_index_0=False, L=None,
                 fused=False, csr=True, downsample_idx=None):
        self.xyz, self.window_size, self.quant_size, self.want_index_0, self.L = xyz, window_size, quant_size, want_index_0, L
        self.fused, self.csr = fused, csr or not fused
        dev = xyz.device
        N, b = xyz.shape[0], len(offset_host)
        self.N = N
        offset = offset.to(device=dev, dtype=torch.int32).contiguous()
        lib = _cabi.load()
        self.ds_idx = downsample_idx          # FPS picks the caller already has (index.fps_prefix), else sampled here
        m = 0 if downsample_idx is None else int(downsample_idx.shape[0])
        if downsample_scale is not None and downsample_idx is None:
            sizes = [int(offset_host[0])] + [int(offset_host[i] - offset_host[i - 1]) for i in range(1, b)]
            new_counts = [n // downsample_scale + 1 for n in sizes]
            m = sum(new_counts)
            new_offset = torch.tensor(new_counts, dtype=torch.int32).cumsum(0).to(torch.int32).to(dev, non_blocking=True)
            self.ds_idx = torch.empty(m, dtype=torch.int32, device=dev)
            ext.furthestsampling_cuda(b, max(sizes), xyz, offset, new_offset, None, self.ds_idx)
            self._keep = (new_offset,)
        nbytes = lib.stb200_pair_builder_workspace_bytes(N)
        self.parts = []
        for parity in (0, 1):
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            offsets = torch.empty(N + 1, dtype=torch.int32, device=dev)
            totals = torch.empty(4, dtype=torch.int32, device=dev)
            _cabi.call("stb200_stratified_pairs_count", N, b, xyz.data_ptr(), offset.data_ptr(), float(window_size), parity,
                       None if self.ds_idx is None else self.ds_idx.data_ptr(), m, ws.data_ptr(), nbytes, offsets.data_ptr(),
                       totals.data_ptr(), _stream())
            plan_part = _plan_count(N, m > 0, ws, FUSED_BLOCKS) if fused else None
            host = torch.empty(44 if fused else 4, dtype=torch.int32, pin_memory=True)
            host[:4].copy_(totals, non_blocking=True)
            if fused:
                host[4:].copy_(plan_part[1], non_blocking=True)
            self.parts.append((ws, offsets, totals, host, plan_part))
        self.m = m
        self.offset = offset
        self.counted = torch.cuda.current_stream().record_event()

    def finish(self) -> "LayerIndex":
        """Call under the same stream as the constructor."""
        self.counted.synchronize()
        dev = self.xyz.device
        built = []
        for ws, offsets, totals, host, plan_part in self.parts:
            hl = host.tolist()
            M, n_max, err, n_win = hl[:4]
            if err:
                raise _cabi.Stb200Error("pair builder: window grid has more than 2^32 cells")
            plan = None
            if self.fused:
                plan = _plan_fill(self.N, self.xyz, self.window_size, self.quant_size, n_win, self.m, ws, plan_part[0], plan_part[1],
                                  hl[4:], FUSED_BLOCKS)
                if not self.csr and plan is not None:
                    built.append(PairIndex(None, None, None, int(n_max), int(M), None, plan.order_s, plan.wstart_s, int(n_win), plan))
                    continue
            index_1 = torch.empty(M, dtype=torch.int32, device=dev)
            rel_idx = torch.empty(M, 3, dtype=torch.int32, device=dev)
            index_0 = torch.empty(M, dtype=torch.int32, device=dev) if self.want_index_0 else None
            row_order = torch.empty(self.N, dtype=torch.int32, device=dev)
            win_offsets = torch.empty(n_win + 1, dtype=torch.int32, device=dev)
            if M:
                _cabi.call("stb200_stratified_pairs_fill", self.N, self.xyz.data_ptr(), float(2 * self.window_size),
                           float(self.quant_size), int(self.m > 0), ws.data_ptr(), ws.numel(), offsets.data_ptr(),
                           index_1.data_ptr(), rel_idx.data_ptr(), None if index_0 is None else index_0.data_ptr(),
                           row_order.data_ptr(), win_offsets.data_ptr(), n_win, M, _stream())
            pi = PairIndex(offsets, index_1, rel_idx, int(n_max), int(M), index_0, row_order if M else None,
                           win_offsets if M else None, int(n_win) if M else 0, plan)
            if self.L is not None and not self.fused:
                pi.c_struct(self.L, backward=True)   # transposed CSR + packed bins, eagerly, on this stream
            built.append(pi)
        self.parts = None
        li = LayerIndex(self.ds_idx, tuple(built))
        li.ready = torch.cuda.current_stream().record_event()
        return li


def record_stream(li: "LayerIndex", stream) -> None:
    """Tell the caching allocator that `stream` consumes the index tensors (they were allocated on another stream)."""
    for pi in li.parity:
        ts = [pi.index_0_offsets, pi.index_1, pi.rel_idx, pi.index_0, pi.row_order, pi.win_offsets]
        if pi.plan is not None:
            ts += pi.plan.tensors()
        if pi._fused is not None:
            ts += list(pi._fused)
        if pi._tcsr is not None:
            ts += [pi._tcsr.t_offsets, pi._tcsr.t_pair, pi._tcsr.t_index0]
        for a, b in pi._packed.values():
            ts += [a, b]
        if pi._len_orders is not None:
            ts += list(pi._len_orders)
        for t in ts:
            if t is not None:
                t.record_stream(stream)
    if li.downsample_idx is not None:
        li.downsample_idx.record_stream(stream)


class GeometryPrefetcher:
    """Double-buffered geometry pipeline for a stack of layers.

        pf = GeometryPrefetcher(layer_cfgs)          # [(window, quant, downsample_scale, table_len), ...]
        pf.submit(xyzs, offsets, offsets_host)        # batch 0
        for batch in loader:
            geo = pf.take()                           # list of LayerIndex for the current batch (main stream waits on it)
            pf.submit(next xyzs, ...)                 # geometry of the next batch starts on the side stream
            ... enqueue attention of the current batch on the main stream ...
            pf.complete()                             # host: wait for the counts, enqueue fill / transpose / pack

    `submit` may be called anywhere between `take` and `complete`: the geometry starts once the main-stream work enqueued
    so far has finished.  FPS occupies up to half of the SMs for ~10 ms, so with >= 4 scenes per GPU it costs least when it
    is submitted after the two machine-filling first layers have been enqueued (bench.py: 71.9 -> 70.4 ms per step); with
    1-2 scenes per GPU submit at the top of the step.
    """

    def __init__(self, layer_cfgs, device=None, fused=False, csr=True):
        self.cfgs = layer_cfgs
        self.fused, self.csr = fused, csr
        # stream priority of the geometry work relative to the attention stream (0 = default, -1 = higher)
        prio = int(os.environ.get("STB200_SIDE_PRIORITY", "0"))
        # One side stream per layer (STB200_GEOMETRY_STREAMS=1: a single one, the round-1 behaviour): the layers' point sets
        # are independent, and each layer's FPS is a sequential loop of ~n/ds iterations that occupies a few SMs, so the
        # four loops run side by side instead of back to back (critical path = layer 0's FPS + its pair lists).
        n_streams = max(1, min(len(layer_cfgs), int(os.environ.get("STB200_GEOMETRY_STREAMS", str(len(layer_cfgs))))))
        self.sides = [torch.cuda.Stream(device=device, priority=prio) for _ in range(n_streams)]
        self.side = self.sides[0]
        self.pending = None
        self.done = None

    def _stream_of(self, layer):
        return self.sides[layer % len(self.sides)]

    def submit(self, xyzs, offsets, offsets_host):
        main = torch.cuda.current_stream()
        self.pending = []
        for i, (x, o, oh, (w, q, ds, L)) in enumerate(zip(xyzs, offsets, offsets_host, self.cfgs)):
            side = self._stream_of(i)
            side.wait_stream(main)            # inputs produced on the main stream are visible,
            if side is not self.side:
                side.wait_stream(self.side)   # and so are uploads the caller enqueued on `self.side`
            with torch.cuda.stream(side):
                self.pending.append(PendingLayerIndex(x, o, w, q, ds, oh, L=L, fused=self.fused, csr=self.csr))
            x.record_stream(side)

    def complete(self):
        if self.pending is None:
            return
        self.done = []
        for i, p in enumerate(self.pending):
            with torch.cuda.stream(self._stream_of(i)):
                self.done.append(p.finish())
        self.pending = None

    def take(self):
        if self.done is None:
            self.complete()
        geo, self.done = self.done, None
        main = torch.cuda.current_stream()
        for li in geo:
            main.wait_event(li.ready)
            record_stream(li, main)
        return geo
