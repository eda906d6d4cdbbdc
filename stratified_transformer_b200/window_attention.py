"""Host-side mirror of the reference's `WindowAttention` module
(/root/reference/model/stratified_transformer.py:114-217), the orchestrator of the hot path.

Same constructor arguments, parameter names (`qkv`, `proj`, `relative_pos_{query,key,value}_table` with
shape [2*quant_grid_length, heads, head_dim, 3]) and forward semantics, so reference checkpoints load.
What differs is only what runs underneath: the pair ops come from `pointops` (libstb200.so), the bias add +
scatter_softmax is one segment-softmax kernel, the relative position index is taken from the `PairIndex`
(computed once per layer parity instead of five elementwise kernels + two asserts per block), and the int32
index tensors are created once instead of `.int()` casts at every call site (:183,194,208).

The qkv / proj Linear layers are plain torch (cuBLAS): dense GEMMs are not part of this path.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import pointops
from .index import PairIndex, rel_pos_index_stratified


class WindowAttention(nn.Module):
    def __init__(self, dim, window_size, num_heads, quant_size, rel_query=True, rel_key=False, rel_value=False,
                 qkv_bias=True, qk_scale=None, attn_drop=0., proj_drop=0.):
        super().__init__()
        self.dim = dim
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = qk_scale or head_dim ** -0.5
        self.window_size = window_size
        self.quant_size = quant_size
        self.rel_query, self.rel_key, self.rel_value = rel_query, rel_key, rel_value
        quant_grid_length = int((2 * window_size + 1e-4) // quant_size)
        self.quant_grid_length = quant_grid_length

        def table():
            t = nn.Parameter(torch.zeros(2 * quant_grid_length, num_heads, head_dim, 3))
            nn.init.trunc_normal_(t, std=.02)
            return t

        if rel_query:
            self.relative_pos_query_table = table()
        if rel_key:
            self.relative_pos_key_table = table()
        if rel_value:
            self.relative_pos_value_table = table()
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop, inplace=True)  # constructed but never applied, like the reference
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop, inplace=True)

    def forward(self, feats, xyz, index_0, index_1=None, index_0_offsets=None, n_max=None):
        """Either the reference's call `forward(feats, xyz, index_0, index_1, index_0_offsets, n_max)`
        (int64 or int32 index tensors) or `forward(feats, xyz, pair_index)` with a prebuilt PairIndex."""
        if isinstance(index_0, PairIndex):
            idx = index_0
        else:
            off32 = index_0_offsets.int().contiguous()
            i1_32 = index_1.int().contiguous()
            idx = PairIndex(off32, i1_32, None, 0, int(index_1.shape[0]), index_0.int().contiguous())
        if idx.rel_idx is None and idx.index_1 is not None:
            idx.rel_idx = rel_pos_index_stratified(xyz, idx.index_0_offsets, idx.index_1, self.window_size, self.quant_size)
        N, C = feats.shape
        h = self.num_heads
        # same math as `qkv(feats).reshape(N,3,h,d).permute(1,0,2,3).contiguous()` followed by `query * scale` and the
        # `.float()` casts at the pointops call sites: ONE GEMM without bias (scale folded into the q rows of the weight),
        # then one kernel that adds the bias and writes q, k, v as contiguous fp32 [N, h, d] (pointops.split_qkv); no other
        # N-sized permute / multiply / cast pass around the pair ops, forward or backward
        Wm, bm = self.qkv.weight, self.qkv.bias
        if C % 8 == 0:
            w_all = torch.cat([Wm[:C] * self.scale, Wm[C:]], 0)
            b_all = None if bm is None else torch.cat([bm[:C] * self.scale, bm[C:]], 0)
            query, key, value = pointops.split_qkv(torch.nn.functional.linear(feats, w_all), b_all, h)
        else:   # channel counts the split kernel does not take: three GEMMs whose outputs are already contiguous
            query = torch.nn.functional.linear(feats, Wm[:C] * self.scale, None if bm is None else bm[:C] * self.scale).view(N, h, C // h)
            key = torch.nn.functional.linear(feats, Wm[C:2 * C], None if bm is None else bm[C:2 * C]).view(N, h, C // h)
            value = torch.nn.functional.linear(feats, Wm[2 * C:], None if bm is None else bm[2 * C:]).view(N, h, C // h)
        if idx.plan is not None and self.rel_query and self.rel_key and self.rel_value and C // h == 16 and \
                not getattr(self, "per_op", False):
            # window-centric fused kernels: the whole pair path (logits + rel-pos bias + softmax + aggregation, and its
            # backward) without any [M,h] tensor; available whenever the index came from the device builder with a plan
            x = pointops.window_attention_plan(query.float(), key.float(), value.float(), self.relative_pos_query_table.float(),
                                               self.relative_pos_key_table.float(), self.relative_pos_value_table.float(), idx.plan)
            x = x.view(N, C)
            if not torch.is_autocast_enabled():
                x = x.to(self.proj.weight.dtype)
            return self.proj_drop(self.proj(x))
        off, i1, rel = idx.index_0_offsets, idx.index_1, idx.rel_idx
        if getattr(self, "fused_forward", False) and self.rel_query and self.rel_key and self.rel_value:
            # opt-in (module.fused_forward = True): whole pair path in one op, per-window tensor-core kernel where the
            # window structure allows it.  First version, currently slower than the per-pair kernels.
            x = pointops.window_attention_fused(query.float(), key.float(), value.float(),
                                                self.relative_pos_query_table.float(), self.relative_pos_key_table.float(),
                                                self.relative_pos_value_table.float(), idx)
            x = x.view(N, C)
            if not torch.is_autocast_enabled():
                x = x.to(self.proj.weight.dtype)
            return self.proj_drop(self.proj(x))
        bias = None
        fused = self.rel_query and self.rel_key
        if fused:   # q.k + rel-pos bias in one pass over the pairs
            attn_flat = pointops.window_logits(query.float(), key.float(), self.relative_pos_query_table.float(),
                                               self.relative_pos_key_table.float(), idx)
        else:
            attn_flat = pointops.attention_step1_v2(query.float(), key.float(), i1, off, idx.n_max)
        if fused:
            pass
        elif self.rel_query and self.rel_key:
            bias = pointops.dot_prod_with_idx_v3(query.float(), off, idx.n_max, key.float(), i1,
                                                 self.relative_pos_query_table.float(),
                                                 self.relative_pos_key_table.float(), rel)
        elif self.rel_query or self.rel_key:
            if idx.index_0 is None:
                idx.index_0 = torch.repeat_interleave(torch.arange(N, device=off.device, dtype=torch.int32),
                                                      (off[1:] - off[:-1]).long())
            if self.rel_query:
                bias = pointops.dot_prod_with_idx(query.float(), idx.index_0, self.relative_pos_query_table.float(), rel)
            else:
                bias = pointops.dot_prod_with_idx(key.float(), i1, self.relative_pos_key_table.float(), rel)
        softmax_attn_flat = pointops.segment_softmax(attn_flat, off, bias)
        if self.rel_value:
            x = pointops.window_aggregate(softmax_attn_flat, value.float(), self.relative_pos_value_table.float(), idx)
        else:
            if idx.index_0 is None:
                idx.index_0 = torch.repeat_interleave(torch.arange(N, device=off.device, dtype=torch.int32),
                                                      (off[1:] - off[:-1]).long())
            x = pointops.attention_step2(softmax_attn_flat, value.float(), idx.index_0, i1)
        x = x.view(N, C)
        if not torch.is_autocast_enabled():
            x = x.to(self.proj.weight.dtype)
        x = self.proj(x)
        x = self.proj_drop(x)
        return x


class SwinWindowAttention(WindowAttention):
    """Mirror of the 3DSwin variant (/root/reference/model/swin3d_transformer.py:81-178): dense window pairs only,
    tables of length 2*int(window/quant)-1, rel-pos index from per-point quantised coordinates (:151-154), forward
    signature `(feats, xyz, index_0, index_0_offsets, n_max, index_1, shift_size)`."""

    def __init__(self, dim, window_size, num_heads, quant_size, rel_query=True, rel_key=False, rel_value=False,
                 qkv_bias=True, qk_scale=None, attn_drop=0., proj_drop=0.):
        nn.Module.__init__(self)
        self.dim = dim
        self.window_size = window_size
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = qk_scale or head_dim ** -0.5
        self.quant_size = quant_size
        self.rel_query, self.rel_key, self.rel_value = rel_query, rel_key, rel_value
        quant_grid_length = int(window_size / quant_size)
        self.quant_grid_length = quant_grid_length

        def table():
            t = nn.Parameter(torch.zeros(2 * quant_grid_length - 1, num_heads, head_dim, 3))
            nn.init.trunc_normal_(t, std=.02)
            return t

        if rel_query:
            self.relative_pos_query_table = table()
        if rel_key:
            self.relative_pos_key_table = table()
        if rel_value:
            self.relative_pos_value_table = table()
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop, inplace=True)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop, inplace=True)

    def forward(self, feats, xyz, index_0, index_0_offsets, n_max, index_1, shift_size):
        from .index import rel_pos_index_swin
        shift = float(shift_size.flatten()[0]) if isinstance(shift_size, torch.Tensor) else float(shift_size)
        off32, i1_32 = index_0_offsets.int().contiguous(), index_1.int().contiguous()
        idx = PairIndex(off32, i1_32, None, 0, int(index_1.shape[0]), None if index_0 is None else index_0.int().contiguous())
        idx.rel_idx = rel_pos_index_swin(xyz, off32, i1_32, self.window_size, self.quant_size, shift)
        return WindowAttention.forward(self, feats, xyz, idx)
