"""Operator-level mirror of /root/reference/lib/pointops2/functions/pointops.py (hot-path subset).

Same public names, argument orders, shapes and dtypes, so `WindowAttention`
(model/stratified_transformer.py:164-217, model/swin3d_transformer.py:132-178) can import this module
in place of `lib.pointops2.functions.pointops`:

    attention_step1_v2(q, k, index1, index0_offsets, n_max)                                  pointops.py:203
    dot_prod_with_idx_v3(q, index_q_offsets, n_max, k, index_k, table_q, table_k, rel_idx)  pointops.py:519
    attention_step2_with_rel_pos_value_v2(attn, v, index0_offsets, n_max, index1, table, rel_idx)   :646
    furthestsampling(xyz, offset, new_offset)                                                 pointops.py:31
    attention_step1 / attention_step2 / attention_step2_v2 / dot_prod_with_idx / dot_prod_with_idx_v2 /
    attention_step2_with_rel_pos_value                                                (v1, :140-581)

plus `segment_softmax`, which replaces `attn_flat + bias` and torch_scatter's `scatter_softmax` on the path.

Differences from the reference, all at the edges: inputs are validated (CUDA, dtype, contiguity) with real
exceptions instead of asserts; `n_max` may be an int or a tensor and is never synchronised on (no launch
shape depends on it); the ops run on torch's current stream; inputs are pinned to fp32 under autocast
(the reference relies on call sites passing `.float()`).
"""
from __future__ import annotations

import ctypes

import torch
from torch.autograd import Function

from . import _cabi
from . import pointops2_cuda as pointops_cuda


def _n_max_int(n_max) -> int:
    # accepted for signature compatibility; kernels ignore it, so never force a device sync on a CUDA scalar
    if isinstance(n_max, torch.Tensor):
        return 0
    return int(n_max)


def _contig(*tensors):
    for t in tensors:
        if not t.is_contiguous():
            raise ValueError("pointops: all tensors must be contiguous (same contract as the reference asserts)")


class FurthestSampling(Function):
    @staticmethod
    def forward(ctx, xyz, offset, new_offset):
        """xyz (n,3) f32, offset (b) i32 cumulative, new_offset (b) i32 cumulative -> idx (m) i32."""
        _contig(xyz)
        offset = offset.int().contiguous()
        new_offset = new_offset.int().contiguous()
        b = offset.shape[0]
        sizes = torch.diff(offset, prepend=offset.new_zeros(1))
        host = torch.stack([sizes.max(), new_offset[b - 1]]).tolist()  # one sync, like the reference's .item()
        idx = torch.zeros(int(host[1]), dtype=torch.int32, device=xyz.device)
        pointops_cuda.furthestsampling_cuda(b, int(host[0]), xyz, offset, new_offset, None, idx)
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, grad):
        return None, None, None


furthestsampling = FurthestSampling.apply


class AttentionStep1_v2(Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, q, k, index1, index0_offsets, n_max):
        _contig(q, k, index0_offsets, index1)
        N_q, h, d = q.shape
        M = index1.shape[0]
        out = torch.empty(M, h, dtype=torch.float32, device=q.device)
        pointops_cuda.attention_step1_forward_cuda_v2(k.shape[0], M, h, h * d, _n_max_int(n_max), q, k,
                                                      index0_offsets, index1, out)
        ctx.save_for_backward(q, k, index0_offsets, index1)
        return out

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_output):
        q, k, index0_offsets, index1 = ctx.saved_tensors
        N_q, h, d = q.shape
        grad_output = grad_output.contiguous()
        grad_q = torch.empty_like(q)
        grad_k = torch.zeros_like(k)
        pointops_cuda.attention_step1_backward_cuda_v2(N_q, index1.shape[0], h, h * d, 0, grad_output,
                                                       index0_offsets, index1, q, k, grad_q, grad_k)
        return grad_q, grad_k, None, None, None


attention_step1_v2 = AttentionStep1_v2.apply


class DotProdWithIdx_v3(Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, q, index_q_offsets, n_max, k, index_k, table_q, table_k, rel_idx):
        _contig(q, index_q_offsets, k, index_k, table_q, table_k, rel_idx)
        N, h, hdim = q.shape
        M = index_k.shape[0]
        if table_k.shape[0] != table_q.shape[0]:
            raise ValueError("table_q and table_k must have the same length")
        out = torch.empty(M, h, dtype=torch.float32, device=q.device)
        pointops_cuda.dot_prod_with_idx_forward_cuda_v3(N, M, h, hdim, _n_max_int(n_max), q, index_q_offsets, k,
                                                        index_k, table_q, table_k, rel_idx, out)
        ctx.save_for_backward(q, index_q_offsets, k, index_k, table_q, table_k, rel_idx)
        return out

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_output):
        q, index_q_offsets, k, index_k, table_q, table_k, rel_idx = ctx.saved_tensors
        N, h, hdim = q.shape
        grad_output = grad_output.contiguous()
        grad_q = torch.empty_like(q)
        grad_k = torch.zeros_like(k)
        grad_tq = torch.zeros_like(table_q)
        grad_tk = torch.zeros_like(table_k)
        pointops_cuda.dot_prod_with_idx_backward_cuda_v3(N, index_k.shape[0], h, hdim, 0, grad_output, q,
                                                         index_q_offsets, k, index_k, table_q, table_k, rel_idx,
                                                         grad_q, grad_k, grad_tq, grad_tk)
        return grad_q, None, None, grad_k, None, grad_tq, grad_tk, None


dot_prod_with_idx_v3 = DotProdWithIdx_v3.apply


class AttentionStep2WithRelPosValue_v2(Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, attn, v, index0_offsets, n_max, index1, table, rel_idx):
        _contig(attn, v, index0_offsets, index1, table, rel_idx)
        M, h = attn.shape
        N, _, hdim = v.shape
        out = torch.empty(N, h, hdim, dtype=torch.float32, device=v.device)
        pointops_cuda.attention_step2_with_rel_pos_value_forward_cuda_v2(N, M, h, hdim, _n_max_int(n_max), attn, v,
                                                                         index0_offsets, index1, table, rel_idx, out)
        ctx.save_for_backward(attn, v, index0_offsets, index1, table, rel_idx)
        return out

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_output):
        attn, v, index0_offsets, index1, table, rel_idx = ctx.saved_tensors
        M, h = attn.shape
        N, _, hdim = v.shape
        grad_output = grad_output.contiguous()
        grad_attn = torch.empty_like(attn)
        grad_v = torch.zeros_like(v)
        grad_table = torch.zeros_like(table)
        pointops_cuda.attention_step2_with_rel_pos_value_backward_cuda_v2(N, M, h, hdim, 0, grad_output,
                                                                          index0_offsets, index1, attn, v, table,
                                                                          rel_idx, grad_attn, grad_v, grad_table)
        return grad_attn, grad_v, None, None, None, grad_table, None


attention_step2_with_rel_pos_value_v2 = AttentionStep2WithRelPosValue_v2.apply


class SegmentSoftmax(Function):
    """p = softmax over each query's pairs of (a + b), per head.  Replaces
    `attn_flat + relative_position_bias` + `scatter_softmax(src, index_0, dim=0)`
    (model/stratified_transformer.py:203,205) with one kernel keyed by the CSR offsets."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, a, b, index0_offsets):
        _contig(a, index0_offsets)
        if b is not None:
            _contig(b)
        M, h = a.shape
        N = index0_offsets.shape[0] - 1
        p = torch.empty_like(a)
        pointops_cuda.segment_softmax_forward_cuda(N, M, h, a, b, index0_offsets, p)
        ctx.save_for_backward(p, index0_offsets)
        ctx.has_b = b is not None
        return p

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_p):
        p, index0_offsets = ctx.saved_tensors
        M, h = p.shape
        grad_s = torch.empty_like(p)
        pointops_cuda.segment_softmax_backward_cuda(index0_offsets.shape[0] - 1, M, h, p, grad_p.contiguous(),
                                                    index0_offsets, grad_s)
        return grad_s, (grad_s if ctx.has_b else None), None


class WindowLogits(Function):
    """logits[m,h] = <q[i0],k[i1]> + <q[i0],Eq(m)> + <k[i1],Ek(m)> in one pass over the pairs: the fused form of
    attention_step1_v2 + dot_prod_with_idx_v3 + add (model/stratified_transformer.py:183-203).  `pair_index` is a
    stratified_transformer_b200.index.PairIndex."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, q, k, table_q, table_k, pair_index):
        _contig(q, k, table_q, table_k)
        N, h, d = q.shape
        L = table_q.shape[0]
        out = torch.empty(pair_index.M, h, dtype=torch.float32, device=q.device)
        _cabi.call("stb200_window_logits_forward", ctypes.byref(pair_index.c_struct(L)), h, d, L, q.data_ptr(), k.data_ptr(),
                   table_q.data_ptr(), table_k.data_ptr(), out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        ctx.save_for_backward(q, k, table_q, table_k)
        ctx.pair_index = pair_index
        return out

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_out):
        q, k, table_q, table_k = ctx.saved_tensors
        N, h, d = q.shape
        L = table_q.shape[0]
        grad_out = grad_out.contiguous()
        gq, gk = torch.empty_like(q), torch.empty_like(k)
        gtq, gtk = torch.zeros_like(table_q), torch.zeros_like(table_k)
        ws = torch.empty(grad_out.numel() + 64, dtype=torch.float32, device=q.device)   # grad rows in transposed order
        _cabi.call("stb200_window_logits_backward_ws", ctypes.byref(ctx.pair_index.c_struct(L, backward=True)), h, d, L,
                   grad_out.data_ptr(), q.data_ptr(), k.data_ptr(), table_q.data_ptr(), table_k.data_ptr(), gq.data_ptr(),
                   gk.data_ptr(), gtq.data_ptr(), gtk.data_ptr(), ws.data_ptr(), ws.numel() * 4,
                   torch.cuda.current_stream().cuda_stream)
        return gq, gk, gtq, gtk, None


def window_logits(q, k, table_q, table_k, pair_index):
    return WindowLogits.apply(q, k, table_q, table_k, pair_index)


class WindowAggregate(Function):
    """attention_step2_with_rel_pos_value_v2 driven by a PairIndex (packed rel-pos bins, shared transposed CSR)."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, attn, v, table_v, pair_index):
        _contig(attn, v, table_v)
        N, h, d = v.shape
        L = table_v.shape[0]
        out = torch.empty(N, h, d, dtype=torch.float32, device=v.device)
        _cabi.call("stb200_window_aggregate_forward", ctypes.byref(pair_index.c_struct(L)), h, d, L, attn.data_ptr(),
                   v.data_ptr(), table_v.data_ptr(), out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        ctx.save_for_backward(attn, v, table_v)
        ctx.pair_index = pair_index
        return out

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_out):
        attn, v, table_v = ctx.saved_tensors
        N, h, d = v.shape
        L = table_v.shape[0]
        grad_out = grad_out.contiguous()
        ga, gv, gt = torch.empty_like(attn), torch.empty_like(v), torch.zeros_like(table_v)
        _cabi.call("stb200_window_aggregate_backward", ctypes.byref(ctx.pair_index.c_struct(L, backward=True)), h, d, L,
                   grad_out.data_ptr(), attn.data_ptr(), v.data_ptr(), table_v.data_ptr(), ga.data_ptr(), gv.data_ptr(),
                   gt.data_ptr(), torch.cuda.current_stream().cuda_stream)
        return ga, gv, gt, None


def window_aggregate(attn, v, table_v, pair_index):
    return WindowAggregate.apply(attn, v, table_v, pair_index)


class WindowAttentionFused(Function):
    """Whole pair path of WindowAttention.forward in one kernel per window tile:
    softmax_seg(q.k + rel-pos bias) applied to (v + rel-pos value), tensor-core table products, no M-sized
    intermediate except the probabilities kept for backward.  Windows the fused kernel cannot take (queries with
    differing key lists, or more than stb200_fused_max_keys() keys) go through the per-pair entry points on the
    complementary row list.  Backward = the single-pass gradient kernels of window_aggregate / segment_softmax /
    window_logits."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, q, k, v, table_q, table_k, table_v, pair_index):
        _contig(q, k, v, table_q, table_k, table_v)
        N, h, d = q.shape
        L = table_q.shape[0]
        plan = pair_index.fused_plan() if (d == 16 and 3 * L <= 256) else None
        stream = torch.cuda.current_stream().cuda_stream
        M = pair_index.M
        out = torch.empty(N, h, d, dtype=torch.float32, device=q.device)
        p = torch.empty(M, h, dtype=torch.float32, device=q.device)
        ix = pair_index.c_struct(L)
        rows = None
        if plan is not None:
            flags, rows = plan
            _cabi.call("stb200_window_attention_forward_fused", ctypes.byref(ix), pair_index.n_win,
                       pair_index.win_offsets.data_ptr(), flags.data_ptr(), h, d, L, q.data_ptr(), k.data_ptr(), v.data_ptr(),
                       table_q.data_ptr(), table_k.data_ptr(), table_v.data_ptr(), out.data_ptr(), p.data_ptr(), stream)
        if plan is None or rows.numel() > 0:
            # per-pair kernels on the rows the fused kernel skipped (all rows when it is not applicable)
            if rows is not None:
                ix.row_order, ix.N = rows.data_ptr(), rows.numel()
            s = torch.empty(M, h, dtype=torch.float32, device=q.device)
            _cabi.call("stb200_window_logits_forward", ctypes.byref(ix), h, d, L, q.data_ptr(), k.data_ptr(),
                       table_q.data_ptr(), table_k.data_ptr(), s.data_ptr(), stream)
            _cabi.call("stb200_segment_softmax_forward_rows", ix.N, None if rows is None else rows.data_ptr(), h, s.data_ptr(),
                       None, pair_index.index_0_offsets.data_ptr(), p.data_ptr(), stream)
            _cabi.call("stb200_window_aggregate_forward", ctypes.byref(ix), h, d, L, p.data_ptr(), v.data_ptr(),
                       table_v.data_ptr(), out.data_ptr(), stream)
        ctx.save_for_backward(q, k, v, table_q, table_k, table_v, p)
        ctx.pair_index = pair_index
        return out

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_out):
        q, k, v, table_q, table_k, table_v, p = ctx.saved_tensors
        pi = ctx.pair_index
        N, h, d = q.shape
        L = table_q.shape[0]
        stream = torch.cuda.current_stream().cuda_stream
        grad_out = grad_out.contiguous()
        ix = pi.c_struct(L, backward=True)
        gp, gv, gtv = torch.empty_like(p), torch.empty_like(v), torch.zeros_like(table_v)
        _cabi.call("stb200_window_aggregate_backward", ctypes.byref(ix), h, d, L, grad_out.data_ptr(), p.data_ptr(), v.data_ptr(),
                   table_v.data_ptr(), gp.data_ptr(), gv.data_ptr(), gtv.data_ptr(), stream)
        gs = torch.empty_like(p)
        pointops_cuda.segment_softmax_backward_cuda(N, pi.M, h, p, gp, pi.index_0_offsets, gs)
        gq, gk = torch.empty_like(q), torch.empty_like(k)
        gtq, gtk = torch.zeros_like(table_q), torch.zeros_like(table_k)
        ws = torch.empty(gs.numel() + 64, dtype=torch.float32, device=q.device)
        _cabi.call("stb200_window_logits_backward_ws", ctypes.byref(ix), h, d, L, gs.data_ptr(), q.data_ptr(), k.data_ptr(),
                   table_q.data_ptr(), table_k.data_ptr(), gq.data_ptr(), gk.data_ptr(), gtq.data_ptr(), gtk.data_ptr(),
                   ws.data_ptr(), ws.numel() * 4, stream)
        return gq, gk, gv, gtq, gtk, gtv, None


def window_attention_fused(q, k, v, table_q, table_k, table_v, pair_index):
    return WindowAttentionFused.apply(q, k, v, table_q, table_k, table_v, pair_index)


class WindowAttentionPlan(Function):
    """The whole pair path of WindowAttention.forward (model/stratified_transformer.py:183-210) and its backward on the
    window-centric fused kernels (include/stb200.h, "Window-centric fused attention"):
    out = softmax_seg(q.k + rel-pos bias) applied to (v + rel-pos value), no [M,h] tensor in either direction; the forward
    keeps only the output and the log-sum-exp of every (query, head) row.  `plan` is a stratified_transformer_b200.index.FusedPlan."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, q, k, v, table_q, table_k, table_v, plan):
        _contig(q, k, v, table_q, table_k, table_v)
        N, h, d = q.shape
        if d != 16:
            raise ValueError("fused window attention supports head dim 16 (use the per-op entry points otherwise)")
        L = table_q.shape[0]
        passes, n_passes = plan.passes(L)
        out = torch.empty(N, h, d, dtype=torch.float32, device=q.device)
        lse = torch.empty(N, h, dtype=torch.float32, device=q.device)
        lsum = torch.empty(N, h, dtype=torch.float32, device=q.device)
        _cabi.call("stb200_fused_attention_forward", passes, n_passes, N, h, L, q.data_ptr(), k.data_ptr(), v.data_ptr(),
                   table_q.data_ptr(), table_k.data_ptr(), table_v.data_ptr(), out.data_ptr(), lse.data_ptr(), lsum.data_ptr(),
                   torch.cuda.current_stream().cuda_stream)
        ctx.save_for_backward(q, k, v, table_q, table_k, table_v, out, lse)
        ctx.plan = plan
        return out

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_out):
        q, k, v, table_q, table_k, table_v, out, lse = ctx.saved_tensors
        plan = ctx.plan
        N, h, d = q.shape
        L = table_q.shape[0]
        passes, n_passes = plan.passes(L)
        grad_out = grad_out.contiguous()
        gq = torch.empty_like(q)
        alloc = torch.zeros_like if plan.needs_zeroed_key_grads else torch.empty_like
        gk, gv = alloc(k), alloc(v)
        gtq, gtk, gtv = torch.zeros_like(table_q), torch.zeros_like(table_k), torch.zeros_like(table_v)
        _cabi.call("stb200_fused_attention_backward", passes, n_passes, N, h, L, grad_out.data_ptr(), out.data_ptr(), lse.data_ptr(),
                   q.data_ptr(), k.data_ptr(), v.data_ptr(), table_q.data_ptr(), table_k.data_ptr(), table_v.data_ptr(),
                   gq.data_ptr(), gk.data_ptr(), gv.data_ptr(), gtq.data_ptr(), gtk.data_ptr(), gtv.data_ptr(),
                   torch.cuda.current_stream().cuda_stream)
        return gq, gk, gv, gtq, gtk, gtv, None


def window_attention_plan(q, k, v, table_q, table_k, table_v, plan):
    return WindowAttentionPlan.apply(q, k, v, table_q, table_k, table_v, plan)


class LayerNormShortRows(Function):
    """F.layer_norm(x, (C,), weight, bias, eps) for fp32 rows of at most 384 elements (include/stb200.h: stb200_layer_norm_*)."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x, weight, bias, eps):
        shape = x.shape
        C = shape[-1]
        x2 = x.reshape(-1, C).contiguous()
        N = x2.shape[0]
        y = torch.empty_like(x2)
        mean = torch.empty(N, dtype=torch.float32, device=x.device)
        rstd = torch.empty(N, dtype=torch.float32, device=x.device)
        if N:
            _cabi.call("stb200_layer_norm_forward", N, C, float(eps), x2.data_ptr(), None if weight is None else weight.contiguous().data_ptr(),
                       None if bias is None else bias.contiguous().data_ptr(), y.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                       torch.cuda.current_stream().cuda_stream)
        ctx.save_for_backward(x2, weight, mean, rstd)
        ctx.has_bias = bias is not None
        ctx.shape = shape
        return y.view(shape)

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, gy):
        x2, weight, mean, rstd = ctx.saved_tensors
        N, C = x2.shape
        gy2 = gy.reshape(-1, C).float().contiguous()
        gx = torch.empty_like(x2)
        want_affine = weight is not None and (ctx.needs_input_grad[1] or ctx.needs_input_grad[2])
        partial = None
        if N:
            if want_affine:
                partial = torch.empty(_cabi.load().stb200_layer_norm_partial_rows(N, C), 2 * C, dtype=torch.float32, device=gx.device)
            _cabi.call("stb200_layer_norm_backward", N, C, gy2.data_ptr(), x2.data_ptr(), None if weight is None else weight.contiguous().data_ptr(),
                       mean.data_ptr(), rstd.data_ptr(), gx.data_ptr(), None if partial is None else partial.data_ptr(),
                       torch.cuda.current_stream().cuda_stream)
        gw = gb = None
        if partial is not None:
            sums = partial.sum(0)
            gw, gb = sums[:C], (sums[C:] if ctx.has_bias else None)
        return gx.view(ctx.shape), gw, gb, None


def layer_norm(x, weight, bias, eps=1e-5):
    return LayerNormShortRows.apply(x, weight, bias, eps)


class KPConvWeighted(Function):
    """weighted[i, k, :] = sum_j max(0, 1 - |s_xyz[nbr[i,j]] - q_xyz[i] - K_k| / extent) * feats[nbr[i,j], :] (include/stb200.h:
    stb200_kpconv_weighted); gradient only with respect to feats."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, q_xyz, s_xyz, nbr, kpts, feats, extent):
        q_xyz, s_xyz, kpts, feats = (t.contiguous() for t in (q_xyz, s_xyz, kpts, feats))
        nbr = nbr.long().contiguous()
        n, nn, K, C = q_xyz.shape[0], nbr.shape[1], kpts.shape[0], feats.shape[1]
        weighted = torch.empty(n, K, C, dtype=torch.float32, device=feats.device)
        _cabi.call("stb200_kpconv_weighted", n, s_xyz.shape[0], nn, K, C, float(extent), q_xyz.data_ptr(), s_xyz.data_ptr(), nbr.data_ptr(),
                   kpts.data_ptr(), feats.data_ptr(), weighted.data_ptr(), torch.cuda.current_stream().cuda_stream)
        ctx.save_for_backward(q_xyz, s_xyz, nbr, kpts)
        ctx.extent, ctx.feat_shape = float(extent), feats.shape
        return weighted

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, gw):
        if not ctx.needs_input_grad[4]:
            return None, None, None, None, None, None
        q_xyz, s_xyz, nbr, kpts = ctx.saved_tensors
        gw = gw.float().contiguous()
        n, nn, K, C = q_xyz.shape[0], nbr.shape[1], kpts.shape[0], ctx.feat_shape[1]
        gf = torch.zeros(ctx.feat_shape, dtype=torch.float32, device=gw.device)
        _cabi.call("stb200_kpconv_weighted_backward", n, s_xyz.shape[0], nn, K, C, ctx.extent, q_xyz.data_ptr(), s_xyz.data_ptr(),
                   nbr.data_ptr(), kpts.data_ptr(), gw.data_ptr(), gf.data_ptr(), torch.cuda.current_stream().cuda_stream)
        return None, None, None, None, gf, None


def kpconv_weighted(q_xyz, s_xyz, nbr, kpts, feats, extent):
    return KPConvWeighted.apply(q_xyz, s_xyz, nbr, kpts, feats, extent)


_QKV_DTYPES = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}


class SplitQKV(Function):
    """qkv [N, 3C] (fp32 / bf16 / fp16: the projection GEMM's output without bias) + bias [3C] -> q, k, v fp32 [N, h, C/h]
    contiguous, in one kernel; backward: one kernel writes grad_qkv in qkv's dtype and the bias gradient.  Replaces the
    reshape / permute / contiguous / multiply of model/stratified_transformer.py:172-175 and the `.float()` casts at the
    pointops call sites (the caller folds `scale` into the q rows of weight and bias)."""

    @staticmethod
    def forward(ctx, qkv, bias, h):
        qkv = qkv.contiguous()
        N, C3 = qkv.shape
        C = C3 // 3
        if qkv.dtype not in _QKV_DTYPES or C3 != 3 * C or C % 8 or C % h:
            raise ValueError(f"SplitQKV: unsupported qkv {tuple(qkv.shape)} {qkv.dtype} with {h} heads")
        if bias is not None:
            bias = bias.float().contiguous()
        q, k, v = (torch.empty(N, h, C // h, dtype=torch.float32, device=qkv.device) for _ in range(3))
        _cabi.call("stb200_qkv_split", N, C, _QKV_DTYPES[qkv.dtype], qkv.data_ptr(), None if bias is None else bias.data_ptr(),
                   q.data_ptr(), k.data_ptr(), v.data_ptr(), torch.cuda.current_stream().cuda_stream)
        ctx.dtype, ctx.has_bias, ctx.shape = qkv.dtype, bias is not None, (N, C)
        return q, k, v

    @staticmethod
    def backward(ctx, gq, gk, gv):
        N, C = ctx.shape
        gq, gk, gv = (g.float().contiguous() for g in (gq, gk, gv))
        g_qkv = torch.empty(N, 3 * C, dtype=ctx.dtype, device=gq.device)
        partial = None
        if ctx.has_bias and ctx.needs_input_grad[1]:
            partial = torch.empty(_cabi.load().stb200_qkv_partial_rows(N, C), 3 * C, dtype=torch.float32, device=gq.device)
        _cabi.call("stb200_qkv_merge", N, C, _QKV_DTYPES[ctx.dtype], gq.data_ptr(), gk.data_ptr(), gv.data_ptr(), g_qkv.data_ptr(),
                   None if partial is None else partial.data_ptr(), torch.cuda.current_stream().cuda_stream)
        return g_qkv, None if partial is None else partial.sum(0), None


def split_qkv(qkv, bias, num_heads):
    return SplitQKV.apply(qkv, bias, num_heads)


@torch.no_grad()
def window_attention_inference_bf16(q, k, v, table_q, table_k, table_v, pair_index, pre_cast=False):
    """Forward-only pair path with bf16 storage of q / k / v and of the staged tables (BASELINE config 3, inference):
    logits and softmax in fp32, output fp32 [N, h, d].  Tolerance vs the fp32 path: 2e-2 of the output scale.
    pre_cast: q / k / v are already contiguous bf16 tensors (an inference stack keeps them in bf16 from the qkv GEMM)."""
    N, h, d = q.shape
    L = table_q.shape[0]
    q16, k16, v16 = (q, k, v) if pre_cast else (t.to(torch.bfloat16).contiguous() for t in (q, k, v))
    tq, tk, tv = (t.float().contiguous() for t in (table_q, table_k, table_v))
    ix = pair_index.c_struct(L)
    stream = torch.cuda.current_stream().cuda_stream
    s = torch.empty(pair_index.M, h, dtype=torch.float32, device=q.device)
    _cabi.call("stb200_window_logits_forward_bf16", ctypes.byref(ix), h, d, L, q16.data_ptr(), k16.data_ptr(), tq.data_ptr(),
               tk.data_ptr(), s.data_ptr(), stream)
    p = torch.empty_like(s)
    pointops_cuda.segment_softmax_forward_cuda(N, pair_index.M, h, s, None, pair_index.index_0_offsets, p)
    out = torch.empty(N, h, d, dtype=torch.float32, device=q.device)
    _cabi.call("stb200_window_aggregate_forward_bf16", ctypes.byref(ix), h, d, L, p.data_ptr(), v16.data_ptr(), tv.data_ptr(),
               out.data_ptr(), stream)
    return out


def segment_softmax(a, index0_offsets, b=None):
    return SegmentSoftmax.apply(a, b, index0_offsets)


def scatter_softmax(src, index, dim=0):
    """Drop-in for torch_scatter.scatter_softmax as the reference calls it (sorted `index`, dim 0)."""
    if dim != 0 or src.dim() != 2:
        raise NotImplementedError("scatter_softmax shim: only src [M,h], dim=0 (the reference's call) is supported")
    n = int(index[-1].item()) + 1 if index.numel() else 0
    counts = torch.bincount(index, minlength=n)
    offsets = torch.cat([counts.new_zeros(1), counts.cumsum(0)]).int()
    return SegmentSoftmax.apply(src.contiguous(), None, offsets)


# ------------------------------------------------------------------------------------------------ v1 family
class AttentionStep1(Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, q, k, index0, index1):
        _contig(q, k, index0, index1)
        N_q, h, d = q.shape
        M = index0.shape[0]
        out = torch.zeros(M, h, dtype=torch.float32, device=q.device)
        pointops_cuda.attention_step1_forward_cuda(k.shape[0], M, h, h * d, q, k, index0, index1, out)
        ctx.save_for_backward(q, k, index0, index1)
        return out

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_output):
        q, k, index0, index1 = ctx.saved_tensors
        N_q, h, d = q.shape
        grad_q, grad_k = torch.zeros_like(q), torch.zeros_like(k)
        pointops_cuda.attention_step1_backward_cuda(N_q, index0.shape[0], h, h * d, grad_output.contiguous(), index0,
                                                    index1, q, k, grad_q, grad_k)
        return grad_q, grad_k, None, None


attention_step1 = AttentionStep1.apply


class AttentionStep2(Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, attn, v, index0, index1):
        _contig(attn, v, index0, index1)
        M, h = attn.shape
        N_q = int(index0.max().item()) + 1
        _, _, d = v.shape
        out = torch.zeros(N_q, h, d, dtype=torch.float32, device=v.device)
        pointops_cuda.attention_step2_forward_cuda(N_q, M, h, h * d, attn, v, index0, index1, out)
        ctx.save_for_backward(attn, v, index0, index1)
        return out

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_output):
        attn, v, index0, index1 = ctx.saved_tensors
        N_q, h, d = grad_output.shape
        grad_attn, grad_v = torch.zeros_like(attn), torch.zeros_like(v)
        pointops_cuda.attention_step2_backward_cuda(N_q, attn.shape[0], h, h * d, grad_output.contiguous(), index0,
                                                    index1, attn, v, grad_attn, grad_v)
        return grad_attn, grad_v, None, None


attention_step2 = AttentionStep2.apply
attention_step2_v2 = AttentionStep2.apply  # the reference's AttentionStep2_v2 calls the same v1 symbols (pointops.py:284)


class DotProdWithIdx(Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, q, index, table, rel_idx):
        _contig(q, index, table, rel_idx)
        N, h, hdim = q.shape
        M = index.shape[0]
        out = torch.zeros(M, h, dtype=torch.float32, device=q.device)
        pointops_cuda.dot_prod_with_idx_forward_cuda(N, M, h, hdim, q, index, table, rel_idx, out)
        ctx.save_for_backward(q, index, table, rel_idx)
        return out

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_output):
        q, index, table, rel_idx = ctx.saved_tensors
        N, h, hdim = q.shape
        grad_q, grad_table = torch.zeros_like(q), torch.zeros_like(table)
        pointops_cuda.dot_prod_with_idx_backward_cuda(N, index.shape[0], h, hdim, grad_output.contiguous(), q, index,
                                                      table, rel_idx, grad_q, grad_table)
        return grad_q, None, grad_table, None


dot_prod_with_idx = DotProdWithIdx.apply


class DotProdWithIdx_v2(Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, q, index_q, k, index_k, table_q, table_k, rel_idx):
        _contig(q, index_q, k, index_k, table_q, table_k, rel_idx)
        N, h, hdim = q.shape
        M = index_q.shape[0]
        out = torch.zeros(M, h, dtype=torch.float32, device=q.device)
        # the reference pre-sorts pairs by merged rel idx here (pointops.py:386-393); the result does not
        # depend on that order, so the sort is skipped
        pointops_cuda.dot_prod_with_idx_forward_cuda_v2(N, M, h, hdim, 0, 0, q, index_q, k, index_k, table_q,
                                                        table_k, rel_idx, None, None, out)
        ctx.save_for_backward(q, index_q, k, index_k, table_q, table_k, rel_idx)
        return out

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_output):
        q, index_q, k, index_k, table_q, table_k, rel_idx = ctx.saved_tensors
        N, h, hdim = q.shape
        grad_q, grad_k = torch.zeros_like(q), torch.zeros_like(k)
        grad_tq, grad_tk = torch.zeros_like(table_q), torch.zeros_like(table_k)
        pointops_cuda.dot_prod_with_idx_backward_cuda_v2(N, index_q.shape[0], h, hdim, 0, 0, grad_output.contiguous(),
                                                         q, index_q, k, index_k, table_q, table_k, rel_idx, None,
                                                         None, grad_q, grad_k, grad_tq, grad_tk)
        return grad_q, None, grad_k, None, grad_tq, grad_tk, None


dot_prod_with_idx_v2 = DotProdWithIdx_v2.apply


class AttentionStep2WithRelPosValue(Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, attn, v, index0, index1, table, rel_idx):
        _contig(attn, v, index0, index1, table, rel_idx)
        M, h = attn.shape
        _, _, hdim = v.shape
        N_q = int(index0.max().item()) + 1
        out = torch.zeros(N_q, h, hdim, dtype=torch.float32, device=v.device)
        pointops_cuda.attention_step2_with_rel_pos_value_forward_cuda(N_q, M, h, hdim, attn, v, index0, index1,
                                                                      table, rel_idx, out)
        ctx.save_for_backward(attn, v, index0, index1, table, rel_idx)
        return out

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_output):
        attn, v, index0, index1, table, rel_idx = ctx.saved_tensors
        N_q, h, hdim = grad_output.shape
        grad_attn, grad_v, grad_table = torch.zeros_like(attn), torch.zeros_like(v), torch.zeros_like(table)
        pointops_cuda.attention_step2_with_rel_pos_value_backward_cuda(N_q, attn.shape[0], h, hdim,
                                                                       grad_output.contiguous(), index0, index1, attn,
                                                                       v, table, rel_idx, grad_attn, grad_v,
                                                                       grad_table)
        return grad_attn, grad_v, None, None, grad_table, None


attention_step2_with_rel_pos_value = AttentionStep2WithRelPosValue.apply


# ------------------------------------------------------------------------------------------------ neighbourhood ops
# (what TransitionDown / Upsample call, model/stratified_transformer.py:103-106, 341)
class KNNQuery(Function):
    @staticmethod
    def forward(ctx, nsample, xyz, new_xyz, offset, new_offset):
        """xyz (n,3), new_xyz (m,3) or None, offset (b), new_offset (b) -> idx (m,nsample) i32, dist (m,nsample) f32
        (Euclidean distance, i.e. sqrt of the kernel's squared distance: pointops.py:34-49)."""
        if new_xyz is None:
            new_xyz = xyz
        _contig(xyz, new_xyz)
        m = new_xyz.shape[0]
        idx = torch.empty(m, nsample, dtype=torch.int32, device=xyz.device)
        dist2 = torch.empty(m, nsample, dtype=torch.float32, device=xyz.device)
        pointops_cuda.knnquery_cuda(m, nsample, xyz, new_xyz, offset.int().contiguous(), new_offset.int().contiguous(), idx, dist2)
        ctx.mark_non_differentiable(idx)
        return idx, torch.sqrt(dist2)

    @staticmethod
    def backward(ctx, *grads):
        return None, None, None, None, None


knnquery = KNNQuery.apply


def queryandgroup(nsample, xyz, new_xyz, feat, idx, offset, new_offset, use_xyz=True, return_indx=False):
    """pointops.py:648-675: kNN + gather of relative coordinates and features -> (m, nsample, [3+]c)."""
    _contig(xyz, feat)
    if new_xyz is None:
        new_xyz = xyz
    if idx is None:
        idx, _ = knnquery(nsample, xyz, new_xyz, offset, new_offset)
    m, c = new_xyz.shape[0], feat.shape[1]
    flat = idx.view(-1).long()
    grouped_xyz = xyz[flat, :].view(m, nsample, 3) - new_xyz.unsqueeze(1)
    grouped_feat = feat[flat, :].view(m, nsample, c)
    out = torch.cat((grouped_xyz, grouped_feat), -1) if use_xyz else grouped_feat
    return (out, idx) if return_indx else out


def interpolation(xyz, new_xyz, feat, offset, new_offset, k=3):
    """pointops.py:756-770: inverse-distance weighted interpolation from the k nearest support points -> (n, c)."""
    _contig(xyz, new_xyz, feat)
    idx, dist = knnquery(k, xyz, new_xyz, offset, new_offset)
    dist_recip = 1.0 / (dist + 1e-8)
    weight = dist_recip / torch.sum(dist_recip, dim=1, keepdim=True)
    new_feat = torch.zeros(new_xyz.shape[0], feat.shape[1], dtype=feat.dtype, device=feat.device)
    for i in range(k):
        new_feat = new_feat + feat[idx[:, i].long(), :] * weight[:, i].unsqueeze(-1)
    return new_feat
