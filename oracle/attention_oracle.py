"""CPU oracle: attention math of the window-attention hot path (TEST INFRASTRUCTURE).

Plain gather / index_add restatement on CPU torch tensors.  Every function works
in the dtype of its float inputs: pass float64 copies for the high-precision
oracle, float32 for the "same arithmetic type" baseline.

Reference it follows (paths under /root/reference):
  * step1            lib/pointops2/src/attention_v2/attention_cuda_kernel_v2.cu:39-45, 78-90
                     (v1: lib/pointops2/src/attention/attention_cuda_kernel.cu:16-19, 32-37)
  * rel-pos q/k dot  lib/pointops2/src/rpe_v2/relative_pos_encoding_cuda_kernel_v2.cu:272-281, 316-339
                     (v1 single table: lib/pointops2/src/rpe/relative_pos_encoding_cuda_kernel.cu:19-24, 39-48)
  * softmax          torch_scatter.scatter_softmax as called at model/stratified_transformer.py:205
                     (third-party, not vendored: mathematical definition, "parity unpinned")
  * step2 + value    lib/pointops2/src/rpe_v2/relative_pos_encoding_cuda_kernel_v2.cu:422-437, 466-483
                     authors' own torch form: lib/pointops2/functions/test_relative_pos_encoding_op_step2.py:32-38
  * plain step2      lib/pointops2/src/attention/attention_cuda_kernel.cu:67-70, 82-86

Pinning: the reference stores no golden vectors for these ops (its test scripts are
seeded generators that need a GPU).  This oracle is pinned two ways:
  1. tests/test_oracle_attention.py: explicit backward formulas == torch autograd of the
     authors' forward formulation, and a literal loop-by-loop emulation on small cases.
  2. tests/test_gpu_parity.py (-m gpu): against the reference's own kernels compiled from
     /root/reference into oracle/_ref/libpointops2_ref.so, on the reference's seeded shapes.
"""
from __future__ import annotations

import torch

_CHUNK = 1 << 18  # pairs per chunk, bounds the [chunk, h, d] temporaries


def index0_from_offsets(offsets: torch.Tensor) -> torch.Tensor:
    """Expand CSR offsets [N+1] into the per-pair query id [M] (model/stratified_transformer.py:312-317 inverse)."""
    offsets = offsets.long()
    counts = offsets[1:] - offsets[:-1]
    return torch.repeat_interleave(torch.arange(counts.numel(), dtype=torch.long), counts)


def table_sum(table: torch.Tensor, rel_idx: torch.Tensor) -> torch.Tensor:
    """E[m,h,c] = T[r0,h,c,0] + T[r1,h,c,1] + T[r2,h,c,2], left-to-right adds (rpe_v2 kernel :276)."""
    r = rel_idx.long()
    return (table[r[:, 0], :, :, 0] + table[r[:, 1], :, :, 1]) + table[r[:, 2], :, :, 2]


def _chunks(M):
    for s in range(0, M, _CHUNK):
        yield s, min(M, s + _CHUNK)


# ----------------------------------------------------------------------------- step1
def step1_fwd(q, k, i0, i1):
    i0, i1 = i0.long(), i1.long()
    M, h = i0.numel(), q.shape[1]
    out = q.new_zeros(M, h)
    for s, e in _chunks(M):
        out[s:e] = (q[i0[s:e]] * k[i1[s:e]]).sum(-1)
    return out


def step1_bwd(g, q, k, i0, i1):
    """grad_q[n] = sum g*k[i1]; grad_k[i1] += g*q[i0]."""
    i0, i1 = i0.long(), i1.long()
    gq, gk = torch.zeros_like(q), torch.zeros_like(k)
    for s, e in _chunks(i0.numel()):
        gg = g[s:e].unsqueeze(-1)
        gq.index_add_(0, i0[s:e], gg * k[i1[s:e]])
        gk.index_add_(0, i1[s:e], gg * q[i0[s:e]])
    return gq, gk


# ----------------------------------------------------------------------------- rel-pos bias
def rpe_fwd(q, k, i0, i1, table_q, table_k, rel_idx):
    """b[m,h] = <q[i0],Eq> + <k[i1],Ek>  (dot_prod_with_idx v2/v3)."""
    i0, i1 = i0.long(), i1.long()
    M, h = i0.numel(), q.shape[1]
    out = q.new_zeros(M, h)
    for s, e in _chunks(M):
        eq = table_sum(table_q, rel_idx[s:e])
        ek = table_sum(table_k, rel_idx[s:e])
        out[s:e] = (q[i0[s:e]] * eq + k[i1[s:e]] * ek).sum(-1)
    return out


def _table_grad_add(gt, rel_idx, contrib):
    """gt[r_a,h,c,a] += contrib[m,h,c] for a in 0..2."""
    r = rel_idx.long()
    for a in range(3):
        tmp = torch.zeros_like(gt[..., a])
        tmp.index_add_(0, r[:, a], contrib)
        gt[..., a] += tmp


def rpe_bwd(g, q, k, i0, i1, table_q, table_k, rel_idx):
    i0, i1 = i0.long(), i1.long()
    gq, gk = torch.zeros_like(q), torch.zeros_like(k)
    gtq, gtk = torch.zeros_like(table_q), torch.zeros_like(table_k)
    for s, e in _chunks(i0.numel()):
        gg = g[s:e].unsqueeze(-1)
        r = rel_idx[s:e]
        gq.index_add_(0, i0[s:e], gg * table_sum(table_q, r))
        gk.index_add_(0, i1[s:e], gg * table_sum(table_k, r))
        _table_grad_add(gtq, r, gg * q[i0[s:e]])
        _table_grad_add(gtk, r, gg * k[i1[s:e]])
    return gq, gk, gtq, gtk


def rpe_single_fwd(x, index, table, rel_idx):
    """v1 dot_prod_with_idx: out[m,h] = <x[index[m]], E(m)>  (rpe kernel :19-24)."""
    index = index.long()
    out = x.new_zeros(index.numel(), x.shape[1])
    for s, e in _chunks(index.numel()):
        out[s:e] = (x[index[s:e]] * table_sum(table, rel_idx[s:e])).sum(-1)
    return out


def rpe_single_bwd(g, x, index, table, rel_idx):
    index = index.long()
    gx, gt = torch.zeros_like(x), torch.zeros_like(table)
    for s, e in _chunks(index.numel()):
        gg = g[s:e].unsqueeze(-1)
        gx.index_add_(0, index[s:e], gg * table_sum(table, rel_idx[s:e]))
        _table_grad_add(gt, rel_idx[s:e], gg * x[index[s:e]])
    return gx, gt


# ----------------------------------------------------------------------------- segment softmax
def softmax_fwd(s, i0, N):
    """p = exp(s - segmax) / segsum over pairs sharing i0, per head (scatter_softmax, dim 0)."""
    i0 = i0.long()
    h = s.shape[1]
    idx = i0.unsqueeze(-1).expand(-1, h)
    mx = s.new_full((N, h), float("-inf")).scatter_reduce(0, idx, s, "amax", include_self=True)
    ex = torch.exp(s - mx[i0])
    den = s.new_zeros(N, h).index_add_(0, i0, ex)
    return ex / den[i0]


def softmax_bwd(p, gp, i0, N):
    """gs = p * (gp - sum_seg p*gp)."""
    i0 = i0.long()
    dot = p.new_zeros(N, p.shape[1]).index_add_(0, i0, p * gp)
    return p * (gp - dot[i0])


# ----------------------------------------------------------------------------- step2
def step2_fwd(p, v, i0, i1, N):
    i0, i1 = i0.long(), i1.long()
    out = v.new_zeros(N, v.shape[1], v.shape[2])
    for s, e in _chunks(i0.numel()):
        out.index_add_(0, i0[s:e], p[s:e].unsqueeze(-1) * v[i1[s:e]])
    return out


def step2_bwd(g, p, v, i0, i1):
    i0, i1 = i0.long(), i1.long()
    gp, gv = torch.zeros_like(p), torch.zeros_like(v)
    for s, e in _chunks(i0.numel()):
        go = g[i0[s:e]]
        gp[s:e] = (go * v[i1[s:e]]).sum(-1)
        gv.index_add_(0, i1[s:e], p[s:e].unsqueeze(-1) * go)
    return gp, gv


def step2_rpv_fwd(p, v, i0, i1, table_v, rel_idx, N):
    """out[n,h,:] = sum_seg p*(Ev + v[i1])  (rpe_v2 kernel :428-430: (table + value) * attn)."""
    i0, i1 = i0.long(), i1.long()
    out = v.new_zeros(N, v.shape[1], v.shape[2])
    for s, e in _chunks(i0.numel()):
        val = table_sum(table_v, rel_idx[s:e]) + v[i1[s:e]]
        out.index_add_(0, i0[s:e], val * p[s:e].unsqueeze(-1))
    return out


def step2_rpv_bwd(g, p, v, i0, i1, table_v, rel_idx):
    i0, i1 = i0.long(), i1.long()
    gp, gv, gt = torch.zeros_like(p), torch.zeros_like(v), torch.zeros_like(table_v)
    for s, e in _chunks(i0.numel()):
        go = g[i0[s:e]]
        val = table_sum(table_v, rel_idx[s:e]) + v[i1[s:e]]
        gp[s:e] = (val * go).sum(-1)
        contrib = p[s:e].unsqueeze(-1) * go
        gv.index_add_(0, i1[s:e], contrib)
        _table_grad_add(gt, rel_idx[s:e], contrib)
    return gp, gv, gt


# ----------------------------------------------------------------------------- whole layer
def layer_fwd(q, k, v, offsets, i1, table_q, table_k, table_v, rel_idx):
    """Forward of the pair path of WindowAttention.forward (model/stratified_transformer.py:183-208)."""
    N = q.shape[0]
    i0 = index0_from_offsets(offsets)
    a = step1_fwd(q, k, i0, i1)
    b = rpe_fwd(q, k, i0, i1, table_q, table_k, rel_idx)
    s = a + b
    p = softmax_fwd(s, i0, N)
    out = step2_rpv_fwd(p, v, i0, i1, table_v, rel_idx, N)
    return dict(i0=i0, a=a, b=b, s=s, p=p, out=out)


def layer_fwd_bwd(q, k, v, offsets, i1, table_q, table_k, table_v, rel_idx, g_out):
    """Forward + explicit backward; returns every intermediate and gradient."""
    N = q.shape[0]
    f = layer_fwd(q, k, v, offsets, i1, table_q, table_k, table_v, rel_idx)
    i0, p = f["i0"], f["p"]
    gp, gv, gtv = step2_rpv_bwd(g_out, p, v, i0, i1, table_v, rel_idx)
    gs = softmax_bwd(p, gp, i0, N)
    gq1, gk1 = step1_bwd(gs, q, k, i0, i1)
    gq2, gk2, gtq, gtk = rpe_bwd(gs, q, k, i0, i1, table_q, table_k, rel_idx)
    f.update(gp=gp, gs=gs, gq=gq1 + gq2, gk=gk1 + gk2, gv=gv, gtq=gtq, gtk=gtk, gtv=gtv,
             gq_step1=gq1, gk_step1=gk1, gq_rpe=gq2, gk_rpe=gk2)
    return f


def layer_autograd(q, k, v, offsets, i1, table_q, table_k, table_v, rel_idx, g_out):
    """Authors' pure-torch formulation differentiated by autograd: the CPU baseline that
    BASELINE.json's north_star asks to time (gather + scatter_add, autograd backward)."""
    N = q.shape[0]
    i0 = index0_from_offsets(offsets)
    i1 = i1.long()
    r = rel_idx.long()
    leaves = [t.detach().clone().requires_grad_(True) for t in (q, k, v, table_q, table_k, table_v)]
    q_, k_, v_, tq, tk, tv = leaves
    qf, kf = q_[i0], k_[i1]
    eq = tq[r[:, 0], :, :, 0] + tq[r[:, 1], :, :, 1] + tq[r[:, 2], :, :, 2]
    ek = tk[r[:, 0], :, :, 0] + tk[r[:, 1], :, :, 1] + tk[r[:, 2], :, :, 2]
    s = (qf * kf).sum(-1) + (qf * eq + kf * ek).sum(-1)
    h = s.shape[1]
    idx = i0.unsqueeze(-1).expand(-1, h)
    mx = s.detach().new_full((N, h), float("-inf")).scatter_reduce(0, idx, s.detach(), "amax")
    ex = torch.exp(s - mx[i0])
    den = torch.zeros(N, h, dtype=s.dtype).index_add(0, i0, ex)
    p = ex / den[i0]
    ev = tv[r[:, 0], :, :, 0] + tv[r[:, 1], :, :, 1] + tv[r[:, 2], :, :, 2]
    out = torch.zeros(N, h, v.shape[2], dtype=s.dtype).index_add(0, i0, p.unsqueeze(-1) * (v_[i1] + ev))
    out.backward(g_out)
    return dict(out=out.detach(), p=p.detach(), s=s.detach(), gq=q_.grad, gk=k_.grad, gv=v_.grad,
                gtq=tq.grad, gtk=tk.grad, gtv=tv.grad)


# ----------------------------------------------------------------------------- literal loops (small cases only)
def layer_loops(q, k, v, offsets, i1, table_q, table_k, table_v, rel_idx):
    """Loop-by-loop emulation of the v2/v3 kernels' forward (one (query, head) CTA, one pair per
    thread, channel loop with the interleaved accumulator of rpe_v2 kernel :275-280)."""
    import numpy as np
    q, k, v = (t.numpy() for t in (q, k, v))
    tq, tk, tv = (t.numpy() for t in (table_q, table_k, table_v))
    off, i1, r = offsets.numpy(), i1.numpy(), rel_idx.numpy()
    N, h, d = q.shape
    M = i1.shape[0]
    dt = q.dtype.type
    a = np.zeros((M, h), q.dtype); b = np.zeros((M, h), q.dtype); p = np.zeros((M, h), q.dtype)
    out = np.zeros((N, h, d), q.dtype)
    for n in range(N):
        for hh in range(h):
            seg = range(off[n], off[n + 1])
            for m in seg:
                sa = dt(0); sb = dt(0)
                for c in range(d):
                    sa = dt(sa + q[n, hh, c] * k[i1[m], hh, c])
                    eq = dt(dt(tq[r[m, 0], hh, c, 0] + tq[r[m, 1], hh, c, 1]) + tq[r[m, 2], hh, c, 2])
                    sb = dt(sb + q[n, hh, c] * eq)
                    ek = dt(dt(tk[r[m, 0], hh, c, 0] + tk[r[m, 1], hh, c, 1]) + tk[r[m, 2], hh, c, 2])
                    sb = dt(sb + k[i1[m], hh, c] * ek)
                a[m, hh] = sa; b[m, hh] = sb
            if len(seg) == 0:
                continue
            s = a[off[n]:off[n + 1], hh] + b[off[n]:off[n + 1], hh]
            ex = np.exp(s - s.max())
            p[off[n]:off[n + 1], hh] = ex / ex.sum()
            for m in seg:
                for c in range(d):
                    ev = dt(dt(tv[r[m, 0], hh, c, 0] + tv[r[m, 1], hh, c, 1]) + tv[r[m, 2], hh, c, 2])
                    out[n, hh, c] += dt(dt(ev + v[i1[m], hh, c]) * p[m, hh])
    return dict(a=torch.from_numpy(a), b=torch.from_numpy(b), p=torch.from_numpy(p), out=torch.from_numpy(out))
