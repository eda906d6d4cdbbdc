"""The callers of the hot path, one level up: drop-in mirrors of the reference's `Mlp`, `SwinTransformerBlock`, `BasicLayer`,
`TransitionDown` and `Upsample` (/root/reference/model/stratified_transformer.py:67-112, 219-342).

Same constructor arguments, parameter names and shapes (reference checkpoints load with `load_state_dict`), same call
forms and return values.  What differs is what runs underneath: `BasicLayer.forward` builds the layer's pair lists ONCE on
the device (`index.build_layer_index`: FPS + both block parities, reference lines 267-317 which re-derive them in Python
per block) and hands each block a `PairIndex`; the blocks run `WindowAttention` on libstb200.  tests/test_gpu_layers.py
runs the reference's own classes (text extracted at build time) over this package's operators beside these mirrors.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import index as st_index
from . import pointops
from .window_attention import WindowAttention


class LayerNorm(nn.LayerNorm):
    """nn.LayerNorm with the same parameters; fp32 CUDA inputs normalised over a last dimension of at most 384 elements go through
    the short-row kernels of libstb200 (torch's kernel spends 28 ms per step of the full model on these 48..384-wide rows)."""

    def forward(self, x):
        if x.is_cuda and len(self.normalized_shape) == 1 and self.normalized_shape[0] <= 384 and x.dtype in (torch.float32, torch.bfloat16, torch.float16):
            return pointops.layer_norm(x, self.weight, self.bias, self.eps)
        return super().forward(x)


class DropPath(nn.Module):
    """Stochastic depth per sample (the first dimension), scaled by the keep probability - what the reference imports from
    timm (`timm.models.layers.DropPath`, stratified_transformer.py:5); identity when not training or drop_prob == 0."""

    def __init__(self, drop_prob: float = 0.0):
        super().__init__()
        self.drop_prob = float(drop_prob)

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
        return x * mask.div_(keep)


class Mlp(nn.Module):
    """stratified_transformer.py:67-85"""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop, inplace=True)

    def forward(self, x):
        return self.drop(self.fc2(self.drop(self.act(self.fc1(x)))))


def transition_down_offsets(offset, ratio):
    """Cumulative sample counts of TransitionDown exactly as stratified_transformer.py:98-102 computes them: the first scene
    contributes int(n_0 * ratio) + 1, every later scene adds the UNTRUNCATED n_i * ratio + 1 to a running Python float, and the
    list is truncated only when it becomes an IntTensor - so fractions carry over between scenes."""
    off = [int(v) for v in (offset.tolist() if isinstance(offset, torch.Tensor) else offset)]
    count = int(off[0] * ratio) + 1
    out = [count]
    for i in range(1, len(off)):
        count += ((off[i] - off[i - 1]) * ratio) + 1
        out.append(count)
    return [int(c) for c in out]      # torch.cuda.IntTensor(list of floats) truncates toward zero


class TransitionDown(nn.Module):
    """stratified_transformer.py:87-112: FPS to ratio * n points, kNN grouping, LayerNorm + Linear per neighbour, max pool."""

    def __init__(self, in_channels, out_channels, ratio, k, norm_layer=LayerNorm):
        super().__init__()
        self.ratio = ratio
        self.k = k
        self.norm = norm_layer(in_channels) if norm_layer else None
        self.linear = nn.Linear(in_channels, out_channels, bias=False)
        self.pool = nn.MaxPool1d(k)

    def sample_offsets(self, offset, device):
        return torch.tensor(transition_down_offsets(offset, self.ratio), dtype=torch.int32, device=device)

    def forward(self, feats, xyz, offset, fps=None):
        """fps = (idx, n_offset): furthest-point picks the caller already computed for exactly these scenes and counts"""
        if fps is None:
            n_offset = self.sample_offsets(offset, xyz.device)
            idx = pointops.furthestsampling(xyz, offset.int(), n_offset)
        else:
            idx, n_offset = fps
        n_xyz = xyz[idx.long(), :]
        grouped = pointops.queryandgroup(self.k, xyz, n_xyz, feats, None, offset.int(), n_offset, use_xyz=False)   # (m, k, c)
        m, k, c = grouped.shape
        if self.norm is not None:
            grouped = self.norm(grouped)
        # max over the k neighbours = `self.pool(x.transpose(1, 2).contiguous()).squeeze(-1)` of the reference (line 109) without the
        # transposed copy; the gradient goes to the arg-max either way (MaxPool1d's backward kernel alone took 10 ms per step of
        # the full model on 8 x 80k points, this one 1 ms)
        pooled = self.linear(grouped).max(dim=1).values if k == self.k else self.pool(self.linear(grouped).transpose(1, 2).contiguous()).squeeze(-1)
        return pooled, n_xyz, n_offset


class Upsample(nn.Module):
    """stratified_transformer.py:328-342"""

    def __init__(self, k, in_channels, out_channels, bn_momentum=0.02):
        super().__init__()
        self.k = k
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.linear1 = nn.Sequential(LayerNorm(out_channels), nn.Linear(out_channels, out_channels))
        self.linear2 = nn.Sequential(LayerNorm(in_channels), nn.Linear(in_channels, out_channels))

    def forward(self, feats, xyz, support_xyz, offset, support_offset, support_feats=None):
        up = pointops.interpolation(xyz, support_xyz, self.linear2(feats), offset, support_offset)
        return self.linear1(support_feats) + up, support_xyz, support_offset


class SwinTransformerBlock(nn.Module):
    """stratified_transformer.py:219-248.  forward(feats, xyz, pair_index) with a prebuilt PairIndex, or the reference's
    forward(feats, xyz, index_0, index_1, index_0_offsets, n_max)."""

    def __init__(self, dim, num_heads, window_size, quant_size, rel_query=True, rel_key=False, rel_value=False, drop_path=0.0,
                 mlp_ratio=4.0, qkv_bias=True, qk_scale=None, act_layer=nn.GELU, norm_layer=LayerNorm, mode=4):
        super().__init__()
        self.mode = mode
        self.norm1 = norm_layer(dim)
        self.attn = WindowAttention(dim, window_size, num_heads=num_heads, quant_size=quant_size, rel_query=rel_query,
                                    rel_key=rel_key, rel_value=rel_value, qkv_bias=qkv_bias, qk_scale=qk_scale)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer)

    def forward(self, feats, xyz, index_0, index_1=None, index_0_offsets=None, n_max=None):
        attn = self.attn(self.norm1(feats), xyz, index_0, index_1, index_0_offsets, n_max)
        feats = feats + self.drop_path(attn)
        return feats + self.drop_path(self.mlp(self.norm2(feats)))


class BasicLayer(nn.Module):
    """stratified_transformer.py:250-326.  forward(feats, xyz, offset) -> (feats, xyz, offset, feats_down, xyz_down, offset_down).

    `layer_index` may carry a prebuilt `index.LayerIndex` (e.g. from `index.GeometryPrefetcher`, computed on a side stream
    during the previous step); otherwise it is built here, once for all blocks.  `fused = True` asks for the window-centric
    plan instead of the per-op pair list."""

    def __init__(self, downsample_scale, depth, channel, num_heads, window_size, grid_size, quant_size, rel_query=True,
                 rel_key=False, rel_value=False, drop_path=0.0, mlp_ratio=4.0, qkv_bias=True, qk_scale=None,
                 norm_layer=LayerNorm, downsample=None, ratio=0.25, k=16, out_channels=None):
        super().__init__()
        self.depth = depth
        self.grid_size = grid_size
        self.max_window_counts = 64
        self.window_size = window_size
        self.quant_size = quant_size
        self.downsample_scale = downsample_scale
        self.fused = False
        self.blocks = nn.ModuleList([
            SwinTransformerBlock(channel, num_heads, window_size, quant_size, rel_query=rel_query, rel_key=rel_key,
                                 rel_value=rel_value, drop_path=drop_path[i] if isinstance(drop_path, list) else drop_path,
                                 mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale, norm_layer=norm_layer)
            for i in range(depth)])
        self.downsample = downsample(channel, out_channels, ratio, k) if downsample else None

    def forward(self, feats, xyz, offset, layer_index=None, fps=None):
        """fps = (idx, n_offset): TransitionDown's furthest-point picks when the caller computed them together with `layer_index`"""
        li = layer_index
        if li is None:
            ds_idx = None
            if isinstance(self.downsample, TransitionDown) and self.downsample_scale is not None:
                # The reference samples twice per layer from the same points: n // ds + 1 key candidates here (line 289) and
                # ratio * n + 1 points in TransitionDown (line 103).  FPS is greedy from point 0 of each scene, so the shorter
                # list is a prefix of the longer one: ONE run serves both (a third of the layer's FPS iterations saved).
                n_offset = self.downsample.sample_offsets(offset, xyz.device)
                k_offset = st_index.fps_new_offset(offset.to(xyz.device), self.downsample_scale)
                if bool((torch.diff(k_offset, prepend=k_offset.new_zeros(1)) <= torch.diff(n_offset, prepend=n_offset.new_zeros(1))).all()):
                    fps = (pointops.furthestsampling(xyz, offset.int(), n_offset), n_offset)
                    ds_idx = st_index.fps_prefix(fps[0], n_offset, k_offset)
            li = st_index.build_layer_index(xyz, offset, self.window_size, self.quant_size, self.downsample_scale, fused=self.fused,
                                            downsample_idx=ds_idx)
        for i, blk in enumerate(self.blocks):
            feats = blk(feats, xyz, li.for_block(i))
        if self.downsample:
            feats_down, xyz_down, offset_down = self.downsample(feats, xyz, offset, fps) if fps is not None else \
                self.downsample(feats, xyz, offset)
        else:
            feats_down, xyz_down, offset_down = None, None, None
        return feats, xyz, offset, feats_down, xyz_down, offset_down
