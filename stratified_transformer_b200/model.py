"""The whole network around the hot path: drop-in mirrors of the reference's `KPConvSimpleBlock`, `KPConvResBlock`
and `Stratified` (/root/reference/model/stratified_transformer.py:344-505), assembled from `layers.py` (BasicLayer /
TransitionDown / Upsample on libstb200) with the reference's constructor arguments, attribute names and return values, so a
reference checkpoint's `state_dict` loads.

`KPConvLayer` and `FastBatchNorm1d` are torch_points3d classes (third party, not vendored: PARITY UNPINNED).  They are restated
here from the published algorithm (rigid KPConv, linear influence, sum aggregation; Thomas et al. 2019 / torch_points3d
`KPConv/kernels.py`, `convolution_ops.py`) with the library's parameter names (`K_points`, `weight`, `batch_norm.*`).  The
kernel-point disposition the library loads from its optimised tables is replaced by a fixed symmetric one for fresh models;
checkpoints bring their own `K_points`.  The stem is outside the hot path; its neighbourhood sums run in one kernel
(`pointops.kpconv_weighted`), the published formula as torch operators stays as `KPConvLayer.forward_torch`.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
from torch.nn.init import trunc_normal_

from . import index as st_index
from . import pointops
from . import prestep
from .layers import BasicLayer, TransitionDown, Upsample, transition_down_offsets


def default_kernel_points(radius: float, n: int = 15) -> torch.Tensor:
    """A fixed symmetric disposition: the centre, the 6 axis directions and the 8 cube diagonals, at 2/3 of the kernel radius
    (torch_points3d instead loads dispositions optimised offline; any checkpoint overrides these values)."""
    dirs = [(0.0, 0.0, 0.0)]
    for a in range(3):
        for s in (1.0, -1.0):
            d = [0.0, 0.0, 0.0]
            d[a] = s
            dirs.append(tuple(d))
    c = 1.0 / math.sqrt(3.0)
    for sx in (c, -c):
        for sy in (c, -c):
            for sz in (c, -c):
                dirs.append((sx, sy, sz))
    pts = torch.tensor(dirs[:n], dtype=torch.float32)
    if n > len(dirs):   # more points than the fixed set: fill a Fibonacci sphere
        k = torch.arange(n - len(dirs), dtype=torch.float32) + 0.5
        phi = torch.acos(1 - 2 * k / (n - len(dirs)))
        th = math.pi * (1 + 5 ** 0.5) * k
        pts = torch.cat([pts, torch.stack([torch.cos(th) * torch.sin(phi), torch.sin(th) * torch.sin(phi), torch.cos(phi)], 1)])
    return pts * (2.0 / 3.0) * radius


class KPConvLayer(nn.Module):
    """Rigid kernel-point convolution: out_i = sum_k ( sum_{j in N(i)} max(0, 1 - |x_j - x_i - K_k| / extent) f_j ) W_k.
    Neighbour index -1 (padding of the radius search) is a shadow point far away with zero features."""
    _INFLUENCE_TO_RADIUS = 1.5

    def __init__(self, num_inputs, num_outputs, point_influence, n_kernel_points=15, fixed="center", KP_influence="linear",
                 aggregation_mode="sum", dimension=3, add_one=False, **kwargs):
        super().__init__()
        if KP_influence != "linear" or aggregation_mode != "sum" or dimension != 3:
            raise NotImplementedError("only the configuration the model uses: linear influence, sum aggregation, 3-D")
        self.kernel_radius = self._INFLUENCE_TO_RADIUS * point_influence
        self.point_influence = point_influence
        self.add_one = add_one
        self.num_inputs = num_inputs + self.add_one * 1
        self.num_outputs = num_outputs
        self.n_kernel_points = n_kernel_points
        self.K_points = nn.Parameter(default_kernel_points(self.kernel_radius, n_kernel_points), requires_grad=False)
        self.weight = nn.Parameter(torch.empty(n_kernel_points, self.num_inputs, num_outputs))
        nn.init.xavier_normal_(self.weight)

    def forward(self, query_points, support_points, neighbors, x):
        if x.is_cuda and x.shape[1] <= 16 and self.n_kernel_points <= 16:
            # neighbourhood sums in one kernel (no [n, nn, K, 3] intermediates), then ONE GEMM over (kernel point, channel)
            wsum = pointops.kpconv_weighted(query_points, support_points, neighbors, self.K_points, x, self.point_influence)
            return torch.matmul(wsum.reshape(wsum.shape[0], -1), self.weight.reshape(-1, self.weight.shape[-1]))
        return self.forward_torch(query_points, support_points, neighbors, x)

    def forward_torch(self, query_points, support_points, neighbors, x):
        """the published algorithm as torch operators (the specification of the kernel path above)"""
        n_sup = support_points.shape[0]
        nb = neighbors.long()
        nb = torch.where(nb < 0, torch.full_like(nb, n_sup), nb)                       # padding -> the shadow row
        sup = torch.cat([support_points, torch.full_like(support_points[:1], 1e6)], 0)
        rel = sup[nb] - query_points.unsqueeze(1)                                       # [n, nn, 3]
        d = torch.linalg.vector_norm(rel.unsqueeze(2) - self.K_points, dim=3)           # [n, nn, K]
        w = torch.clamp(1.0 - d / self.point_influence, min=0.0).transpose(1, 2)       # [n, K, nn]
        feats = torch.cat([x, torch.zeros_like(x[:1])], 0)[nb]                          # [n, nn, Cin]
        per_kernel = torch.matmul(w, feats).permute(1, 0, 2)                            # [K, n, Cin]
        return torch.matmul(per_kernel, self.weight).sum(0)                             # [n, Cout]


class FastBatchNorm1d(nn.Module):
    """torch_points3d.core.common_modules.FastBatchNorm1d: BatchNorm1d over [N, C] (or [B, N, C]) inputs; the inner module is
    called `batch_norm`, which is what the checkpoints' keys say."""

    def __init__(self, num_features, momentum=0.1, **kwargs):
        super().__init__()
        self.batch_norm = nn.BatchNorm1d(num_features, momentum=momentum, **kwargs)

    def forward(self, x):
        if x.dim() == 2:
            return self.batch_norm(x)
        if x.dim() == 3:
            return self.batch_norm(x.transpose(1, 2)).transpose(1, 2)
        raise ValueError(f"FastBatchNorm1d: expected 2-D or 3-D input, got {x.dim()}-D")


class KPConvSimpleBlock(nn.Module):
    """stratified_transformer.py:344-360"""

    def __init__(self, in_channels, out_channels, prev_grid_size, sigma=1.0, negative_slope=0.2, bn_momentum=0.02):
        super().__init__()
        self.kpconv = KPConvLayer(in_channels, out_channels, point_influence=prev_grid_size * sigma, add_one=False)
        self.bn = FastBatchNorm1d(out_channels, momentum=bn_momentum)
        self.activation = nn.LeakyReLU(negative_slope=negative_slope)

    def forward(self, feats, xyz, batch, neighbor_idx):
        return self.activation(self.bn(self.kpconv(xyz, xyz, neighbor_idx, feats)))


class KPConvResBlock(nn.Module):
    """stratified_transformer.py:363-396"""

    def __init__(self, in_channels, out_channels, prev_grid_size, sigma=1.0, negative_slope=0.2, bn_momentum=0.02):
        super().__init__()
        d_2 = out_channels // 4
        activation = nn.LeakyReLU(negative_slope=negative_slope)
        self.unary_1 = nn.Sequential(nn.Linear(in_channels, d_2, bias=False), FastBatchNorm1d(d_2, momentum=bn_momentum), activation)
        self.unary_2 = nn.Sequential(nn.Linear(d_2, out_channels, bias=False), FastBatchNorm1d(out_channels, momentum=bn_momentum),
                                     activation)
        self.kpconv = KPConvLayer(d_2, d_2, point_influence=prev_grid_size * sigma, add_one=False)
        self.bn = FastBatchNorm1d(out_channels, momentum=bn_momentum)
        self.activation = activation
        if in_channels != out_channels:
            self.shortcut_op = nn.Sequential(nn.Linear(in_channels, out_channels, bias=False),
                                             FastBatchNorm1d(out_channels, momentum=bn_momentum))
        else:
            self.shortcut_op = nn.Identity()

    def forward(self, feats, xyz, batch, neighbor_idx):
        out = self.unary_2(self.kpconv(xyz, xyz, neighbor_idx, self.unary_1(feats)))
        return out + self.shortcut_op(feats)


class Stratified(nn.Module):
    """stratified_transformer.py:399-505: KPConv stem -> BasicLayers with TransitionDown -> Upsample chain -> classifier and the
    offset regressor of this fork.  forward(feats, xyz, offset, batch, neighbor_idx) -> (logits [N, num_classes], shift [N, 3])."""

    def __init__(self, downsample_scale, depths, channels, num_heads, window_size, up_k, grid_sizes, quant_sizes, rel_query=True,
                 rel_key=False, rel_value=False, drop_path_rate=0.2, num_layers=4, concat_xyz=False, num_classes=13, ratio=0.25, k=16,
                 prev_grid_size=0.04, sigma=1.0, stem_transformer=False, activation="Relu"):
        super().__init__()
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, sum(depths))]
        c_in = 6 if concat_xyz else 3
        if stem_transformer:
            self.stem_layer = nn.ModuleList([KPConvSimpleBlock(c_in, channels[0], prev_grid_size, sigma=sigma)])
            self.layer_start = 0
        else:
            self.stem_layer = nn.ModuleList([KPConvSimpleBlock(c_in, channels[0], prev_grid_size, sigma=sigma),
                                             KPConvResBlock(channels[0], channels[0], prev_grid_size, sigma=sigma)])
            self.downsample = TransitionDown(channels[0], channels[1], ratio, k)
            self.layer_start = 1
        self.layers = nn.ModuleList([
            BasicLayer(downsample_scale, depths[i], channels[i], num_heads[i], window_size[i], grid_sizes[i], quant_sizes[i],
                       rel_query=rel_query, rel_key=rel_key, rel_value=rel_value, drop_path=dpr[sum(depths[:i]):sum(depths[:i + 1])],
                       downsample=TransitionDown if i < num_layers - 1 else None, ratio=ratio, k=k,
                       out_channels=channels[i + 1] if i < num_layers - 1 else None)
            for i in range(self.layer_start, num_layers)])
        self.upsamples = nn.ModuleList([Upsample(up_k, channels[i], channels[i - 1]) for i in range(num_layers - 1, 0, -1)])
        self.classifier = nn.Sequential(nn.Linear(channels[0], channels[0]), nn.BatchNorm1d(channels[0]), nn.ReLU(inplace=True),
                                        nn.Linear(channels[0], num_classes))
        act_reg = nn.Tanh() if activation == "Tanh" else nn.ReLU(inplace=True)
        self.regressor = nn.Sequential(nn.Linear(channels[0], channels[0]), nn.BatchNorm1d(channels[0]), act_reg,
                                       nn.Linear(channels[0], 3))
        self.init_weights()

    def forward(self, feats, xyz, offset, batch, neighbor_idx, geometry=None):
        """geometry: what `GeometryChain.take()` returns for this batch (sampling + pair lists of every layer, prefetched on a side
        stream during the previous step); None builds them inline, layer by layer."""
        # (Building them on a side stream under the stem of the SAME step was tried: 155.7 -> 154.0 ms on 8 x 80k points - the
        # sampling clusters take their SMs from the stem's bandwidth-bound kernels.  Across steps it pays: GeometryChain.)
        stages = iter(geometry["stages"]) if geometry is not None else None
        stack = []
        for layer in self.stem_layer:
            feats = layer(feats, xyz, batch, neighbor_idx)
        feats = feats.contiguous()
        if self.layer_start == 1:
            stack.append((feats, xyz, offset))
            fps = next(stages)[1] if stages is not None else None
            feats, xyz, offset = self.downsample(feats, xyz, offset, fps)
        for layer in self.layers:
            li, fps = next(stages) if stages is not None else (None, None)
            feats, xyz, offset, feats_down, xyz_down, offset_down = layer(feats, xyz, offset, layer_index=li, fps=fps)
            stack.append((feats, xyz, offset))
            feats, xyz, offset = feats_down, xyz_down, offset_down
        feats, xyz, offset = stack.pop()
        for upsample in self.upsamples:
            s_feats, s_xyz, s_offset = stack.pop()
            feats, xyz, offset = upsample(feats, xyz, s_xyz, offset, s_offset, support_feats=s_feats)
        return self.classifier(feats), self.regressor(feats)

    def init_weights(self):
        def _init(m):
            if isinstance(m, nn.Linear):
                trunc_normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, (nn.LayerNorm, nn.BatchNorm1d)):
                nn.init.constant_(m.bias, 0)
                nn.init.constant_(m.weight, 1.0)
        self.apply(_init)


class GeometryChain:
    """Everything `Stratified.forward` derives from the coordinates alone, for one batch, on a side stream: the per-point scene id and
    the radius neighbour lists of the stem (train.py:319-325), and per layer TransitionDown's furthest-point picks, the stratified key
    candidates (their prefix) and the pair lists of both block parities.  None of it depends on features, so the chain of batch t+1
    can run under the compute of batch t (the sampling loops are long sequential kernels on a few SMs).  Measured on the full
    S3DIS network, 8 x 80k points: 143.2 -> 141.6 ms per step - the sampling clusters mostly wait for SMs - so `bench.py` does not use it:

        chain = GeometryChain(model); chain.submit(xyz0, offset0, radius)
        for batch in loader:
            geo = chain.take()                                  # current batch: main stream waits for its events
            out = model(feats, xyz, offset, geo["batch"], geo["neighbor_idx"], geometry=geo)
            chain.submit(next_xyz, next_offset, radius)         # between forward and backward works best
            loss.backward(); chain.complete()                   # host: sizes of the pair lists, then their fill kernels
    """

    def __init__(self, model: Stratified, max_neighbors: int = 34):
        self.model, self.max_neighbors = model, max_neighbors
        self.side = None
        self.pending = self.done = None

    def submit(self, xyz, offset, radius=None):
        model = self.model
        main = torch.cuda.current_stream()
        if self.side is None:
            self.side = torch.cuda.Stream(device=xyz.device)
        side = self.side
        side.wait_stream(main)
        xyz.record_stream(side)
        off_host = [int(v) for v in offset.tolist()]
        stages = []
        with torch.cuda.stream(side):
            off_t = offset.to(torch.int32).contiguous()
            batch = prestep.batch_from_offset(off_t, xyz.shape[0])
            nbr = None
            if radius is not None:
                nbr = prestep.ball_query(radius, self.max_neighbors, xyz, xyz, mode="partial_dense", batch_x=batch, batch_y=batch)[0]
            downs = ([model.downsample] if model.layer_start == 1 else []) + [None] * len(model.layers)
            layers = ([None] if model.layer_start == 1 else []) + list(model.layers)
            for pre, layer in zip(downs, layers):
                td = pre if layer is None else layer.downsample
                fps = n_off_t = pend = n_host = None
                if isinstance(td, TransitionDown):
                    n_host = transition_down_offsets(off_host, td.ratio)
                    n_off_t = torch.tensor(n_host, dtype=torch.int32, device=xyz.device)
                if layer is not None:
                    ds = layer.downsample_scale
                    k_host, c = [], 0
                    for i, o in enumerate(off_host):
                        c += (o - (off_host[i - 1] if i else 0)) // ds + 1
                        k_host.append(c)
                    k_off_t = torch.tensor(k_host, dtype=torch.int32, device=xyz.device)
                    prefix_ok = n_host is not None and all(
                        (k_host[i] - (k_host[i - 1] if i else 0)) <= (n_host[i] - (n_host[i - 1] if i else 0)) for i in range(len(k_host)))
                    if prefix_ok:
                        fps = self._fps(xyz, off_t, off_host, n_off_t, n_host)
                        ds_idx = st_index.fps_prefix(fps, n_off_t, k_off_t, total=k_host[-1])
                    else:
                        ds_idx = self._fps(xyz, off_t, off_host, k_off_t, k_host)
                    pend = st_index.PendingLayerIndex(xyz, off_t, layer.window_size, layer.quant_size, ds, off_host, fused=layer.fused,
                                                      downsample_idx=ds_idx)
                if n_host is not None and fps is None:
                    fps = self._fps(xyz, off_t, off_host, n_off_t, n_host)
                stages.append((fps, n_off_t, pend))
                if fps is not None:
                    xyz = xyz[fps.long(), :].contiguous()
                    off_t, off_host = n_off_t, n_host
            event = side.record_event()
        self.pending = (batch, nbr, stages, event)

    @staticmethod
    def _fps(xyz, off_t, off_host, new_off_t, new_host):
        """pointops.furthestsampling without its device-to-host read: the sizes are known on the host here, and a synchronisation
        inside submit() would keep the host from enqueueing the step's backward pass while the sampling loop runs"""
        from . import pointops2_cuda as ext
        n_max = max(o - (off_host[i - 1] if i else 0) for i, o in enumerate(off_host))
        idx = torch.zeros(int(new_host[-1]), dtype=torch.int32, device=xyz.device)
        ext.furthestsampling_cuda(len(off_host), n_max, xyz.contiguous(), off_t, new_off_t, None, idx)
        return idx

    def complete(self):
        if self.pending is None:
            return
        batch, nbr, stages, event = self.pending
        with torch.cuda.stream(self.side):
            done = [(None if pend is None else pend.finish(), None if fps is None else (fps, n_off)) for fps, n_off, pend in stages]
            event = self.side.record_event()
        self.done, self.pending = (batch, nbr, done, event), None

    def take(self):
        if self.done is None:
            self.complete()
        (batch, nbr, done, event), self.done = self.done, None
        main = torch.cuda.current_stream()
        main.wait_event(event)
        for t in (batch, nbr):
            if t is not None:
                t.record_stream(main)
        for li, fps in done:
            if li is not None:
                st_index.record_stream(li, main)
            if fps is not None:
                fps[0].record_stream(main)
                fps[1].record_stream(main)
        return {"batch": batch, "neighbor_idx": nbr, "stages": done}
