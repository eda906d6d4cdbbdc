"""Host-side logic of the callers around the hot path (no GPU): offsets arithmetic, the FPS prefix property on plain tensors, the
restated stem blocks as torch operators, state-dict compatibility of the mirrors."""
import numpy as np
import torch

from oracle import fps_oracle


def test_transition_down_offsets_follow_the_reference_arithmetic():
    """model/stratified_transformer.py:98-102: int() only on the first scene, fractions accumulate in a Python float, the final
    IntTensor truncates."""
    from stratified_transformer_b200.layers import transition_down_offsets
    assert transition_down_offsets([1001, 2003, 3006, 4010], 0.25) == [251, 502, 754, 1006]
    assert transition_down_offsets(torch.tensor([80000]), 0.25) == [20001]
    assert transition_down_offsets([7], 0.25) == [2]


def test_fps_prefix_on_the_c_oracle():
    """the shorter FPS run is a prefix of the longer one, per scene (index.fps_prefix gathers it)"""
    from stratified_transformer_b200 import index
    rng = np.random.default_rng(0)
    sizes = [400, 650, 90]
    xyz = rng.random((sum(sizes), 3)).astype(np.float32)
    offset = np.cumsum(sizes).astype(np.int32)
    long_off = np.cumsum([n // 4 + 1 for n in sizes]).astype(np.int32)
    short_off = np.cumsum([n // 8 + 1 for n in sizes]).astype(np.int32)
    long_idx = fps_oracle.furthestsampling(xyz, offset, long_off)
    want = fps_oracle.furthestsampling(xyz, offset, short_off)
    got = index.fps_prefix(torch.from_numpy(long_idx), torch.from_numpy(long_off), torch.from_numpy(short_off))
    assert np.array_equal(got.numpy(), want)
    got2 = index.fps_prefix(torch.from_numpy(long_idx), torch.from_numpy(long_off), torch.from_numpy(short_off), total=int(short_off[-1]))
    assert np.array_equal(got2.numpy(), want)


def test_kpconv_restatement_as_torch_operators():
    """KPConvLayer.forward_torch: shapes, padding neighbours contribute nothing, a point on a kernel point gets full weight"""
    from stratified_transformer_b200.model import KPConvLayer, default_kernel_points
    kp = default_kernel_points(0.06, 15)
    assert kp.shape == (15, 3) and torch.allclose(kp[0], torch.zeros(3)) and torch.allclose(kp[1:].norm(dim=1), torch.full((14,), 0.04))
    torch.manual_seed(0)
    layer = KPConvLayer(4, 5, point_influence=0.04)
    xyz = torch.zeros(3, 3)
    xyz[1] = layer.K_points[3]                                   # neighbour sitting exactly on kernel point 3 of query 0
    feats = torch.randn(3, 4)
    nbr = torch.tensor([[1, -1], [-1, -1], [-1, -1]])
    out = layer.forward_torch(xyz, xyz, nbr, feats)
    assert out.shape == (3, 5)
    assert torch.allclose(out[1], torch.zeros(5)) and torch.allclose(out[2], torch.zeros(5))
    d = (xyz[1] - layer.K_points).norm(dim=1)
    w = torch.clamp(1 - d / 0.04, min=0)
    assert abs(float(w[3]) - 1.0) < 1e-6
    assert torch.allclose(out[0], torch.einsum("k,c,kco->o", w, feats[1], layer.weight), atol=1e-6)


def test_mirror_modules_keep_the_reference_parameter_names():
    from stratified_transformer_b200.layers import BasicLayer, LayerNorm, TransitionDown
    layer = BasicLayer(8, 2, 48, 3, 0.16, 0.04, 0.01, rel_query=True, rel_key=True, rel_value=True, downsample=TransitionDown,
                       out_channels=96)
    keys = set(layer.state_dict())
    for k in ("blocks.0.norm1.weight", "blocks.0.attn.qkv.weight", "blocks.0.attn.relative_pos_query_table", "blocks.1.mlp.fc2.bias",
              "downsample.norm.weight", "downsample.linear.weight"):
        assert k in keys, k
    assert layer.blocks[0].attn.relative_pos_query_table.shape == (64, 3, 16, 3)
    ln = LayerNorm(48)
    x = torch.randn(10, 48)
    assert torch.allclose(ln(x), torch.nn.functional.layer_norm(x, (48,), ln.weight, ln.bias, ln.eps))   # CPU tensors: torch's path
