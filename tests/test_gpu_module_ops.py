"""stb200_qkv_split / stb200_qkv_merge (the elementwise passes of WindowAttention.forward around the pair ops,
/root/reference/model/stratified_transformer.py:172-175) against the torch statements they replace."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("N,C,h", [(1, 48, 3), (1000, 48, 3), (4097, 96, 6), (777, 192, 12), (300, 384, 24), (50, 8, 1)])
def test_split_qkv_forward_backward_match_torch(dtype, N, C, h):
    from stratified_transformer_b200 import pointops
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(N + C)
    qkv = torch.randn(N, 3 * C, device=dev, generator=g).to(dtype).requires_grad_(True)
    bias = torch.randn(3 * C, device=dev, generator=g).requires_grad_(True)
    q, k, v = pointops.split_qkv(qkv, bias, h)
    # the reference's statements (bias inside the Linear, fp32 casts at the call sites)
    ref = (qkv.detach().float() + bias.detach()).reshape(N, 3, h, C // h).permute(1, 0, 2, 3).contiguous()
    for got, want in zip((q, k, v), ref):
        assert got.dtype == torch.float32 and got.is_contiguous() and got.shape == (N, h, C // h)
        assert torch.equal(got, want)      # one fp32 add per element: exact
    gq, gk, gv = (torch.randn(N, h, C // h, device=dev, generator=g) for _ in range(3))
    torch.autograd.backward((q, k, v), (gq, gk, gv))
    want_g = torch.stack([gq, gk, gv], 1).reshape(N, 3 * C)
    assert qkv.grad.dtype == dtype
    assert torch.equal(qkv.grad, want_g.to(dtype))      # round-to-nearest cast: exact
    want_b = want_g.double().sum(0)
    assert torch.allclose(bias.grad.double(), want_b, rtol=1e-5, atol=1e-4 * max(1.0, N ** 0.5))


def test_split_qkv_without_bias_and_bad_shapes():
    from stratified_transformer_b200 import pointops
    dev = torch.device("cuda")
    qkv = torch.randn(33, 144, device=dev)
    q, k, v = pointops.split_qkv(qkv, None, 3)
    assert torch.equal(torch.cat([q.reshape(33, 48), k.reshape(33, 48), v.reshape(33, 48)], 1), qkv)
    with pytest.raises(ValueError):
        pointops.split_qkv(torch.randn(4, 3 * 12, device=dev), None, 3)     # C not a multiple of 8
    with pytest.raises(ValueError):
        pointops.split_qkv(torch.randn(4, 144, device=dev).double(), None, 3)


def test_bias_gradient_is_deterministic():
    from stratified_transformer_b200 import pointops
    dev = torch.device("cuda")
    outs = []
    for _ in range(3):
        torch.manual_seed(1)
        qkv = torch.randn(50000, 144, device=dev).bfloat16().requires_grad_(True)
        bias = torch.zeros(144, device=dev, requires_grad=True)
        q, k, v = pointops.split_qkv(qkv, bias, 3)
        (q.sum() + (k * 2).sum() + (v * q.detach()).sum()).backward()
        outs.append(bias.grad.clone())
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


@pytest.mark.parametrize("shape", [(1, 48), (1000, 48), (4097, 96), (777, 192), (300, 384), (50, 16, 48), (33, 7), (64, 100)])
@pytest.mark.parametrize("affine", [True, False])
def test_layer_norm_short_rows_matches_torch(shape, affine):
    from stratified_transformer_b200 import pointops
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(sum(shape))
    C = shape[-1]
    x = (torch.randn(*shape, device=dev, generator=g) * 3 + 1).requires_grad_(True)
    w = (torch.rand(C, device=dev, generator=g) + 0.5).requires_grad_(True) if affine else None
    b = torch.randn(C, device=dev, generator=g).requires_grad_(True) if affine else None
    gy = torch.randn(*shape, device=dev, generator=g)
    y = pointops.layer_norm(x, w, b, 1e-5)
    y.backward(gy)
    got = [y.detach(), x.grad.clone()] + ([w.grad.clone(), b.grad.clone()] if affine else [])
    x64 = x.detach().double().requires_grad_(True)
    w64 = w.detach().double().requires_grad_(True) if affine else None
    b64 = b.detach().double().requires_grad_(True) if affine else None
    y64 = torch.nn.functional.layer_norm(x64, (C,), w64, b64, 1e-5)
    y64.backward(gy.double())
    want = [y64.detach(), x64.grad] + ([w64.grad, b64.grad] if affine else [])
    for a, r, name in zip(got, want, ("y", "gx", "gw", "gb")):
        scale = max(1.0, float(r.abs().max()))
        assert float((a.double() - r).abs().max()) <= 2e-5 * scale, name


def test_layer_norm_module_is_a_drop_in_under_autocast():
    from stratified_transformer_b200.layers import LayerNorm
    ref = torch.nn.LayerNorm(96).cuda()
    mine = LayerNorm(96).cuda()
    mine.load_state_dict(ref.state_dict())
    x = torch.randn(5000, 96, device="cuda")
    with torch.autocast("cuda", dtype=torch.bfloat16):
        a, b = mine(x.bfloat16()), ref(x.bfloat16())
    assert a.dtype == b.dtype == torch.float32
    assert torch.allclose(a, b, atol=2e-5, rtol=1e-5)
    assert isinstance(mine, torch.nn.LayerNorm)
