"""Drop-in proof (-m gpu): the REFERENCE's own `WindowAttention` class (model/stratified_transformer.py:114-217, text
extracted at build time into the git-ignored oracle/_ref/ref_window_attention.py with only its imports redirected to this
package's `pointops` and the scatter_softmax shim) runs unmodified over libstb200, and agrees with this package's mirror
module — same state dict, same inputs, forward and all parameter / input gradients — on both of the mirror's paths
(per-op kernels and fused plan)."""
import importlib.util
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GEN = os.path.join(ROOT, "oracle", "_ref", "ref_window_attention.py")


def _load_reference_class():
    if not os.path.exists(GEN):
        pytest.skip("oracle/_ref/ref_window_attention.py not generated (needs /root/reference at build time)")
    spec = importlib.util.spec_from_file_location("ref_window_attention", GEN)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.WindowAttention


@pytest.mark.parametrize("parity", [0, 1])
def test_reference_window_attention_runs_on_libstb200_and_matches_mirror(parity):
    from oracle import fps_oracle, index_oracle as io
    from stratified_transformer_b200 import index
    from stratified_transformer_b200.synthetic import make_batch
    from stratified_transformer_b200.window_attention import WindowAttention
    RefWindowAttention = _load_reference_class()
    xyz, _, offset = make_batch(2, 3000, seed0=41, n_raw=80000)
    ds = fps_oracle.furthestsampling(xyz, offset, io.fps_new_offset(offset, 8))
    xd, od = torch.from_numpy(xyz).cuda(), torch.from_numpy(offset).cuda()
    window, quant, C, h = 0.32, 0.02, 48, 3
    pi = index.build_stratified_index(xd, od, window, quant, torch.from_numpy(ds).cuda(), parity, want_index_0=True, fused=True)
    torch.manual_seed(0)
    ref = RefWindowAttention(C, window, h, quant, rel_query=True, rel_key=True, rel_value=True).cuda()
    for t in (ref.relative_pos_query_table, ref.relative_pos_key_table, ref.relative_pos_value_table):
        torch.nn.init.uniform_(t, -0.3, 0.3)
    mine = WindowAttention(C, window, h, quant, rel_query=True, rel_key=True, rel_value=True).cuda()
    mine.load_state_dict(ref.state_dict())        # same parameter names and shapes: reference checkpoints load
    feats = torch.randn(xyz.shape[0], C, device="cuda")
    n_max = torch.tensor(pi.n_max, device="cuda")

    def run(module, *args):
        module.zero_grad(set_to_none=True)
        f = feats.clone().requires_grad_(True)
        y = module(f, xd, *args)
        y.square().sum().backward()
        return dict(y=y.detach(), gf=f.grad, **{n: p.grad.clone() for n, p in module.named_parameters()})

    # the reference's call form: int64 index tensors, n_max as a 0-dim CUDA tensor (stratified_transformer.py:312-319)
    want = run(ref, pi.index_0.long(), pi.index_1.long(), pi.index_0_offsets.long(), n_max)
    got_perop = run(mine, pi.index_0.long(), pi.index_1.long(), pi.index_0_offsets.long(), n_max)
    got_plan = run(mine, pi)
    for name, w in want.items():
        scale = max(1.0, float(w.abs().max()))
        for tag, got in (("per-op", got_perop), ("plan", got_plan)):
            err = float((got[name] - w).abs().max())
            assert err <= 5e-4 * scale, f"{name} ({tag}): max err {err:.3e} at scale {scale:.2f}"
