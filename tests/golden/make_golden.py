"""Generate tests/golden/*.npz by running the REFERENCE's own Python on CPU (build container only).

Nothing from /root/reference is copied into the repository: this script reads
/root/reference/model/stratified_transformer.py at run time, extracts the source text of
`get_indice_pairs` and `grid_sample` with `ast`, strips the hard-coded `.cuda()` calls and
executes them with CPU torch.  The only thing it has to supply is `voxel_grid`, which the
reference imports from torch_geometric (third party, absent here): oracle.index_oracle.voxel_grid
is used for that, so these fixtures pin grid_sample / get_indice_pairs / sort+CSR, not voxel_grid.

The rel-pos index fixtures evaluate the three torch statements of
model/stratified_transformer.py:186-188 (and model/swin3d_transformer.py:151-154) with CPU torch.

Because the reference sorts with unstable sorts, `index_1` is stored
canonicalised: keys sorted inside each query segment.

Run:  python tests/golden/make_golden.py      (needs /root/reference; not run on the GPU box)
"""
import ast
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import fps_oracle, index_oracle as io  # noqa: E402
from stratified_transformer_b200.synthetic import make_batch  # noqa: E402

REF_MODEL = "/root/reference/model/stratified_transformer.py"
OUT = os.path.dirname(os.path.abspath(__file__))


def load_reference_functions():
    src = open(REF_MODEL).read()
    tree = ast.parse(src)
    ns = {"torch": torch}

    def voxel_grid(pos, batch, size, start=None):
        st = None if start is None else start.numpy()
        return torch.from_numpy(io.voxel_grid(pos.numpy(), batch.numpy(), size.numpy(), st))

    ns["voxel_grid"] = voxel_grid
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("get_indice_pairs", "grid_sample"):
            text = ast.get_source_segment(src, node).replace(".cuda()", "")
            exec(compile(text, REF_MODEL, "exec"), ns)
    return ns["grid_sample"], ns["get_indice_pairs"]


def reference_layer_index(grid_sample, get_indice_pairs, xyz, offset, window_size, ds_idx, parity):
    """BasicLayer.forward:267-317 restated call-for-call around the two extracted functions."""
    xyz_t = torch.from_numpy(xyz)
    batch = torch.from_numpy(io.batch_from_offset(offset))
    w = torch.tensor([window_size] * 3).type_as(xyz_t)
    nw = 2 * torch.tensor([window_size] * 3).type_as(xyz_t)
    if parity % 2 == 0:
        _, p2v, cnt = grid_sample(xyz_t, batch, w, start=None)
        _, np2v, ncnt = grid_sample(xyz_t, batch, nw, start=None)
    else:
        _, p2v, cnt = grid_sample(xyz_t + 1 / 2 * w, batch, w, start=xyz_t.min(0)[0])
        _, np2v, ncnt = grid_sample(xyz_t + 1 / 2 * nw, batch, nw, start=xyz_t.min(0)[0])
    i0, i1 = get_indice_pairs(p2v, cnt, np2v, ncnt, torch.from_numpy(ds_idx), batch, xyz_t, w, parity)
    i0, perm = torch.sort(i0)
    i1 = i1[perm]
    counts = i0.bincount()
    n_max = int(counts.max())
    offsets = torch.cat([torch.zeros(1, dtype=torch.long), counts.cumsum(-1)], 0)
    return i0.numpy(), i1.numpy(), offsets.numpy(), n_max


def main():
    grid_sample, get_indice_pairs = load_reference_functions()
    cases = [
        # name, scenes, pts/scene, voxel, window, quant, downsample_scale, lattice
        ("s3dis_small", 2, 1500, 0.04, 0.16, 0.01, 8, False),
        ("s3dis_lattice", 1, 4000, 0.04, 0.16, 0.01, 8, True),
        ("scannet_small", 1, 2000, 0.02, 0.2, 0.01, 4, False),
    ]
    for name, b, n, voxel, w, quant, ds, lattice in cases:
        xyz, _, offset = make_batch(b, n, voxel, seed0=7, n_raw=200_000, lattice=lattice)
        new_offset = io.fps_new_offset(offset, ds)
        ds_idx = fps_oracle.furthestsampling(xyz, offset, new_offset)
        out = dict(xyz=xyz, offset=offset, new_offset=new_offset, downsample_idx=ds_idx,
                   window_size=np.float64(w), quant_size=np.float64(quant), downsample_scale=np.int64(ds))
        for parity in (0, 1):
            i0, i1, offsets, n_max = reference_layer_index(grid_sample, get_indice_pairs, xyz, offset, w, ds_idx, parity)
            i1c = io.canonicalize(offsets, i1)
            # rel-pos index, the reference's three torch statements (stratified_transformer.py:186-188)
            xyz_t = torch.from_numpy(xyz)
            i0_t, i1_t = torch.from_numpy(i0), torch.from_numpy(i1c)
            rel = xyz_t[i0_t] - xyz_t[i1_t]
            rel = torch.round(rel * 100000) / 100000
            rel_idx = ((rel + 2 * w - 0.0001) // quant).int().numpy()
            out[f"p{parity}_offsets"] = offsets.astype(np.int32)
            out[f"p{parity}_index_1"] = i1c.astype(np.int32)
            out[f"p{parity}_n_max"] = np.int64(n_max)
            out[f"p{parity}_rel_idx"] = rel_idx.astype(np.int8 if rel_idx.max() < 127 and rel_idx.min() > -128 else np.int32)
            print(name, "parity", parity, "N", xyz.shape[0], "M", i1.shape[0], "n_max", n_max,
                  "rel_idx range", rel_idx.min(), rel_idx.max())
        np.savez_compressed(os.path.join(OUT, f"index_{name}.npz"), **out)

    # Swin rel-pos index (swin3d_transformer.py:151-154, 129-130) on one case
    xyz, _, offset = make_batch(1, 1500, 0.04, seed0=11, n_raw=200_000)
    xyz_t = torch.from_numpy(xyz)
    rng = np.random.default_rng(3)
    i0 = np.sort(rng.integers(0, xyz.shape[0], 20000))
    i1 = rng.integers(0, xyz.shape[0], 20000)
    res = dict(xyz=xyz, index_0=i0.astype(np.int32), index_1=i1.astype(np.int32))
    for shift, tag in ((0.0, "noshift"), (0.08, "shift")):
        window_size, quant_size = 0.16, 0.01
        qgl = int(window_size / quant_size)
        xq = (xyz_t - xyz_t.min(0)[0] + shift) % window_size
        xq = xq // quant_size
        rp = xq[torch.from_numpy(i0)] - xq[torch.from_numpy(i1)]
        res[f"rel_idx_{tag}"] = (rp + qgl - 1).int().numpy().astype(np.int8)
    np.savez_compressed(os.path.join(OUT, "relidx_swin.npz"), **res)
    print("wrote fixtures to", OUT)


if __name__ == "__main__":
    main()
