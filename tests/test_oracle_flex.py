"""The FlexiCubes oracle restatement vs. golden outputs of the unmodified reference (tests/golden/flex_*.npz)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import flexicubes_oracle as fo

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "flex_*.npz")))


def load(path):
    z = np.load(path)
    return {k: torch.from_numpy(z[k]) if z[k].shape != () else z[k] for k in z.files}


def test_voxel_grid_matches_reference_layout():
    g = load([p for p in GOLDEN if p.endswith("flex_r4.npz")][0])
    verts, cubes = fo.voxel_grid(4)
    assert torch.equal(cubes, g["cubes"])


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[5:-4] for p in GOLDEN])
def test_oracle_matches_reference(path):
    g = load(path)
    res = int(g["res"])
    leaves = [g[k].clone().requires_grad_() for k in ("x", "s", "nu")]
    w = g["w"].clone().requires_grad_() if "w" in g else None
    args = (w[:, :12], w[:, 12:20], w[:, 20]) if w is not None else (None, None, None)
    vo, fa, L, ex = fo.gflexicubes(leaves[0], leaves[1], leaves[2], g["cubes"], res, *args)
    if "n_out" in g:
        assert vo.shape[0] == 0 and fa.shape[0] == 0
        return
    assert torch.equal(fa, g["faces_open"])
    assert torch.equal(ex["faces_watertight"], g["faces_watertight"])
    assert ex["n_verts_watertight"] == int(g["n_verts_watertight"])
    for got, key in ((vo, "vertices_open"), (ex["vertices_watertight"], "vertices_watertight"), (ex["msdf"], "msdf"),
                     (ex["msdf_watertight"], "msdf_watertight"), (ex["msdf_boundary"], "msdf_boundary"), (L, "L_dev")):
        torch.testing.assert_close(got, g[key], rtol=1e-5, atol=1e-6, msg=key)
    probe = (vo * g["wv"]).sum() + (ex["msdf"] * g["wm"]).sum() + (L * g["wl"]).sum() + (ex["vertices_watertight"] * g["ww"]).sum()
    grads = torch.autograd.grad(probe, leaves + ([w] if w is not None else []), allow_unused=True)
    for nm, got in zip(("x", "s", "nu", "w"), grads):
        want = g["g_" + nm]
        got = torch.zeros_like(want) if got is None else got
        scale = want.abs().max().clamp(min=1.0)
        assert (got - want).abs().max() <= 1e-4 * scale, nm
