"""Tangent frame of the extraction (reference gshell_tets.py:40-78, 210-239, 318-319, 337-380): the host
composition in gshell_b200/geometry/tangents.py against goldens of the UNMODIFIED reference (tests/golden/mt_*.npz:
`v_tng_aug` and the gradients `gtng_*` of <v_tng_aug, Wt> with respect to pos / sdf / msdf).

tangents.py is torch ops around one CUDA kernel (vertex normals); here that kernel is replaced by the oracle's restatement of
the same sum so that the composition -- atlas arithmetic, per-face tangents, Gram-Schmidt, boundary weights and in particular
the GRADIENT STRUCTURE (the boundary weights are built from `msdf_vert`, whose gradient reaches the SDF through the
interpolation weights, reference :287-288, :345-365) -- is checked on the CPU.  The extraction's own outputs that feed it are
taken from the oracle (bit-exact with the kernels, tests/test_mt_gpu.py)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import mt_oracle as mo

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "mt_*.npz")))


def _grid_edges(tets):
    """All grid edges, endpoints sorted, rows in lexicographic order (what geometry/tet_tables.py holds as `edge_v`)."""
    e = tets[:, [0, 1, 0, 2, 0, 3, 1, 2, 1, 3, 2, 3]].reshape(-1, 2)
    return torch.unique(torch.sort(e, dim=1)[0], dim=0).int()


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[3:-4] for p in GOLDEN])
def test_tangent_composition_matches_reference_values_and_gradients(path, monkeypatch):
    z = np.load(path)
    g = {k: torch.from_numpy(z[k]) if z[k].shape != () else z[k] for k in z.files}
    if "v_tng_aug" not in g or g["v_tng_aug"].shape[0] == 0:
        pytest.skip("empty surface")
    import gshell_b200.geometry.tangents as tg
    monkeypatch.setattr(tg, "vertex_normals", lambda v, f: mo.smooth_normals(v, f.long()))
    pos, sdf, msdf = (g[k].clone().requires_grad_() for k in ("pos", "sdf", "msdf"))
    tets = g["tets"]
    with torch.no_grad():
        valid, case, vmap, edge = mo.crossing_edges(sdf, tets, "packed", None)
    verts, _, m_sg = mo.lerp_on_sdf(pos, sdf, msdf, edge)          # m_sg = extra['msdf_watertight'] (stop-gradient form)
    with torch.no_grad():
        faces, one, two = mo.watertight_faces(case, vmap)
        tri, quad = mo.polygon_loops(case, vmap, one, two)
    slot_a = torch.cat([tri[:, :, 0].reshape(-1), quad[:, :, 0].reshape(-1)]).int()
    v_tng, v_aug = tg.tangent_frame_aug(verts, faces, m_sg, slot_a, tets.shape[0], tri.shape[0],
                                        sdf=sdf, msdf=msdf, edge_v=_grid_edges(tets))
    # values: same IEEE op order as the reference on the CPU
    torch.testing.assert_close(v_aug.detach(), g["v_tng_aug"], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(v_tng.detach(), g["v_tng_watertight"], rtol=1e-4, atol=1e-5)
    grads = torch.autograd.grad((v_aug * g["wt"]).sum(), [pos, sdf, msdf], allow_unused=True)
    for name, got in zip(("pos", "sdf", "msdf"), grads):
        want = g[f"gtng_{name}"]
        got = torch.zeros_like(want) if got is None else got
        l2 = float((got - want).norm() / want.norm().clamp(min=1e-20))
        assert l2 < 1e-5, (name, l2)
    # without the grid values the boundary weights only see the mSDF: the SDF gradient is then NOT the reference's
    v_tng2, v_aug2 = tg.tangent_frame_aug(verts, faces, m_sg, slot_a, tets.shape[0], tri.shape[0])
    assert torch.equal(v_aug2.detach(), v_aug.detach())            # the value does not depend on which form is used
