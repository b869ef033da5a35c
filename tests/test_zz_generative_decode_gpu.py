"""GPU parity of the generative decode path `GShell_Tets.marching_from_auggrid` (reference gshell_tets.py:446-629) against goldens
of the unmodified reference.  Kept in its own file that sorts last: the path was added after this round's GPU budget was spent,
so its host logic is verified on CPU (tests/test_oracle_auggrid.py) but this device run is its first; a surprise here must not
stop the `-x` run before the established suites."""
import glob
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


AUGGRID = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "auggrid_*.npz")))


@pytest.mark.parametrize("path", AUGGRID, ids=[os.path.basename(p) for p in AUGGRID])
def test_marching_from_auggrid_matches_reference_golden(path):
    """Generative decode path (reference gshell_tets.py:446-629) on the device against goldens of the unmodified reference
    (generator tests/golden/make_golden_auggrid.py): topology bit-exact, positions / mSDF to fp32 rounding, tangents in bulk."""
    from gshell_b200.geometry.gshell_tets import GShell_Tets
    z = np.load(path)
    g = {k: torch.from_numpy(z[k]) for k in z.files}
    d = _dev()
    va, fa, a, b, tng, v, gidx, m_aug, m = GShell_Tets().marching_from_auggrid(
        g["pos"].to(d), g["sdf"].to(d), g["tets"].to(d), g["sorted_edges"].to(d), g["coeff"].to(d), g["disc"].to(d),
        g["msdf_sign"].to(d), g["occ"].to(d))
    assert a is None and b is None
    assert torch.equal(fa.cpu().long(), g["faces_aug"].long())
    assert torch.equal(gidx.cpu().long(), g["valid_tet_gidx"].long())
    for got, key in ((va, "verts_aug"), (v, "verts"), (m_aug, "msdf_aug"), (m, "msdf")):
        assert got.shape == g[key].shape, key
        if got.numel():
            assert float((got.cpu() - g[key]).abs().max()) <= 1e-5, key
    assert tng.shape == g["v_tng_aug"].shape
    if tng.numel():
        want = g["v_tng_aug"]
        ok = torch.isfinite(want).all(-1) & torch.isfinite(tng.cpu()).all(-1)
        err = (tng.cpu()[ok] - want[ok]).abs().max(-1)[0]
        # atomics sum the per-face tangents in launch order: one ill-conditioned (nearly cancelling) row may flip per run; on a
        # 48-row fixture a single row is already 2 %
        assert float(err.median()) < 1e-4 and float((err > 1e-3).float().mean()) < max(0.03, 1.5 / err.numel())
