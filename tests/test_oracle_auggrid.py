"""marching_from_auggrid (generative decode path, reference gshell_tets.py:446-629): the oracle restatement and the product's
table-driven host implementation against goldens of the UNMODIFIED reference (tests/golden/auggrid_*.npz, generator beside
them).  Topology must be bit-exact, floats within 1e-6 (same fp32 ops in the same order)."""
import glob
import os

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDENS = sorted(glob.glob(os.path.join(HERE, "golden", "auggrid_*.npz")))


def _load(path):
    z = np.load(path)
    return {k: torch.from_numpy(z[k]) for k in z.files}


def _determined_rows(g):
    """Rows of v_tng_aug that the inputs determine (oracle/mt_oracle.py::determined_tangent_rows): the generated coefficients are
    clamped to [0, 1], so some vertices coincide with grid vertices and their face normals cancel -- those rows are rounding
    residue that follows the summation order."""
    from oracle import mt_oracle as mo
    valid, case, vmap, edge = mo.crossing_edges(g["sdf"].float().reshape(-1), g["tets"], "packed", None)
    faces, one, two = mo.watertight_faces(case, vmap)
    tri, quad = mo.polygon_loops(case, vmap, one, two)
    return mo.determined_tangent_rows(g["verts"], faces, tri, quad, g["tets"].shape[0])


def _compare(out, g, with_tangents, exact_order=True):
    va, fa, a, b, tng, v, gidx, m_aug, m = out
    assert a is None and b is None
    assert torch.equal(fa.long(), g["faces_aug"].long())
    assert torch.equal(gidx.long(), g["valid_tet_gidx"].long())
    for got, key in ((va, "verts_aug"), (v, "verts"), (m_aug, "msdf_aug"), (m, "msdf")):
        assert got.shape == g[key].shape, key
        if got.numel():
            assert float((got - g[key]).abs().max()) <= 1e-6, key
    if with_tangents:
        assert tng.shape == g["v_tng_aug"].shape
        if tng.numel():
            ok = torch.isfinite(g["v_tng_aug"]).all(-1)
            if not exact_order:                  # kernels: another summation order than the reference's scatter_add_ passes
                det = _determined_rows(g)
                assert float(det.float().mean()) > 0.9
                ok = ok & det
            assert float((tng[ok] - g["v_tng_aug"][ok]).abs().max()) <= 1e-4


@pytest.mark.parametrize("path", GOLDENS, ids=[os.path.basename(p) for p in GOLDENS])
def test_oracle_matches_reference(path):
    from oracle.mt_oracle import gshell_marching_from_auggrid
    g = _load(path)
    out = gshell_marching_from_auggrid(g["pos"], g["sdf"], g["tets"], g["sorted_edges"], g["coeff"], g["disc"], g["msdf_sign"],
                                       g["occ"])
    _compare(out, g, with_tangents=True)


@pytest.mark.parametrize("order_seed", [0, 3])
@pytest.mark.parametrize("path", GOLDENS, ids=[os.path.basename(p) for p in GOLDENS])
def test_product_decode_kernels_match_reference(path, order_seed, host_kernels_lib, monkeypatch):
    """The product's decode (csrc/auggrid.cu + the tangent kernels behind GShell_Tets._marching_from_auggrid) on the CPU: the
    UNMODIFIED kernel source compiled as host code behind the same C ABI (tests/native/host_kernels.py), the product's Python
    layer and ctypes signatures unchanged, threads in ascending (0) and shuffled (3) order.  The public method refuses non-CUDA
    tensors, hence the internal entry point; the GPU run of the same code: tests/test_zz3_generative_decode_gpu.py."""
    stand_in, host_kernels = host_kernels_lib
    import gshell_b200.geometry.gshell_tets as gt
    import gshell_b200.geometry.tangents as tg
    import gshell_b200.render.mesh as mesh
    for mod in (gt, tg, mesh):
        monkeypatch.setattr(mod, "_lib", stand_in)
    host_kernels.set_thread_order(stand_in.lib, order_seed)
    g = _load(path)
    for with_tangents in (False, True):
        out = gt.GShell_Tets(with_tangents=with_tangents)._marching_from_auggrid(g["pos"], g["sdf"], g["tets"], g["sorted_edges"],
                                                                                 g["coeff"], g["disc"], g["msdf_sign"], g["occ"])
        _compare(out, g, with_tangents=with_tangents, exact_order=False)
        assert (out[4] is None) == (not with_tangents)
    assert out[1].dtype == torch.int64 and out[6].dtype == torch.int64


def test_product_decode_rejects_foreign_edge_lists(host_kernels_lib, monkeypatch):
    stand_in, _ = host_kernels_lib
    import gshell_b200.geometry.gshell_tets as gt
    monkeypatch.setattr(gt, "_lib", stand_in)
    g = _load(GOLDENS[-1])
    bad = g["sorted_edges"].clone()
    bad[0, 0] = bad[0, 0].flip(0)
    # a fresh index tensor: the check is cached per table
    with pytest.raises(ValueError):
        gt.GShell_Tets(with_tangents=False)._marching_from_auggrid(g["pos"], g["sdf"], g["tets"].clone(), bad, g["coeff"], g["disc"],
                                                                   g["msdf_sign"], g["occ"])


@pytest.mark.parametrize("n,seed,kind", [(5, 11, "rand"), (6, 12, "rand"), (7, 13, "sphere"), (8, 14, "rand"), (2, 15, "rand")])
def test_product_decode_kernels_match_oracle_on_fresh_grids(n, seed, kind, host_kernels_lib, monkeypatch):
    """More sizes and seeds than the fixtures hold: the host-compiled kernels against the oracle restatement (itself pinned to the
    reference's goldens above) on inputs built the way the golden generator builds them -- topology exact, floats to rounding."""
    import sys
    from oracle.mt_oracle import gshell_marching_from_auggrid
    sys.path.insert(0, os.path.join(HERE, "golden"))
    try:
        import make_golden_auggrid as mk
    finally:
        sys.path.remove(os.path.join(HERE, "golden"))
    stand_in, host_kernels = host_kernels_lib
    import gshell_b200.geometry.gshell_tets as gt
    import gshell_b200.geometry.tangents as tg
    import gshell_b200.render.mesh as mesh
    for mod in (gt, tg, mesh):
        monkeypatch.setattr(mod, "_lib", stand_in)
    host_kernels.set_thread_order(stand_in.lib, seed)
    a = mk.inputs(n, seed, kind)
    args = (a["pos"], a["sdf"], a["tets"], a["sorted_edges"], a["coeff"], a["disc"], a["msdf_sign"], a["occ"])
    want = gshell_marching_from_auggrid(*args, with_tangents=False)
    got = gt.GShell_Tets(with_tangents=False)._marching_from_auggrid(*args)
    assert want[0].shape[0] > 0
    assert torch.equal(got[1].long(), want[1].long()) and torch.equal(got[6].long(), want[6].long())
    for i in (0, 5, 7, 8):
        assert got[i].shape == want[i].shape and float((got[i] - want[i]).abs().max()) <= 1e-6, i


def test_geometry_decodes_generated_grid(host_kernels_lib, monkeypatch, tmp_path):
    """`GShellTetsGeometry(extract_from_generative=True).getMesh_from_augmented_grid_withocc` (reference gshell_tets_geometry.py:
    64-78, 166-189): the lattice discretisation and per-tet edge list the geometry derives from the grid file, the decode, the
    normals and the tangent frame of the returned mesh -- on the CPU through the host-compiled kernels, against the oracle."""
    import sys
    from oracle.mt_oracle import gshell_marching_from_auggrid, smooth_normals
    sys.path.insert(0, os.path.join(HERE, "golden"))
    try:
        import make_golden_auggrid as mk
    finally:
        sys.path.remove(os.path.join(HERE, "golden"))
    stand_in, host_kernels = host_kernels_lib
    import gshell_b200.geometry.gshell_tets as gt
    import gshell_b200.geometry.gshell_tets_geometry as gg
    import gshell_b200.geometry.tangents as tg
    import gshell_b200.render.mesh as mesh
    from gshell_b200.grids import save_tets_npz
    for mod in (gt, tg, mesh):
        monkeypatch.setattr(mod, "_lib", stand_in)
    # the public method refuses CPU tensors (no CPU path in the product); the test reaches the same code one level below
    monkeypatch.setattr(gt.GShell_Tets, "marching_from_auggrid", lambda self, *a: self._marching_from_auggrid(*a))
    host_kernels.set_thread_order(stand_in.lib, 0)
    n = 4
    npz = str(tmp_path / "tets.npz")
    save_tets_npz(npz, n)
    geo = gg.GShellTetsGeometry(64, 2.0, gg.default_flags(), tet_init_file=npz, extract_from_generative=True, device="cpu")
    a = mk.inputs(n, 21, "rand")
    assert torch.equal(geo.verts_discretized, a["disc"]) and torch.equal(geo.sorted_tetedges, a["sorted_edges"])
    out = geo.getMesh_from_augmented_grid_withocc(None, a["sdf"], a["coeff"], a["msdf_sign"], a["occ"])
    want = gshell_marching_from_auggrid(geo.verts, a["sdf"], geo.indices, geo.sorted_tetedges, a["coeff"], geo.verts_discretized,
                                        a["msdf_sign"], a["occ"], with_tangents=False)
    im = out["imesh"]
    assert torch.equal(im.t_pos_idx.long(), want[1].long())
    assert float((im.v_pos - want[0]).abs().max()) <= 1e-6 and float((out["v_msdf"] - want[7]).abs().max()) == 0
    nrm = smooth_normals(want[0], want[1])
    agree = ((im.v_nrm - nrm).abs().max(-1).values < 1e-4).float().mean()
    assert float(agree) > 0.95                                   # the rest: unreferenced / cancelling vertices (fallback normal)
    unit = im.v_tng.norm(dim=-1)
    assert bool(torch.isfinite(im.v_tng).all()) and float(((unit - 1).abs() < 1e-3).float().mean()) > 0.9
    assert float((im.v_tng * im.v_nrm).sum(-1).abs()[(unit - 1).abs() < 1e-3].max()) < 1e-3      # orthogonal to the vertex normals
