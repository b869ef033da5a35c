"""GPU parity of the CUDA marching-tets path (through the C ABI) against
  (1) golden outputs of the unmodified reference (tests/golden/mt_*.npz),
  (2) the CPU oracle on larger seeded grids,
  (3) size-independent properties at the BASELINE "256" grid size (BCC N=103).
Bars: faces bit-exact; vertex positions / mSDF values exact (same IEEE op order, tolerance fallback
1e-4 relative as stated by north_star); gradients within 1e-4 relative to the largest entry."""
import glob
import os

import numpy as np
import pytest
import torch

from _device import DEVICE, device      # cuda:0, or the CPU under the host emulator (tests/_device.py)

pytestmark = pytest.mark.gpu

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "mt_*.npz")))


def _dev():
    return device()


def _run_cuda(pos, sdf, msdf, tets, index_dtype=torch.int64):
    from gshell_b200.geometry.gshell_tets import GShell_Tets
    dev = _dev()
    leaves = [x.clone().to(dev).requires_grad_() for x in (pos, sdf, msdf)]
    out = GShell_Tets(index_dtype=index_dtype, with_tangents=False)(*leaves, tets.to(dev))
    return leaves, out


def _check_forward(out, want, exact=True):
    va, fa, _, _, _, extra = out
    assert fa.dtype == torch.int64
    assert torch.equal(fa.cpu(), want["faces_aug"])
    assert torch.equal(extra["faces_watertight"].cpu(), want["faces_watertight"])
    assert extra["n_verts_watertight"] == int(want["n_verts_watertight"])
    assert out[4] is None          # built with_tangents=False, like the training path; the tangent frame: tests/test_zz2_tangents_gpu.py
    for got, key in ((va, "verts_aug"), (extra["vertices_watertight"], "vertices_watertight"),
                     (extra["msdf"], "msdf_aug"), (extra["msdf_watertight"], "msdf_watertight"),
                     (extra["msdf_boundary"], "msdf_boundary")):
        if exact:
            assert torch.equal(got.detach().cpu(), want[key]), key
        else:
            torch.testing.assert_close(got.detach().cpu(), want[key], rtol=1e-4, atol=1e-6)


def _check_grads(leaves, out, w, want_grads):
    va, _, _, _, _, extra = out
    dev = va.device
    probe = (va * w["wa"].to(dev)).sum() + (extra["msdf"] * w["wm"].to(dev)).sum() + \
            (extra["vertices_watertight"] * w["ww"].to(dev)).sum()
    grads = torch.autograd.grad(probe, leaves, allow_unused=True)
    for name, got, want in zip(("pos", "sdf", "msdf"), grads, want_grads):
        scale = want.abs().max().clamp(min=1.0)
        err = (got.cpu() - want).abs().max()
        assert err <= 1e-4 * scale, f"grad {name}: err {err} scale {scale}"


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[3:-4] for p in GOLDEN])
def test_cuda_matches_reference_golden(path):
    z = np.load(path)
    g = {k: torch.from_numpy(z[k]) if z[k].shape != () else z[k] for k in z.files}
    leaves, out = _run_cuda(g["pos"], g["sdf"], g["msdf"], g["tets"])
    _check_forward(out, g)
    if out[0].shape[0]:
        _check_grads(leaves, out, g, [g["gmain_pos"], g["gmain_sdf"], g["gmain_msdf"]])


# (52, ...) and (103, ...) are BASELINE.json's "128" and "256" grids (configs[1], configs[3]); the random-SDF field of the
# benchmark uses msdf = U(0,1) - 0.01 ("bench")
@pytest.mark.parametrize("n,seed,kind", [(16, 11, "rand"), (26, 12, "rand"), (26, 13, "sphere"), (33, 14, "rand"),
                                         (52, 15, "bench"), (52, 16, "sphere"), (103, 0, "bench"), (103, 17, "sphere")])
def test_cuda_matches_oracle(n, seed, kind):
    from gshell_b200.grids import bcc_tet_grid
    from oracle.mt_oracle import gshell_marching_tets
    v, t = bcc_tet_grid(n)
    g = torch.Generator().manual_seed(seed)
    pos = torch.tensor(v) - 0.5
    nv = v.shape[0]
    if kind == "rand":
        sdf = torch.rand(nv, generator=g) - 0.1
        msdf = (torch.rand(nv, generator=g) - 0.3).clamp(-1, 1)
    elif kind == "bench":
        sdf = torch.rand(nv, generator=g) - 0.1
        msdf = (torch.rand(nv, generator=g) - 0.01).clamp(-1, 1)
    else:
        sdf = pos.norm(dim=1) - 0.3 + 0.002 * torch.rand(nv, generator=g)
        msdf = pos[:, 2] + 0.1
    tets = torch.tensor(t)
    ol = [x.clone().requires_grad_() for x in (pos, sdf, msdf)]
    ova, ofa, _, _, _, oex = gshell_marching_tets(*ol, tets, unique_mode="packed", with_tangents=False)
    want = {"faces_aug": ofa, "faces_watertight": oex["faces_watertight"],
            "n_verts_watertight": oex["n_verts_watertight"], "verts_aug": ova.detach(),
            "vertices_watertight": oex["vertices_watertight"].detach(), "msdf_aug": oex["msdf"].detach(),
            "msdf_watertight": oex["msdf_watertight"].detach(), "msdf_boundary": oex["msdf_boundary"].detach()}
    leaves, out = _run_cuda(pos, sdf, msdf, tets)
    _check_forward(out, want)
    gw = torch.Generator().manual_seed(seed + 100)
    w = {"wa": torch.randn(ova.shape, generator=gw), "wm": torch.randn(oex["msdf"].shape, generator=gw),
         "ww": torch.randn(oex["vertices_watertight"].shape, generator=gw)}
    probe = (ova * w["wa"]).sum() + (oex["msdf"] * w["wm"]).sum() + (oex["vertices_watertight"] * w["ww"]).sum()
    want_grads = torch.autograd.grad(probe, ol)
    _check_grads(leaves, out, w, want_grads)


def test_int32_faces_option_and_determinism():
    from gshell_b200.grids import bcc_tet_grid
    v, t = bcc_tet_grid(20)
    torch.manual_seed(3)
    pos, sdf = torch.tensor(v), torch.rand(v.shape[0]) - 0.2
    msdf = torch.rand(v.shape[0]) - 0.4
    _, a = _run_cuda(pos, sdf, msdf, torch.tensor(t), index_dtype=torch.int32)
    _, b = _run_cuda(pos, sdf, msdf, torch.tensor(t), index_dtype=torch.int64)
    assert a[1].dtype == torch.int32 and torch.equal(a[1].long(), b[1])
    assert torch.equal(a[0], b[0]) and torch.equal(a[5]["msdf"], b[5]["msdf"])


def test_full_size_properties():
    """BASELINE '256' grid (BCC N=103: 2.22M verts / 12.99M tets): structural invariants that do not
    need the (minutes-long) CPU oracle."""
    from gshell_b200.grids import bcc_tet_grid
    v, t = bcc_tet_grid(103)
    dev = _dev()
    g = torch.Generator().manual_seed(0)
    pos = (torch.tensor(v) - 0.5).to(dev)
    sdf = (torch.rand(v.shape[0], generator=g) - 0.1).to(dev).requires_grad_()
    msdf = (torch.rand(v.shape[0], generator=g) - 0.01).clamp(-1, 1).to(dev).requires_grad_()
    tets = torch.tensor(t).to(dev)
    from gshell_b200.geometry.gshell_tets import GShell_Tets
    va, fa, _, _, _, ex = GShell_Tets(index_dtype=torch.int32, with_tangents=False)(pos, sdf, msdf, tets)
    n_wt = ex["n_verts_watertight"]
    occ = sdf > 0
    # number of watertight vertices == number of sign-crossing unique edges
    from gshell_b200.geometry.tet_tables import tables_for
    tab = tables_for(tets, v.shape[0])
    cross = occ[tab.edge_v[:, 0].long()] != occ[tab.edge_v[:, 1].long()]
    assert n_wt == int(cross.sum())
    # face indices in range; every referenced row is non-zero-able and every unreferenced row is zero
    assert int(fa.min()) >= 0 and int(fa.max()) < va.shape[0]
    used = torch.zeros(va.shape[0], dtype=torch.bool, device=dev)
    used[fa.reshape(-1).long()] = True
    assert torch.all(va[~used] == 0)
    assert torch.equal(used[:n_wt], ex["msdf_watertight"] > 0)
    # watertight faces: T1 + 2*T2 rows, each watertight edge shared by exactly two faces away from the grid border
    fw = ex["faces_watertight"].long()
    assert int(fw.min()) >= 0 and int(fw.max()) < n_wt
    # boundary vertices lie on the mSDF zero level: interpolated mSDF ~ 0 where referenced
    mb = ex["msdf_boundary"]
    assert float(mb[used[n_wt:]].abs().max()) < 1e-4
    # backward runs and is finite
    (va.sum() + ex["msdf"].sum()).backward()
    assert torch.isfinite(sdf.grad).all() and torch.isfinite(msdf.grad).all()
    assert float(sdf.grad.abs().sum()) > 0 and float(msdf.grad.abs().sum()) > 0


def test_empty_surface_returns_empty_tensors():
    from gshell_b200.grids import bcc_tet_grid
    v, t = bcc_tet_grid(4)
    _, out = _run_cuda(torch.tensor(v), torch.ones(v.shape[0]), torch.ones(v.shape[0]), torch.tensor(t))
    assert out[0].shape == (0, 3) and out[1].shape == (0, 3) and out[5]["n_verts_watertight"] == 0


OPEN_GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "mtopen_*.npz")))


@pytest.mark.parametrize("path", OPEN_GOLDEN, ids=[os.path.basename(p)[7:-4] for p in OPEN_GOLDEN])
def test_cuda_matches_reference_golden_without_watertight_template(path):
    """output_watertight_template=False (reference gshell_tets.py:260-263, 435-441; goldens of the unmodified reference): other
    vertex numbering, fewer rows, `extra` reduced to its mSDF entries."""
    from gshell_b200.geometry.gshell_tets import GShell_Tets
    z = np.load(path)
    g = {k: torch.from_numpy(z[k]) for k in z.files}
    dev = _dev()
    leaves = [g[k].clone().to(dev).requires_grad_() for k in ("pos", "sdf", "msdf")]
    va, fa, _, _, _, ex = GShell_Tets(with_tangents=False)(*leaves, g["tets"].to(dev), output_watertight_template=False)
    assert set(ex) == {"msdf", "msdf_watertight", "msdf_boundary"}
    assert torch.equal(fa.cpu(), g["faces_aug"]) and torch.equal(va.detach().cpu(), g["verts_aug"])
    for k, w in (("msdf", "msdf_aug"), ("msdf_watertight", "msdf_watertight"), ("msdf_boundary", "msdf_boundary")):
        assert torch.equal(ex[k].detach().cpu(), g[w]), k
    grads = torch.autograd.grad((va * g["wa"].to(dev)).sum() + (ex["msdf"] * g["wm"].to(dev)).sum(), leaves, allow_unused=True)
    for nm, a in zip(("pos", "sdf", "msdf"), grads):
        want = g[f"g_{nm}"]
        got = torch.zeros_like(want) if a is None else a.cpu()
        assert float((got - want).abs().max()) <= 1e-4 * max(1.0, float(want.abs().max())), nm


def test_without_watertight_template_matches_oracle_at_the_128_grid():
    from gshell_b200.geometry.gshell_tets import GShell_Tets
    from gshell_b200.grids import bcc_tet_grid
    from oracle.mt_oracle import gshell_marching_tets
    v, t = bcc_tet_grid(52)
    g = torch.Generator().manual_seed(31)
    pos = torch.tensor(v) - 0.5
    sdf = torch.rand(v.shape[0], generator=g) - 0.1
    msdf = torch.where(pos[:, 0] > 0.05, -torch.rand(v.shape[0], generator=g) - 0.01, torch.rand(v.shape[0], generator=g) - 0.3)
    tets = torch.tensor(t)
    ova, ofa, _, _, _, oex = gshell_marching_tets(pos, sdf, msdf, tets, unique_mode="packed", with_tangents=False,
                                                  output_watertight_template=False)
    dev = _dev()
    va, fa, _, _, _, ex = GShell_Tets(with_tangents=False)(pos.to(dev), sdf.to(dev), msdf.to(dev), tets.to(dev),
                                                          output_watertight_template=False)
    assert torch.equal(fa.cpu(), ofa) and torch.equal(va.cpu(), ova) and torch.equal(ex["msdf"].cpu(), oex["msdf"])
    full = GShell_Tets(with_tangents=False)(pos.to(dev), sdf.to(dev), msdf.to(dev), tets.to(dev))
    assert va.shape[0] < full[0].shape[0]            # the pre-filter really removed tets
