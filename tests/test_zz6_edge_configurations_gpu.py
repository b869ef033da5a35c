"""tick() on the shapes and settings the other suites do not use: image sizes that are no multiple of any kernel tile (33 x 47,
17 x 9, 4 x 4), 1 / 3 views, odd sample counts (n = 1, 3, 5), non-square light probes, the shadow ramp half-way, no denoiser, no
background, a camera INSIDE the shell (vertices behind the near plane, w <= 0), the random-SDF soup, FlexiCubes, multisampling.
Asserted: nothing raises, the loss and every parameter gradient are finite, the SDF receives a gradient.  (Written while driving
the pipeline on the host emulator, which is how the spp > 1 composite bug was found; sorts last.)"""
import os

import numpy as np
import pytest
import torch

from _device import DEVICE, device      # cuda:0, or the CPU under the host emulator (tests/_device.py)

pytestmark = pytest.mark.gpu

CASES = [dict(B=1, res=[33, 47], n=1, it=0, probe=(16, 16)),
         dict(B=3, res=[17, 9], n=3, it=500, probe=(16, 32)),
         dict(B=2, res=[4, 4], n=2, it=1500, probe=(32, 16)),
         dict(B=1, res=[64, 24], n=5, it=1500, probe=(16, 16), den=False),
         dict(B=2, res=[40, 40], n=2, it=1500, probe=(64, 64), bg=False),
         dict(B=1, res=[31, 31], n=2, it=1500, probe=(16, 16), inside=True),
         dict(B=2, res=[24, 56], n=4, it=700, probe=(16, 16), kind="flex"),
         dict(B=1, res=[25, 25], n=2, it=1500, probe=(16, 16), sdf="random"),
         dict(B=2, res=[20, 28], n=2, it=1500, probe=(16, 16), spp=2)]


@pytest.mark.parametrize("c", CASES, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items() if k not in ("probe",)).replace(" ", ""))
def test_tick_runs_and_stays_finite(c, tmp_path):
    from gshell_b200 import synthetic
    from gshell_b200.denoiser.denoiser import BilateralDenoiser
    from gshell_b200.geometry.gshell_flexicubes_geometry import GShellFlexiCubesGeometry
    from gshell_b200.geometry.gshell_tets_geometry import GShellTetsGeometry, default_flags
    from gshell_b200.grids import save_tets_npz
    from gshell_b200.render import light, util
    from gshell_b200.render import renderutils as ru
    d = device()
    torch.manual_seed(0)
    FLAGS = default_flags(n_samples=c["n"], sphere_init=c.get("sdf") != "random")
    if c.get("kind") == "flex":
        geo = GShellFlexiCubesGeometry(8, 2.0, FLAGS, device=d)
    else:
        npz = str(tmp_path / "tets.npz")
        save_tets_npz(npz, 6)
        geo = GShellTetsGeometry(64, 2.0, FLAGS, tet_init_file=npz, device=d)
    B, res, spp = c["B"], c["res"], c.get("spp", 1)
    mat = synthetic.LeafMaterialField(B, res[0], res[1], d)
    lgt = light.EnvironmentLight((torch.rand(*c["probe"], 3, device=d) * 0.5 + 0.25).requires_grad_())
    mvp, campos = synthetic.random_cameras(B, res, d, np.random.RandomState(1))
    if c.get("inside"):
        mv = util.translate(0, 0, -0.3)
        mvp = (util.perspective(0.8, 1.0, 0.1, 1000.0) @ mv)[None].to(d)
        campos = torch.linalg.inv(mv)[:3, 3][None].to(d)
    img, bg = synthetic.random_target(B, res, d)
    target = {"mvp": mvp, "campos": campos, "img": img, "background": bg if c.get("bg", True) else None, "resolution": res, "spp": spp}
    lgt.update_pdf()
    il, dl, rl = geo.tick(None, target, lgt, {"kd_ks": mat, "bsdf": "pbr"}, lambda a, b: ru.image_loss(a, b, loss="l1", tonemapper="log_srgb"),
                          c["it"], BilateralDenoiser() if c.get("den", True) else None)
    total = il + dl + rl
    assert bool(torch.isfinite(total))
    total.backward()
    for name, p in (("sdf", geo.sdf), ("msdf", geo.msdf), ("deform", geo.deform), ("light", lgt.base), ("material", mat.tex)):
        assert p.grad is not None and bool(torch.isfinite(p.grad).all()), name
    assert float(geo.sdf.grad.abs().sum()) > 0


@pytest.mark.parametrize("kind", ["tets", "flex"])
def test_tick_on_a_field_without_a_surface(kind, tmp_path):
    """SDF > 0 everywhere: the extraction returns size-0 tensors ("must return size-0 tensors, not raise"), every later
    stage gets an empty mesh -- no launch with an empty grid, no exception; the image loss is the background's, finite.  (The SDF
    regulariser is a mean over the sign-changing edges, of which there are none: NaN, as in the reference.)"""
    from gshell_b200 import synthetic
    from gshell_b200.denoiser.denoiser import BilateralDenoiser
    from gshell_b200.geometry.gshell_flexicubes_geometry import GShellFlexiCubesGeometry
    from gshell_b200.geometry.gshell_tets_geometry import GShellTetsGeometry, default_flags
    from gshell_b200.grids import save_tets_npz
    from gshell_b200.render import light
    from gshell_b200.render import renderutils as ru
    d = device()
    FLAGS = default_flags(n_samples=2, sphere_init=True)
    if kind == "flex":
        geo = GShellFlexiCubesGeometry(8, 2.0, FLAGS, device=d)
    else:
        npz = str(tmp_path / "tets.npz")
        save_tets_npz(npz, 6)
        geo = GShellTetsGeometry(64, 2.0, FLAGS, tet_init_file=npz, device=d)
    with torch.no_grad():
        geo.sdf.fill_(1.0)
    B, res = 1, [32, 32]
    mat = synthetic.LeafMaterialField(B, res[0], res[1], d)
    lgt = light.create_trainable_env_rnd(16, device=d)
    mvp, campos = synthetic.random_cameras(B, res, d, np.random.RandomState(1))
    img, bg = synthetic.random_target(B, res, d)
    target = {"mvp": mvp, "campos": campos, "img": img, "background": bg, "resolution": res, "spp": 1}
    md = geo.getMesh({"kd_ks": mat, "bsdf": "pbr"})
    assert md["imesh"].v_pos.shape == (0, 3) and md["imesh"].t_pos_idx.shape == (0, 3)
    il, dl, rl = geo.tick(None, target, lgt, {"kd_ks": mat, "bsdf": "pbr"}, lambda a, b: ru.image_loss(a, b, loss="l1", tonemapper="log_srgb"),
                          1200, BilateralDenoiser())
    assert bool(torch.isfinite(il)) and bool(torch.isfinite(dl))
    (il + dl).backward()
