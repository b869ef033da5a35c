"""Host-side helpers of gshell_b200/render/util.py that have no kernel behind them."""
import torch

from gshell_b200.render import util


def test_texture_linear_wrap_conventions():
    """nvdiffrast's `texture(..., filter_mode='linear')` conventions as EnvironmentLight.generate_image relies on them
    (reference light.py:61-64): texel centres at (i + 0.5) / N, bilinear blend, coordinates wrap."""
    g = torch.Generator().manual_seed(0)
    tex = torch.rand(5, 7, 3, generator=g)
    same = util.texture_linear_wrap(tex, util.pixel_grid(7, 5, device="cpu"))
    assert torch.allclose(same, tex, atol=1e-6)                                   # sampling at the texel centres returns the texels
    uv = torch.tensor([[[1.0 / 7, 0.5 / 5]]])                                      # half-way between texels (0,0) and (0,1)
    assert torch.allclose(util.texture_linear_wrap(tex, uv)[0, 0], 0.5 * (tex[0, 0] + tex[0, 1]), atol=1e-6)
    left_edge = util.texture_linear_wrap(tex, torch.tensor([[[0.0, 0.5 / 5]]]))[0, 0]       # x = 0 blends the first and the LAST column
    assert torch.allclose(left_edge, 0.5 * (tex[0, 0] + tex[0, 6]), atol=1e-6)
    up = util.texture_linear_wrap(tex, util.pixel_grid(14, 10, device="cpu"))
    assert up.shape == (10, 14, 3) and float(up.min()) >= float(tex.min()) - 1e-6 and float(up.max()) <= float(tex.max()) + 1e-6
    assert abs(float(up.mean()) - float(tex.mean())) < 1e-5                       # wrapping bilinear upsampling by 2 keeps the mean


def test_geometry_modules_expose_the_reference_parameter_names(tmp_path):
    """The train scripts pick the optimiser groups by parameter NAME (train_gshelltet_deepfashion.py:327-330: 'deform', 'msdf',
    'sdf', the rest) and store `geometry.state_dict()` (:691).  Names as registered by the reference: gshell_tets_geometry.py:88-144,
    gshell_flexicubes_geometry.py:68-106 (the per-cube weights under two names)."""
    from gshell_b200.geometry.gshell_flexicubes_geometry import GShellFlexiCubesGeometry
    from gshell_b200.geometry.gshell_tets_geometry import GShellTetsGeometry, default_flags
    from gshell_b200.grids import save_tets_npz
    npz = str(tmp_path / "tets.npz")
    save_tets_npz(npz, 3)
    tets = GShellTetsGeometry(64, 2.0, default_flags(), tet_init_file=npz, device="cpu")
    assert sorted(n for n, _ in tets.named_parameters()) == ["deform", "msdf", "sdf"]
    assert sorted(tets.state_dict()) == ["deform", "msdf", "sdf"]
    flex = GShellFlexiCubesGeometry(4, 2.0, default_flags(), device="cpu")
    assert sorted(n for n, _ in flex.named_parameters()) == ["deform", "msdf", "per_cube_weights", "sdf"]
    assert sorted(flex.state_dict()) == ["deform", "msdf", "per_cube_weights", "sdf", "weight"]
    flex.load_state_dict(flex.state_dict())
    mlp = GShellTetsGeometry(64, 2.0, default_flags(use_sdf_mlp=True, sdf_mlp_pretrain_steps=1, n_hidden=2, d_hidden=8), tet_init_file=npz,
                             device="cpu")
    names = [n for n, _ in mlp.named_parameters()]
    assert {"deform", "msdf", "sdf"} <= set(names) and any(n.startswith("sdf_net.") for n in names)
