"""`render.render.render_uv` (reference render/render.py:449-468; texture baking after training, `xatlas_uvmap` in the train
scripts): rasterisation in texture space + interpolation of the world positions + one sample of the material field, against the
rasteriser oracle.  Sorts last: added after the round's last GPU run (it only composes operators that the established suites
cover, tests/test_raster_gpu.py)."""
import pytest
import torch

from _device import DEVICE, device      # cuda:0, or the CPU under the host emulator (tests/_device.py)

pytestmark = pytest.mark.gpu


def test_render_uv_matches_oracle_composition():
    from gshell_b200.render import mesh, render
    from oracle import raster_oracle
    d = device()
    g = torch.Generator().manual_seed(0)
    # two charts in uv space, a bent quad in world space; uv and position index buffers differ, as after a uv unwrap
    v_pos = torch.tensor([[-1.0, -1.0, 0.0], [1.0, -1.0, 0.3], [1.0, 1.0, 0.0], [-1.0, 1.0, -0.2]]) + 0.05 * torch.randn(4, 3, generator=g)
    t_pos = torch.tensor([[0, 1, 2], [0, 2, 3]])
    v_tex = torch.tensor([[0.05, 0.05], [0.45, 0.07], [0.44, 0.46], [0.55, 0.52], [0.95, 0.9], [0.53, 0.93]])
    t_tex = torch.tensor([[0, 1, 2], [3, 4, 5]])
    m = mesh.Mesh(v_pos.to(d), t_pos.to(d), v_tex=v_tex.to(d), t_tex_idx=t_tex.to(d))

    class Field:
        def sample(self, p):
            return torch.cat([p * 0.5 + 0.5, torch.sin(p)], -1)
    H = W = 48
    mask, kd, ks = render.render_uv(None, m, [H, W], Field())
    uv = v_tex[None] * 2.0 - 1.0
    clip = torch.cat([uv, torch.zeros(1, 6, 1), torch.ones(1, 6, 1)], -1)
    rast = raster_oracle.rasterize(clip, t_tex, H, W)
    pos = raster_oracle.interpolate(v_pos[None], rast, t_pos)
    want_mask = (rast[..., 3:] > 0).float()
    assert mask.shape == (1, H, W, 1) and kd.shape == (1, H, W, 3) and ks.shape == (1, H, W, 3)
    assert torch.equal(mask.cpu(), want_mask) and 0.1 < float(want_mask.mean()) < 0.25
    cov = want_mask[..., 0] > 0
    assert float((kd.cpu() - (pos * 0.5 + 0.5))[cov].abs().max()) < 1e-5
    assert float((ks.cpu() - torch.sin(pos))[cov].abs().max()) < 1e-5


def test_environment_light_generate_image():
    """EnvironmentLight.generate_image (reference light.py:61-64; validation images): the probe resampled with wrapping bilinear
    taps; at the probe's own resolution it is the probe."""
    from gshell_b200.render import light
    lgt = light.create_trainable_env_rnd(16, device=device())
    img = lgt.generate_image([16, 16])
    assert img.shape == (16, 16, 3) and torch.allclose(img, lgt.base.detach(), atol=1e-6)
    big = lgt.generate_image([32, 64])
    assert big.shape == (32, 64, 3) and not big.requires_grad
