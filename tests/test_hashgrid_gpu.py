"""csrc/hashgrid.cu through the C ABI against oracle/hashgrid_oracle.py (same formulas in torch; tiny-cuda-nn itself is absent:
parity unpinned), and the `MLPTexture3D` surface of the reference's material field."""
import pytest
import torch

from gshell_b200.render import mlptexture
from oracle import hashgrid_oracle as ho

from _device import DEVICE, device      # cuda:0, or the CPU under the host emulator (tests/_device.py)

pytestmark = pytest.mark.gpu


def _points(n, gen):
    x = torch.rand(n, 3, generator=gen)
    x[:8] = torch.tensor([[0, 0, 0], [1, 1, 1], [1, 0, 0], [0, 1, 0], [0, 0, 1], [0.5, 0.5, 0.5], [1, 1, 0], [0.25, 1, 0.75]])
    return x


@pytest.mark.parametrize("cfg", [{}, {"n_levels": 6, "base_resolution": 4, "desired_resolution": 128, "log2_hashmap_size": 10}])
def test_encoding_forward_and_backward_match_oracle(cfg):
    dev = device()
    gen = torch.Generator().manual_seed(3)
    layout = mlptexture.hashgrid_levels(**cfg)
    offs, ress, scales = [a.tolist() for a in layout]
    table = torch.randn(offs[-1], 2, generator=gen)
    x = _points(4096, gen)
    w = torch.randn(4096, 2 * len(ress), generator=gen)
    xo, to = x.clone().requires_grad_(True), table.clone().requires_grad_(True)
    want = ho.encode(xo, to, offs, ress, scales)
    (want * w).sum().backward()
    xc, tc = x.to(dev).requires_grad_(True), table.to(dev).requires_grad_(True)
    got = mlptexture._HashGrid.apply(xc, tc, layout)
    (got * w.to(dev)).sum().backward()
    assert got.shape == want.shape
    assert float((got.cpu() - want).abs().max()) <= 2e-6 * float(want.abs().max())
    gt, gto = tc.grad.cpu(), to.grad
    assert float((gt - gto).norm()) <= 1e-5 * float(gto.norm())
    # gradient w.r.t. the position: piecewise constant in a cell, scale up to 4095 -> relative to its own size per level mix
    gx, gxo = xc.grad.cpu(), xo.grad
    assert float((gx - gxo).norm()) <= 1e-4 * float(gxo.norm())


def test_material_field_surface():
    dev = device()
    torch.manual_seed(0)
    aabb = torch.tensor([[-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]], device=dev)
    mn = torch.tensor([0.0, 0.0, 0.0, 0.0, 0.08, 0.0], device=dev)
    mx = torch.tensor([1.0, 1.0, 1.0, 0.0, 1.0, 1.0], device=dev)
    tex = mlptexture.MLPTexture3D(aabb, channels=6, min_max=[mn, mx])
    pos = (torch.rand(2, 16, 16, 3, device=dev) * 2.4 - 1.2).requires_grad_(True)           # some points outside the box: clamped
    out = tex.sample(pos)
    assert out.shape == (2, 16, 16, 6)
    assert bool((out >= mn - 1e-6).all()) and bool((out <= mx + 1e-6).all())
    out.sum().backward()
    assert float(tex.encoder.params.grad.abs().sum()) > 0 and all(float(p.grad.abs().sum()) > 0 for p in tex.net.parameters())
    assert pos.grad is not None and bool(torch.isfinite(pos.grad).all())
    # the encoder's parameters see the gradient scaled by 128 (reference mlptexture.py:30,76); the positions do not
    g_params = tex.encoder.params.grad.clone()
    tex.zero_grad()
    pos.grad = None
    tex.gradient_scaling = 1.0
    tex.sample(pos).sum().backward()
    assert float((g_params - 128.0 * tex.encoder.params.grad).norm()) <= 1e-4 * float(g_params.norm())
    tex.clamp_()
    tex.cleanup()
