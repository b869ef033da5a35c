"""Synthetic workload pieces shared by bench.py, smoke() and the tests: cameras (recipe of the reference's
dataset/dataset_mesh.py:65-87), a leaf-tensor material field standing in for the tcnn hash-grid MLP
(render/mlptexture.py, third-party field, out of scope), targets."""
import numpy as np
import torch

from .render import util


class LeafMaterialField:
    """`material['kd_ks'].sample(pos) -> [B,H,W,6]` backed by a per-pixel leaf tensor (kd rgb, ks = (o, roughness,
    metallic)); gradients land in `.tex.grad` (= d/dkd, d/dks buffers)."""

    def __init__(self, B, H, W, device, generator=None):
        kd = torch.rand(B, H, W, 3, generator=generator)
        ks = torch.stack([torch.zeros(B, H, W), 0.08 + 0.92 * torch.rand(B, H, W, generator=generator),
                          torch.rand(B, H, W, generator=generator)], -1)
        self.tex = torch.cat([kd, ks], -1).to(device).requires_grad_()

    def sample(self, pos):
        return self.tex

    def parameters(self):
        return [self.tex]


def random_cameras(B, res, device, rng, fovy=np.deg2rad(45), cam_radius=3.0, cam_near_far=(0.1, 1000.0)):
    """mv = translate(0,0,-r) @ random_rotation_translation(0.25); mvp = proj @ mv; campos = inv(mv)[:3,3]."""
    proj = util.perspective(fovy, res[1] / res[0], cam_near_far[0], cam_near_far[1])
    mvs = [util.translate(0, 0, -cam_radius) @ util.random_rotation_translation(0.25, rng=rng) for _ in range(B)]
    mv = torch.stack(mvs, 0)
    mvp = proj[None] @ mv
    campos = torch.linalg.inv(mv)[:, :3, 3]
    return mvp.to(device), campos.to(device)


def random_target(B, res, device, generator=None):
    """RGBA target with a binary alpha disc + random background."""
    H, W = res
    img = torch.rand(B, H, W, 4, generator=generator)
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, H), torch.linspace(-1, 1, W), indexing="ij")
    img[..., 3] = ((xx * xx + yy * yy) < 0.45).float()[None]
    bg = torch.rand(B, H, W, 3, generator=generator)
    return img.to(device), bg.to(device)
