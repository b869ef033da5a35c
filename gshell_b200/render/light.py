"""Monte-Carlo sampled environment light with PDF / CDF tables (reference render/light.py:21-105)."""
import numpy as np
import torch

from . import util


class EnvironmentLight:
    LIGHT_MIN_RES = 16
    MIN_ROUGHNESS = 0.08
    MAX_ROUGHNESS = 0.5

    def __init__(self, base):
        self.mtx = None
        self.base = base
        self.pdf_scale = (self.base.shape[0] * self.base.shape[1]) / (2 * np.pi * np.pi)
        self.update_pdf()

    def xfm(self, mtx):
        self.mtx = mtx

    def parameters(self):
        return [self.base]

    def clone(self):
        return EnvironmentLight(self.base.clone().detach())

    def clamp_(self, min=None, max=None):
        self.base.clamp_(min, max)

    @torch.no_grad()
    def update_pdf(self):
        """pdf = max-channel * sin(theta), row / column CDFs (reference light.py:46-59).  256x256 texels: three
        cumsum-sized torch ops per iteration, not worth a kernel ('tiny')."""
        h, w = self.base.shape[0], self.base.shape[1]
        Y = util.pixel_grid(w, h, device=self.base.device)[..., 1]
        pdf = torch.max(self.base, dim=-1)[0] * torch.sin(Y * np.pi)
        self._pdf = pdf / torch.sum(pdf)
        self.cols = torch.cumsum(self._pdf, dim=1)
        self.rows = torch.cumsum(self.cols[:, -1:].repeat([1, self.cols.shape[1]]), dim=0)
        self.cols = self.cols / torch.where(self.cols[:, -1:] > 0, self.cols[:, -1:], torch.ones_like(self.cols))
        self.rows = self.rows / torch.where(self.rows[-1:, :] > 0, self.rows[-1:, :], torch.ones_like(self.rows))


def create_trainable_env_rnd(base_res, scale=0.5, bias=0.25, device="cuda"):
    base = torch.rand(base_res, base_res, 3, dtype=torch.float32, device=device) * scale + bias
    return EnvironmentLight(base.clone().detach().requires_grad_(True))
