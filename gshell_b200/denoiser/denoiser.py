"""Drop-in for the reference's denoiser/denoiser.py (BilateralDenoiser :21-35)."""
import math

import torch

from ..render import optixutils as ou


def _safe_normalize(x, eps=1e-20):
    return x / torch.sqrt(torch.clamp(torch.sum(x * x, -1, keepdim=True), min=eps))


class BilateralDenoiser(torch.nn.Module):
    def __init__(self, influence=1.0):
        super().__init__()
        self.set_influence(influence)

    def set_influence(self, factor):
        self.sigma = max(factor * 2, 0.0001)
        self.variance = self.sigma ** 2.
        self.N = 2 * math.ceil(self.sigma * 2.5) + 1

    def forward(self, input):
        """input [B,H,W,8] = (rgb, normal, z, dz) as assembled at reference render.py:141."""
        col = input[..., 0:3]
        nrm = _safe_normalize(input[..., 3:6])      # bent normals can be shorter than 1
        zdz = input[..., 6:8]
        return ou.bilateral_denoiser(col, nrm, zdz, self.sigma)

    def forward_pair(self, col_a, col_b, normal, depth):
        """Diffuse + specular light in one pass (shared guides)."""
        return ou.bilateral_denoiser_pair(col_a, col_b, _safe_normalize(normal), depth, self.sigma)
