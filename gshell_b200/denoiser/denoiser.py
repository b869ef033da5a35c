"""`BilateralDenoiser` with the reference's interface (denoiser/denoiser.py:21-35): a cross-bilateral filter of the Monte-Carlo
light buffers guided by normals and depth; the filter itself is csrc/denoise.cu behind `render.optixutils`."""
import math

import torch

from ..render import optixutils as ou


def _unit(v, eps=1e-20):
    """v / |v| with the squared length clamped (bent normals reach the filter shorter than 1)"""
    return v / torch.sqrt((v * v).sum(-1, keepdim=True).clamp(min=eps))


class BilateralDenoiser(torch.nn.Module):
    def __init__(self, influence=1.0):
        super().__init__()
        self.set_influence(influence)

    def set_influence(self, factor):
        """Filter width from the shadow ramp of `tick` (reference :26-29): sigma = 2 factor, window of 2 ceil(2.5 sigma) + 1 taps."""
        sigma = max(2 * factor, 1e-4)
        self.sigma, self.variance, self.N = sigma, sigma ** 2.0, 2 * math.ceil(2.5 * sigma) + 1

    def forward(self, input):
        """input [B,H,W,8] = (rgb, normal, z, dz), the layout `render.shade` assembles (reference render.py:141)."""
        rgb, normal, depth = input[..., 0:3], input[..., 3:6], input[..., 6:8]
        return ou.bilateral_denoiser(rgb, _unit(normal), depth, self.sigma)

    def forward_pair(self, col_a, col_b, normal, depth):
        """Diffuse and specular light in one pass over shared guides (one kernel launch instead of two)."""
        return ou.bilateral_denoiser_pair(col_a, col_b, _unit(normal), depth, self.sigma)
