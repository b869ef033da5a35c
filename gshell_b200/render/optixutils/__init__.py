"""Drop-in for the reference's `render/optixutils` package (render/optixutils/ops.py:128-147), without OptiX."""
from .ops import OptiXContext, optix_build_bvh, optix_env_shade, bilateral_denoiser, bilateral_denoiser_pair

__all__ = ["OptiXContext", "optix_build_bvh", "optix_env_shade", "bilateral_denoiser", "bilateral_denoiser_pair"]
