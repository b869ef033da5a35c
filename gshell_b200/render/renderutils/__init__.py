"""Drop-in for the reference's `render/renderutils` package: the three operators the G-Shell training
path calls (reference call sites render/render.py:118,374 and createLoss in the train scripts)."""
from .ops import xfm_points, prepare_shading_normal, image_loss

__all__ = ["xfm_points", "prepare_shading_normal", "image_loss"]
