"""Drop-in for the reference's `render/renderutils` package.  The three operators the G-Shell training path calls (reference
call sites render/render.py:118,374 and createLoss in the train scripts) and the pointwise BSDF operators the package also exports
(renderutils/__init__.py:10-11; the reference's tests/test_bsdf.py drives them).  Not rebuilt: `diffuse_cubemap` / `specular_cubemap`
-- the split-sum cube-map light of nvdiffrec, which G-Shell's Monte-Carlo `EnvironmentLight` (render/light.py) never calls and for
which the reference holds no PyTorch statement to pin against (`use_python` asserts False, ops.py:412,458)."""
from .ops import (xfm_points, xfm_vectors, image_loss, prepare_shading_normal, lambert, frostbite_diffuse, pbr_specular, pbr_bsdf,
                  _fresnel_shlick, _ndf_ggx, _lambda_ggx, _masking_smith)

__all__ = ["xfm_vectors", "xfm_points", "image_loss", "prepare_shading_normal", "lambert", "frostbite_diffuse", "pbr_specular", "pbr_bsdf",
           "_fresnel_shlick", "_ndf_ggx", "_lambda_ggx", "_masking_smith"]
