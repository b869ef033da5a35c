"""Monte-Carlo sampled environment light with PDF / CDF tables (reference render/light.py:21-105)."""
import math
import os

import torch

from .. import _lib


class EnvironmentLight:
    """Lat-long environment probe `base` [h, w, 3] with the importance-sampling tables of the integrator (`_pdf`, `cols`, `rows`);
    attribute and method names as the reference's class (render/light.py:21-64)."""
    LIGHT_MIN_RES, MIN_ROUGHNESS, MAX_ROUGHNESS = 16, 0.08, 0.5

    def __init__(self, base):
        self.base, self.mtx = base, None
        h, w = base.shape[0], base.shape[1]
        self.pdf_scale = (h * w) / (2 * math.pi * math.pi)      # texels per steradian-ish unit of the lat-long map (reference :31)
        self.update_pdf()

    def xfm(self, mtx):
        self.mtx = mtx

    def parameters(self):
        return [self.base]

    def clone(self):
        return type(self)(self.base.detach().clone())

    def clamp_(self, min=None, max=None):
        self.base.clamp_(min=min, max=max)

    @torch.no_grad()
    def generate_image(self, res):
        """Reference light.py:61-64: the probe resampled to res = [h, w] (bilinear, wrapping) -- validation images only."""
        from . import util
        return util.texture_linear_wrap(self.base.detach(), util.pixel_grid(res[1], res[0], device=self.base.device))

    @torch.no_grad()
    def update_pdf(self):
        """Probe tables for importance sampling (reference light.py:46-59): two kernels on the h x w texels
        (csrc/tick_ops.cu::gsb_light_pdf) -> `_pdf`, `cols` (per-row column CDF), `rows` (row CDF, [h,w] like the reference)."""
        base = self.base.detach()
        _lib.require_cuda(base, "EnvironmentLight")
        base = base.float().contiguous()
        h, w = base.shape[0], base.shape[1]
        ws = torch.empty(h, dtype=torch.float32, device=base.device)
        self._pdf, self.cols, self.rows = (torch.empty((h, w), dtype=torch.float32, device=base.device) for _ in range(3))
        _lib.check(_lib.lib.gsb_light_pdf(_lib.ptr(base), h, w, _lib.ptr(ws), _lib.ptr(self._pdf), _lib.ptr(self.cols), _lib.ptr(self.rows),
                                          _lib.current_stream(base.device)), "gsb_light_pdf")


def create_trainable_env_rnd(base_res, scale=0.5, bias=0.25, device="cuda"):
    base = torch.rand(base_res, base_res, 3, dtype=torch.float32, device=device) * scale + bias
    return EnvironmentLight(base.clone().detach().requires_grad_(True))


@torch.no_grad()
def load_env(fn, scale=1.0, res=None, trainable=False, device="cuda"):
    """Reference light.py:70-91: a lat-long .hdr probe, optionally resampled to res = [h, w].  The file is read by the reference's
    `util.load_image` (image IO is not rebuilt here; see render/util.py)."""
    from . import util
    if os.path.splitext(fn)[1].lower() != ".hdr":
        raise ValueError("Unknown envlight extension %s" % os.path.splitext(fn)[1])
    img = torch.tensor(util.load_image(fn), dtype=torch.float32, device=device) * scale
    if res is not None:
        img = torch.clamp(util.texture_linear_wrap(img, util.pixel_grid(res[1], res[0], device=device)), min=0.0001)
    return EnvironmentLight(img.clone().detach().requires_grad_(True) if trainable else img)


@torch.no_grad()
def save_env_map(fn, light):
    """Reference light.py:93-97 (written by the reference's `util.save_image_raw`)."""
    from . import util
    assert isinstance(light, EnvironmentLight)
    util.save_image_raw(fn, light.generate_image([512, 1024]).detach().cpu().numpy())
