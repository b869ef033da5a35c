"""Vector / image / camera helpers used by the hot path (subset of the reference's render/util.py:19-35,
61-65,195-210,238-332; image IO and GLFW display are out of scope)."""
import importlib.util
import os
import sys

import numpy as np
import torch

_reference_module = None


def __getattr__(name):
    """Names this module does not define -- image IO (`save_image`, `load_image`, ...), the GLFW display, cube-map and text
    helpers: off the hot path, not rebuilt -- resolve to the reference's own render/util.py when `gshell_b200.dropin.install()` has
    put the reference's `render/` directory on this package's search path.  Loaded on first use (it imports nvdiffrast / imageio)."""
    global _reference_module
    if name.startswith("__"):
        raise AttributeError(name)
    if _reference_module is None:
        here = os.path.dirname(os.path.abspath(__file__))
        for d in list(getattr(sys.modules.get(__package__), "__path__", []))[1:]:
            cand = os.path.join(d, "util.py")
            if os.path.isfile(cand) and os.path.abspath(d) != here:
                spec = importlib.util.spec_from_file_location(__package__ + "._reference_util", cand)
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
                _reference_module = mod
                break
    if _reference_module is not None and hasattr(_reference_module, name):
        return getattr(_reference_module, name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r} (image IO / display helpers come from the reference's "
                         "render/util.py: run through gshell_b200.dropin)")



def dot(x, y):
    return torch.sum(x * y, -1, keepdim=True)


def length(x, eps=1e-20):
    return torch.sqrt(torch.clamp(dot(x, x), min=eps))


def safe_normalize(x, eps=1e-20):
    return x / length(x, eps)


def rgb_to_srgb(f):
    return torch.where(f <= 0.0031308, f * 12.92, torch.pow(torch.clamp(f, 0.0031308), 1.0 / 2.4) * 1.055 - 0.055)


def pixel_grid(width, height, center_x=0.5, center_y=0.5, device="cuda"):
    y, x = torch.meshgrid((torch.arange(0, height, dtype=torch.float32, device=device) + center_y) / height,
                          (torch.arange(0, width, dtype=torch.float32, device=device) + center_x) / width, indexing="ij")
    return torch.stack((x, y), dim=-1)


def scale_img_nhwc(x, size, mag="bilinear", min="area"):
    y = x.permute(0, 3, 1, 2)
    if x.shape[1] > size[0] and x.shape[2] > size[1]:
        y = torch.nn.functional.interpolate(y, size, mode=min)
    elif mag in ("bilinear", "bicubic"):
        y = torch.nn.functional.interpolate(y, size, mode=mag, align_corners=True)
    else:
        y = torch.nn.functional.interpolate(y, size, mode=mag)
    return y.permute(0, 2, 3, 1).contiguous()


def avg_pool_nhwc(x, size):
    return torch.nn.functional.avg_pool2d(x.permute(0, 3, 1, 2), size).permute(0, 2, 3, 1).contiguous()


def bilinear_tap(img, uv):
    """dr.texture(img, uv, filter_mode='linear', boundary_mode='clamp') for an NHWC image and [B,H,W,2] uv in [0,1]
    (texel centres at (i+0.5)/N).  Regulariser taps only (reference render.py:59,110): plain torch."""
    grid = uv * 2.0 - 1.0
    out = torch.nn.functional.grid_sample(img.permute(0, 3, 1, 2), grid, mode="bilinear", padding_mode="border",
                                          align_corners=False)
    return out.permute(0, 2, 3, 1)


def texture_linear_wrap(tex, uv):
    """dr.texture(tex, uv, filter_mode='linear') with nvdiffrast's default boundary mode 'wrap', for a [H,W,C] image and [h,w,2]
    coordinates in [0,1] (texel centres at (i + 0.5) / N) -> [h,w,C].  Plain torch: used off the hot path only
    (EnvironmentLight.generate_image, validation images)."""
    H, W = tex.shape[0], tex.shape[1]
    x, y = uv[..., 0] * W - 0.5, uv[..., 1] * H - 0.5
    x0, y0 = torch.floor(x), torch.floor(y)
    fx, fy = (x - x0).unsqueeze(-1), (y - y0).unsqueeze(-1)
    x0, y0 = x0.long(), y0.long()
    xa, xb, ya, yb = x0 % W, (x0 + 1) % W, y0 % H, (y0 + 1) % H
    top = tex[ya, xa] * (1 - fx) + tex[ya, xb] * fx
    bot = tex[yb, xa] * (1 - fx) + tex[yb, xb] * fx
    return top * (1 - fy) + bot * fy


def perspective(fovy=0.7854, aspect=1.0, n=0.1, f=1000.0, device=None):
    y = np.tan(fovy / 2)
    return torch.tensor([[1 / (y * aspect), 0, 0, 0], [0, 1 / -y, 0, 0], [0, 0, -(f + n) / (f - n), -(2 * f * n) / (f - n)],
                         [0, 0, -1, 0]], dtype=torch.float32, device=device)


def translate(x, y, z, device=None):
    return torch.tensor([[1, 0, 0, x], [0, 1, 0, y], [0, 0, 1, z], [0, 0, 0, 1]], dtype=torch.float32, device=device)


@torch.no_grad()
def random_rotation_translation(t, device=None, rng=np.random):
    m = rng.normal(size=[3, 3])
    m[1] = np.cross(m[0], m[2])
    m[2] = np.cross(m[0], m[1])
    m = m / np.linalg.norm(m, axis=1, keepdims=True)
    m = np.pad(m, [[0, 1], [0, 1]], mode="constant")
    m[3, 3] = 1.0
    m[:3, 3] = rng.uniform(-t, t, size=[3])
    return torch.tensor(m, dtype=torch.float32, device=device)
