"""Generate tests/golden/shade_*.npz from the reference's own PyTorch code on CPU (through _ref_shim):
  render/renderutils/ops.py (use_python=True paths), bsdf.py, loss.py, render/light.py and the Python
  BilateralDenoiser embedded in render/optixutils/tests/filter_test.py:31-74.
Run in the build container only:   python tests/golden/make_golden_shade.py"""
import math
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _ref_shim import REFERENCE_ROOT, reference_on_cpu   # noqa: E402


def unit(x):
    return x / x.norm(dim=-1, keepdim=True)


def save(name, d):
    path = os.path.join(HERE, f"shade_{name}.npz")
    np.savez_compressed(path, **{k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()})
    print(name, os.path.getsize(path) // 1024, "KiB")


def grads(out, leaves, w):
    return torch.autograd.grad((out * w).sum(), leaves, allow_unused=True)


def main():
    warnings.filterwarnings("ignore")
    g = torch.Generator().manual_seed(7)
    R = lambda *s: torch.rand(*s, generator=g)          # noqa: E731
    N = lambda *s: torch.randn(*s, generator=g)         # noqa: E731
    with reference_on_cpu() as imp:
        ru = imp("render.renderutils")
        rb = imp("render.renderutils.bsdf")
        light_mod = imp("render.light")

        # ---- xfm_points ---------------------------------------------------------------------------
        pts = N(1, 257, 3).requires_grad_()
        mtx = N(3, 4, 4)
        out = ru.xfm_points(pts, mtx, use_python=True)
        w = N(*out.shape)
        save("xfm", {"points": pts, "matrix": mtx, "out": out, "w": w, "g_points": grads(out, [pts], w)[0]})

        # ---- prepare_shading_normal ---------------------------------------------------------------
        B, H, W = 2, 9, 13
        for tag, with_pert in (("nrm_plain", False), ("nrm_perturbed", True)):
            pos = N(B, H, W, 3).requires_grad_()
            view = (N(B, 1, 1, 3) * 3).requires_grad_()
            pert = (unit(N(B, H, W, 3)) * torch.tensor([0.3, 0.3, 1.0])).requires_grad_() if with_pert else None
            sn = N(B, H, W, 3).requires_grad_()
            st = N(B, H, W, 3).requires_grad_()
            gn = unit(N(B, H, W, 3)).requires_grad_()
            out = ru.prepare_shading_normal(pos, view, pert, sn, st, gn, two_sided_shading=True, opengl=True, use_python=True)
            w = N(*out.shape)
            leaves = [pos, view, sn, st, gn] + ([pert] if with_pert else [])
            gs = grads(out, leaves, w)
            d = {"pos": pos, "view_pos": view, "smooth_nrm": sn, "smooth_tng": st, "geom_nrm": gn, "out": out, "w": w}
            for nme, gg in zip(["pos", "view_pos", "smooth_nrm", "smooth_tng", "geom_nrm", "perturbed_nrm"], gs):
                d["g_" + nme] = gg
            if with_pert:
                d["perturbed_nrm"] = pert
            save(tag, d)

        # ---- image_loss -----------------------------------------------------------------------------
        img = (R(2, 8, 16, 3) * 2.0).requires_grad_()
        tgt = (R(2, 8, 16, 3) * 2.0).requires_grad_()
        d = {"img": img, "target": tgt}
        for loss, tm in (("l1", "none"), ("l1", "log_srgb"), ("mse", "log_srgb"), ("smape", "none"), ("relmse", "none"), ("mse", "none")):
            if tm == "log_srgb":
                # The reference's Python path multiplies by exposure=5 inside _tonemap_srgb (loss.py:16-18) but
                # its CUDA kernel -- the one training runs -- does not (loss.cu:38-41).  The CUDA kernel is the
                # contract, so the golden uses the reference's own helper with exposure=1.
                lm = imp("render.renderutils.loss")
                ti = lm._tonemap_srgb(torch.log(torch.clamp(img, min=0, max=65535) + 1), exposure=1)
                tt = lm._tonemap_srgb(torch.log(torch.clamp(tgt, min=0, max=65535) + 1), exposure=1)
                val = lm.image_loss_fn(ti, tt, loss, "none")
            else:
                val = ru.image_loss(img, tgt, loss=loss, tonemapper=tm, use_python=True)
            gi, gt = torch.autograd.grad(val, [img, tgt])
            d[f"{loss}_{tm}"] = val
            d[f"{loss}_{tm}_g_img"] = gi
            d[f"{loss}_{tm}_g_target"] = gt
        save("loss", d)

        # ---- BSDF pieces ----------------------------------------------------------------------------
        P = 4096
        nrm = unit(N(P, 3)).requires_grad_()
        wi = unit(nrm.detach() + 0.8 * N(P, 3)).requires_grad_()
        wo = unit(nrm.detach() + 0.8 * N(P, 3)).requires_grad_()
        col = R(P, 3).requires_grad_()
        alpha = (R(P, 1) ** 2).requires_grad_()
        spec = rb.bsdf_pbr_specular(col, nrm, wo, wi, alpha, min_roughness=0.08)
        lam = rb.bsdf_lambert(nrm, wi)
        ws = N(P, 3)
        gs = grads(spec, [col, nrm, wo, wi, alpha], ws)
        cosv = (R(P, 1) * 1.2 - 0.1)
        a2 = R(P, 1)
        save("bsdf", {"nrm": nrm, "wi": wi, "wo": wo, "col": col, "alpha": alpha, "spec": spec, "lambert": lam, "ws": ws,
                      "g_col": gs[0], "g_nrm": gs[1], "g_wo": gs[2], "g_wi": gs[3], "g_alpha": gs[4],
                      "cos": cosv, "a2": a2, "ndf": rb.bsdf_ndf_ggx(a2, cosv), "lambda": rb.bsdf_lambda_ggx(a2, cosv),
                      "fresnel": rb.bsdf_fresnel_shlick(col[:, 0:1].detach(), 1.0, cosv),
                      "masking": rb.bsdf_masking_smith_ggx_correlated(a2, cosv, torch.flip(cosv, [0]))})

        # ---- light pdf / cdf tables -----------------------------------------------------------------
        base = R(16, 32, 3) * 0.5 + 0.25
        base[3:5, 7:9] = 20.0
        lgt = light_mod.EnvironmentLight(base)
        save("light", {"base": base, "pdf": lgt._pdf, "rows": lgt.rows[:, 0], "cols": lgt.cols})

        # ---- bilateral denoiser (reference's Python filter, filter_test.py:31-74) -------------------
        src = open(os.path.join(REFERENCE_ROOT, "render/optixutils/tests/filter_test.py")).read()
        cls = src[src.index("class BilateralDenoiser"):src.index("def relative_loss")]
        ns = {"torch": torch, "np": np, "math": math, "dot": lambda a, b: torch.sum(a * b, -1, keepdim=True)}
        exec(cls, ns)
        for sigma, tag in ((0.6, "denoise_s06"), (1.0, "denoise_s10")):
            inp = R(2, 12, 20, 11)
            inp[..., 3:6] = unit(N(2, 12, 20, 3) * 0.2 + torch.tensor([0.0, 0.0, 1.0]))
            inp[..., 9] = 0.9 + 0.05 * R(2, 12, 20)          # z
            inp[..., 10] = 0.001 + 0.004 * R(2, 12, 20)      # dz
            inp = inp.requires_grad_()
            out = ns["BilateralDenoiser"](sigma=sigma).forward(inp)
            w = N(*out.shape)
            g_in = grads(out, [inp], w)[0]
            save(tag, {"col": inp[..., 0:3], "nrm": inp[..., 3:6], "zdz": inp[..., 9:11], "sigma": np.array(sigma),
                       "out": out, "w": w, "g_col": g_in[..., 0:3]})


if __name__ == "__main__":
    main()
