"""Generate tests/golden/glue_*.npz: the pure-PyTorch loss-assembly pieces of tick() evaluated by the
UNMODIFIED reference on CPU (through _ref_shim): render/regularizer.py (chroma_loss, shading_loss, material_smoothness_grad)
and geometry/gshell_tets_geometry.py:33-39 (compute_sdf_reg_loss; the module itself cannot be imported without OptiX, so the
function is compiled from its source lines at generation time -- nothing is copied into this repository).
Run in the build container only:   python tests/golden/make_golden_glue.py"""
import ast
import os
import sys
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _ref_shim import REFERENCE_ROOT, reference_on_cpu   # noqa: E402


def function_from_reference(rel_path, name):
    src = open(os.path.join(REFERENCE_ROOT, rel_path)).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == name)
    ns = {"torch": torch}
    exec(compile(ast.Module(body=[node], type_ignores=[]), rel_path, "exec"), ns)
    return ns[name]


def main():
    warnings.filterwarnings("ignore")
    g = torch.Generator().manual_seed(11)
    R = lambda *s: torch.rand(*s, generator=g)          # noqa: E731
    out = {}
    # render/regularizer.py -> mesh -> obj -> material -> mlptexture imports tiny-cuda-nn at module scope (unused here)
    sys.modules.setdefault("tinycudann", types.ModuleType("tinycudann"))
    with reference_on_cpu() as imp:
        reg = imp("render.regularizer")
        B, H, W = 2, 12, 10
        diff, spec, kd = (R(B, H, W, 3).requires_grad_() for _ in range(3))
        alpha = (R(B, H, W, 1) > 0.3).float()
        color_ref = torch.cat([R(B, H, W, 3), alpha], -1)
        kd_grad = torch.cat([R(B, H, W, 3), alpha], -1).requires_grad_()
        ks_grad = torch.cat([R(B, H, W, 3) * torch.tensor([0.0, 1.0, 1.0]), alpha], -1).requires_grad_()
        nrm_grad = torch.cat([R(B, H, W, 3), alpha], -1).requires_grad_()
        l_sh = reg.shading_loss(diff, spec, color_ref, 0.15, 0.0025)
        l_ch = reg.chroma_loss(kd, color_ref, 0.3)
        l_ms = reg.material_smoothness_grad(kd_grad, ks_grad, nrm_grad, lambda_kd=0.25, lambda_ks=0.1, lambda_nrm=0.05)
        gs = torch.autograd.grad(l_sh + 2.0 * l_ch + 3.0 * l_ms, [diff, spec, kd, kd_grad, ks_grad, nrm_grad])
        out.update(diff=diff, spec=spec, kd=kd, color_ref=color_ref, kd_grad=kd_grad, ks_grad=ks_grad, nrm_grad=nrm_grad,
                   shading_loss=l_sh, chroma_loss=l_ch, material_smoothness=l_ms,
                   **{f"g_{k}": v for k, v in zip(("diff", "spec", "kd", "kd_grad", "ks_grad", "nrm_grad"), gs)})
    sdf_reg = function_from_reference("geometry/gshell_tets_geometry.py", "compute_sdf_reg_loss")
    sdf = (R(500) - 0.35).requires_grad_()
    sdf.data[::17] = 0.0                                   # torch.sign(0) = 0 differs from both signs: the reference's edge case
    edges = torch.randint(0, 500, (1800, 2), generator=g)
    l_sdf = sdf_reg(sdf, edges)
    out.update(sdf=sdf, edges=edges, sdf_reg=l_sdf, g_sdf=torch.autograd.grad(l_sdf, sdf)[0])
    path = os.path.join(HERE, "glue_losses.npz")
    np.savez_compressed(path, **{k: v.detach().numpy() for k, v in out.items()})
    print("glue_losses", os.path.getsize(path) // 1024, "KiB", {k: float(out[k]) for k in ("shading_loss", "chroma_loss", "material_smoothness", "sdf_reg")})


if __name__ == "__main__":
    main()
