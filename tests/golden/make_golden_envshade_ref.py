"""Generate tests/golden/shade_envshade_ref.npz: outputs AND gradients of the reference's own env-light integrator
(render/optixutils/c_src/envsampling/kernel.cu, unmodified, compiled for the CPU by oracle/build_ref.py) on seeded inputs, with
and without occluders.  Run in the build container only:   python tests/golden/make_golden_envshade_ref.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_env_shade as ref          # noqa: E402
from oracle import shade_oracle as so            # noqa: E402


def unit(x):
    return x / x.norm(dim=-1, keepdim=True)


def scene(seed, B, H, W, rough_min):
    g = torch.Generator().manual_seed(seed)
    R = lambda *s: torch.rand(*s, generator=g)          # noqa: E731
    pos = R(B, H, W, 3) - 0.5
    view = torch.tensor([0.2, 0.1, 3.0]).view(1, 1, 1, 3)
    nrm = unit(torch.randn(B, H, W, 3, generator=g) * 0.4 + unit(view - pos))
    kd = R(B, H, W, 3)
    ks = torch.stack([torch.zeros(B, H, W), rough_min + (0.98 - rough_min) * R(B, H, W), R(B, H, W)], -1)
    mask = (R(B, H, W) > 0.15).float()
    light = R(16, 32, 3) * 0.5 + 0.1
    light[3:5, 7:12] = 12.0                                # a sun: exercises the CDF importance sampling
    # occluders: a handful of large triangles around the points
    verts = (R(24, 3) - 0.5) * 3.0
    tris = torch.arange(24, dtype=torch.int32).view(8, 3)
    return dict(mask=mask, pos=pos, nrm=nrm, view=view, kd=kd, ks=ks, light=light, verts=verts, tris=tris)


def main():
    out = {}
    cases = [("pbr_n4", 0, 4, 11, 0.3, 0.0), ("pbr_rough008_n3", 0, 3, 12, 0.08, 0.0), ("diffuse_n4", 1, 4, 13, 0.08, 0.0),
             ("white_n2", 2, 2, 14, 0.08, 0.0), ("pbr_shadow_n3", 0, 3, 15, 0.3, 1.0), ("pbr_halfshadow_n2", 0, 2, 16, 0.3, 0.5)]
    for name, bsdf, n, seed, rough_min, shadow in cases:
        s = scene(seed, 2, 10, 9, rough_min)
        pdf, rows, cols = so.light_pdf_tables(s["light"])
        g = torch.Generator().manual_seed(100 + seed)
        perms = torch.argsort(torch.rand(97, n * n, generator=g), dim=-1).int()
        ro = s["pos"] + 0.001 * s["nrm"]
        occ = dict(verts=s["verts"], tris=s["tris"]) if shadow > 0 else {}
        args = (s["mask"], ro, s["pos"], s["nrm"], s["view"], s["kd"], s["ks"], s["light"], pdf, rows, cols, perms)
        kw = dict(bsdf=bsdf, n_samples_x=n, rnd_seed=1000 + seed, shadow_scale=shadow, **occ)
        diff, spec = ref.env_shade_fwd(*args, **kw)
        gd, gs = torch.rand(diff.shape, generator=g), torch.rand(spec.shape, generator=g)
        grads = ref.env_shade_bwd(*args, gd, gs, **kw)
        rec = dict(s, perms=perms, diff=diff, spec=spec, gd=gd, gs=gs, bsdf=torch.tensor(bsdf), n=torch.tensor(n), seed=torch.tensor(1000 + seed),
                   shadow=torch.tensor(shadow), **{f"g_{k}": v for k, v in zip(("pos", "nrm", "kd", "ks", "light"), grads)})
        out.update({f"{name}/{k}": v.numpy() for k, v in rec.items()})
        print(name, "diff mean", float(diff.mean()), "spec mean", float(spec.mean()))
    path = os.path.join(HERE, "shade_envshade_ref.npz")
    np.savez_compressed(path, **out)
    print(os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
