"""Generate tests/golden/mt_*.npz by running the UNMODIFIED reference `GShell_Tets.__call__`
(/root/reference/geometry/gshell_tets.py:245) on CPU through `_ref_shim.py`.

Run in the build container only:   python tests/golden/make_golden_mt.py
Each fixture stores the seeded inputs, every output of the call, and the gradients of two fixed
scalar probes with respect to (pos, sdf, msdf):
    probe_main = <verts_aug, Wa> + <extra.msdf, Wm> + <extra.vertices_watertight, Ww>
    probe_tng  = <v_tng_aug, Wt>
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from _ref_shim import reference_on_cpu          # noqa: E402
from gshell_b200.grids import bcc_tet_grid      # noqa: E402

CASES = [  # name, N, seed, sdf kind
    ("n3_rand", 3, 0, "rand"),
    ("n4_rand", 4, 1, "rand"),
    ("n6_rand", 6, 2, "rand"),
    ("n10_rand", 10, 3, "rand"),
    ("n8_sphere", 8, 4, "sphere"),
    ("n4_empty", 4, 5, "empty"),
    ("n5_zeros", 5, 6, "zeros"),      # exact-zero SDF / mSDF values exercise the sign() guards
]


def make_inputs(n, seed, kind):
    v, t = bcc_tet_grid(n)
    g = torch.Generator().manual_seed(seed)
    pos = torch.tensor(v) - 0.5 + 0.02 * (torch.rand(v.shape, generator=g) - 0.5)
    nv = v.shape[0]
    if kind == "rand":
        sdf = torch.rand(nv, generator=g) - 0.1
        msdf = (torch.rand(nv, generator=g) - 0.01).clamp(-1, 1)
        msdf = torch.where(torch.rand(nv, generator=g) < 0.5, msdf, -msdf)   # exercise all cut cases
    elif kind == "sphere":
        sdf = pos.norm(dim=1) - 0.35 + 0.01 * (torch.rand(nv, generator=g) - 0.5)
        msdf = pos[:, 1] + 0.05 + 0.02 * (torch.rand(nv, generator=g) - 0.5)
    elif kind == "empty":
        sdf = torch.rand(nv, generator=g) + 0.5
        msdf = torch.rand(nv, generator=g) - 0.5
    elif kind == "zeros":
        sdf = torch.rand(nv, generator=g) - 0.4
        sdf[torch.rand(nv, generator=g) < 0.2] = 0.0
        msdf = torch.rand(nv, generator=g) - 0.5
        msdf[torch.rand(nv, generator=g) < 0.3] = 0.0
    return pos.float(), sdf.float(), msdf.float(), torch.tensor(t)


def main():
    warnings.filterwarnings("ignore")
    with reference_on_cpu() as imp:
        ref = imp("geometry.gshell_tets").GShell_Tets()
        for name, n, seed, kind in CASES:
            pos, sdf, msdf, tets = make_inputs(n, seed, kind)
            leaves = [x.clone().requires_grad_() for x in (pos, sdf, msdf)]
            out = {"pos": pos, "sdf": sdf, "msdf": msdf, "tets": tets}
            if kind == "empty":
                # the reference fails on an empty surface inside map_uv/gather on some torch builds;
                # record what it does.
                try:
                    va, fa, _, _, tng, extra = ref(*leaves, tets)
                except Exception as e:          # pragma: no cover
                    print(name, "reference raised:", type(e).__name__, e)
                    out["reference_raised"] = np.array(1)
                    np.savez_compressed(os.path.join(HERE, f"mt_{name}.npz"),
                                        **{k: np.asarray(v) for k, v in out.items()})
                    continue
            else:
                va, fa, _, _, tng, extra = ref(*leaves, tets)
            g = torch.Generator().manual_seed(1000 + seed)
            wa = torch.randn(va.shape, generator=g)
            wm = torch.randn(extra["msdf"].shape, generator=g)
            ww = torch.randn(extra["vertices_watertight"].shape, generator=g)
            wt = torch.randn(tng.shape, generator=g)
            probe_main = (va * wa).sum() + (extra["msdf"] * wm).sum() + (extra["vertices_watertight"] * ww).sum()
            gm = torch.autograd.grad(probe_main, leaves, retain_graph=True, allow_unused=True)
            gt = torch.autograd.grad((tng * wt).sum(), leaves, allow_unused=True)
            out.update(verts_aug=va, faces_aug=fa, v_tng_aug=tng, wa=wa, wm=wm, ww=ww, wt=wt,
                       n_verts_watertight=np.array(extra["n_verts_watertight"]),
                       vertices_watertight=extra["vertices_watertight"],
                       faces_watertight=extra["faces_watertight"],
                       v_tng_watertight=extra["v_tng_watertight"],
                       msdf_aug=extra["msdf"], msdf_watertight=extra["msdf_watertight"],
                       msdf_boundary=extra["msdf_boundary"])
            for nm, a, b in zip(("pos", "sdf", "msdf"), gm, gt):
                out[f"gmain_{nm}"] = torch.zeros_like(leaves[0 if nm == 'pos' else 1]) if a is None else a
                out[f"gtng_{nm}"] = torch.zeros_like(leaves[0 if nm == 'pos' else 1]) if b is None else b
            arrs = {k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in out.items()}
            path = os.path.join(HERE, f"mt_{name}.npz")
            np.savez_compressed(path, **arrs)
            print(name, "Vw", int(extra["n_verts_watertight"]), "Va", va.shape[0], "Fa", fa.shape[0],
                  "Fw", extra["faces_watertight"].shape[0], os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
