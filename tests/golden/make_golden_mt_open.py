"""Generate tests/golden/mtopen_*.npz: `GShell_Tets.__call__(..., output_watertight_template=False)` of the UNMODIFIED reference
(geometry/gshell_tets.py:260-263: tetrahedra whose four mSDF values are all non-positive are dropped BEFORE the edge de-duplication,
which changes the vertex numbering, the row counts and the keys of `extra`).  Run in the build container only:
    python tests/golden/make_golden_mt_open.py"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from _ref_shim import reference_on_cpu   # noqa: E402
from make_golden_mt import make_inputs    # noqa: E402

CASES = [("n4_rand", 4, 21, "rand"), ("n6_rand", 6, 22, "rand"), ("n5_zeros", 5, 23, "zeros"), ("n6_sphere", 6, 24, "sphere")]


def main():
    warnings.filterwarnings("ignore")
    with reference_on_cpu() as imp:
        ref = imp("geometry.gshell_tets").GShell_Tets()
        for name, n, seed, kind in CASES:
            pos, sdf, msdf, tets = make_inputs(n, seed, kind)
            if kind == "rand":        # make whole regions mSDF-negative so that tets really drop out
                msdf = torch.where(pos[:, 0] > 0.1, -msdf.abs() - 0.01, msdf)
            leaves = [x.clone().requires_grad_() for x in (pos, sdf, msdf)]
            va, fa, _, _, tng, extra = ref(*leaves, tets, output_watertight_template=False)
            assert set(extra) == {"msdf", "msdf_watertight", "msdf_boundary"}
            g = torch.Generator().manual_seed(2000 + seed)
            wa, wm = torch.randn(va.shape, generator=g), torch.randn(extra["msdf"].shape, generator=g)
            gm = torch.autograd.grad((va * wa).sum() + (extra["msdf"] * wm).sum(), leaves, allow_unused=True)
            out = {"pos": pos, "sdf": sdf, "msdf": msdf, "tets": tets, "verts_aug": va, "faces_aug": fa, "wa": wa, "wm": wm,
                   "msdf_aug": extra["msdf"], "msdf_watertight": extra["msdf_watertight"], "msdf_boundary": extra["msdf_boundary"]}
            for nm, a, leaf in zip(("pos", "sdf", "msdf"), gm, leaves):
                out[f"g_{nm}"] = torch.zeros_like(leaf) if a is None else a
            path = os.path.join(HERE, f"mtopen_{name}.npz")
            np.savez_compressed(path, **{k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in out.items()})
            full = ref(pos, sdf, msdf, tets)
            print(name, "Va", va.shape[0], "Fa", fa.shape[0], "Vw", extra["msdf_watertight"].shape[0], "| template=True: Va", full[0].shape[0],
                  "Vw", full[5]["n_verts_watertight"], os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
