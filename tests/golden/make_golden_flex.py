"""Generate (a) gshell_b200/geometry/flexicubes_tables.npz -- the Dual-Marching-Cubes / FlexiCubes look-up tables as a
binary data file (values: reference geometry/flexicubes_table.py) -- and (b) tests/golden/flex_*.npz: outputs and
gradients of the UNMODIFIED reference `GShellFlexiCubes.__call__` (geometry/gshell_flexicubes.py:136) on CPU.
Run in the build container only:   python tests/golden/make_golden_flex.py"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
from _ref_shim import reference_on_cpu   # noqa: E402

CASES = [("r4", 4, 0, "rand", False), ("r6", 6, 1, "rand", True), ("r8_sphere", 8, 2, "sphere", True),
         ("r10", 10, 3, "rand", True), ("r5_empty", 5, 4, "empty", False), ("r6_allopen", 6, 5, "closed_msdf", True)]


def main():
    warnings.filterwarnings("ignore")
    with reference_on_cpu() as imp:
        tab = imp("geometry.flexicubes_table")
        np.savez_compressed(os.path.join(ROOT, "gshell_b200", "geometry", "flexicubes_tables.npz"),
                            dmc_table=np.array(tab.dmc_table, dtype=np.int8), num_vd_table=np.array(tab.num_vd_table, dtype=np.int8),
                            check_table=np.array(tab.check_table, dtype=np.int16),
                            gflex_num_triangles_table=np.array(tab.gflex_num_triangles_table, dtype=np.int8),
                            gflex_configuration_table=np.array(tab.gflex_configuration_table, dtype=np.int8))
        fc = imp("geometry.gshell_flexicubes").GShellFlexiCubes(device="cpu")
        for name, res, seed, kind, weights in CASES:
            g = torch.Generator().manual_seed(seed)
            verts, cubes = fc.construct_voxel_grid(res)
            nv, nc = verts.shape[0], cubes.shape[0]
            x = (verts + 0.2 / res * (torch.rand(nv, 3, generator=g) - 0.5)).float()
            if kind == "rand":
                s = torch.rand(nv, generator=g) - 0.45
                nu = torch.rand(nv, generator=g) - 0.3
            elif kind == "sphere":
                s = verts.norm(dim=1) - 0.33 + 0.01 * torch.rand(nv, generator=g)
                nu = verts[:, 2] + 0.1 + 0.02 * torch.rand(nv, generator=g)
            elif kind == "empty":
                s = torch.rand(nv, generator=g) + 0.1
                nu = torch.rand(nv, generator=g) - 0.5
            else:
                s = torch.rand(nv, generator=g) - 0.45
                nu = -torch.rand(nv, generator=g) - 0.1          # no face fully inside: early return :566
            w = torch.randn(nc, 21, generator=g) * 0.5 if weights else None
            leaves = [t.clone().requires_grad_() for t in (x, s, nu)] + ([w.clone().requires_grad_()] if weights else [])
            out = {"x": x, "s": s, "nu": nu, "cubes": cubes, "res": np.array(res)}
            if weights:
                out["w"] = w
                wl = leaves[3]
                r = fc(leaves[0], leaves[1], leaves[2], cubes, res, wl[:, :12], wl[:, 12:20], wl[:, 20])
            else:
                r = fc(leaves[0], leaves[1], leaves[2], cubes, res)
            if kind == "empty":
                out["n_out"] = np.array(r[0].shape[0])
                np.savez_compressed(os.path.join(HERE, f"flex_{name}.npz"), **{k: np.asarray(v) for k, v in out.items()})
                print(name, "empty ->", r[0].shape, r[1].shape)
                continue
            vo, fo, L, ex = r
            gw = torch.Generator().manual_seed(100 + seed)
            wv, wm, wl_, ww = (torch.randn(t.shape, generator=gw) for t in (vo, ex["msdf"], L, ex["vertices_watertight"]))
            probe = (vo * wv).sum() + (ex["msdf"] * wm).sum() + (L * wl_).sum() + (ex["vertices_watertight"] * ww).sum()
            grads = torch.autograd.grad(probe, leaves, allow_unused=True)
            out.update(vertices_open=vo, faces_open=fo, L_dev=L, n_verts_watertight=np.array(ex["n_verts_watertight"]),
                       vertices_watertight=ex["vertices_watertight"], faces_watertight=ex["faces_watertight"], msdf=ex["msdf"],
                       msdf_watertight=ex["msdf_watertight"], msdf_boundary=ex["msdf_boundary"], wv=wv, wm=wm, wl=wl_, ww=ww)
            for nm, gg, lf in zip(("x", "s", "nu", "w"), grads, leaves):
                out["g_" + nm] = torch.zeros_like(lf) if gg is None else gg
            np.savez_compressed(os.path.join(HERE, f"flex_{name}.npz"),
                                **{k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in out.items()})
            print(name, "Vw", int(ex["n_verts_watertight"]), "Vo", vo.shape[0], "Fo", fo.shape[0], "Fw", ex["faces_watertight"].shape[0],
                  os.path.getsize(os.path.join(HERE, f"flex_{name}.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
