"""Generate tests/golden/auggrid_*.npz: the UNMODIFIED reference `GShell_Tets.marching_from_auggrid`
(geometry/gshell_tets.py:446-629) on CPU (through _ref_shim) on seeded synthetic augmented grids.
Inputs mirror eval_gmeshdiffusion_generated_samples.py:118-170 and gshell_tets_geometry.py:70-78 on a BCC grid with N cells per
axis: vertices discretised with dx = half the spacing of the unique coordinate values (corners on multiples of 4, centres on
4k+2, so every edge midpoint is an integer point of the (4N+1)^3 lattice), an SDF sign per vertex, interpolation coefficients
and mSDF signs stored at those midpoints, occupancy values on the (8N+1)^3 lattice of polygon-edge midpoints.
Run in the build container only:   python tests/golden/make_golden_auggrid.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from _ref_shim import reference_on_cpu          # noqa: E402
from gshell_b200.grids import bcc_tet_grid       # noqa: E402

ENDS = [0, 1, 0, 2, 0, 3, 1, 2, 1, 3, 2, 3]


def inputs(n, seed, kind):
    v, t = bcc_tet_grid(n)
    g = torch.Generator().manual_seed(seed)
    verts = torch.tensor(v, dtype=torch.float32)
    tets = torch.tensor(t, dtype=torch.long)
    disc = torch.round(verts * (4 * n)).long().float()                        # reference passes `.long().float()` (:78)
    pos = (verts - 0.5) * 2.0 + 0.2 / n * (torch.rand(verts.shape, generator=g) - 0.5)   # "deformed" vertices
    if kind == "rand":
        sdf = torch.sign(torch.rand(verts.shape[0], generator=g) - 0.35)
    elif kind == "sphere":
        sdf = torch.sign((verts - 0.5).norm(dim=1) - 0.33)
    else:                                                                       # no surface at all
        sdf = torch.ones(verts.shape[0])
    sdf[sdf == 0] = 1.0
    ends = tets[:, ENDS].reshape(-1, 6, 2)
    sorted_edges = torch.sort(ends, dim=-1)[0]
    G = 4 * n + 1
    coeff = torch.rand(G, G, G, generator=g) * 1.4 - 0.2                        # exercises the clamp(0, 1)
    msdf_sign = torch.sign(torch.rand(G, G, G, generator=g) - 0.4)
    occ = torch.rand(8 * n + 1, 8 * n + 1, 8 * n + 1, generator=g) * 2 - 1
    return dict(pos=pos, sdf=sdf, tets=tets, sorted_edges=sorted_edges, coeff=coeff, disc=disc, msdf_sign=msdf_sign, occ=occ)


def main():
    with reference_on_cpu() as imp:
        mod = imp("geometry.gshell_tets")
        ref = mod.GShell_Tets()
        for name, n, seed, kind in (("n3_rand", 3, 1, "rand"), ("n4_rand", 4, 2, "rand"), ("n4_sphere", 4, 3, "sphere"),
                                    ("n3_empty", 3, 4, "empty")):
            a = inputs(n, seed, kind)
            try:
                out = ref.marching_from_auggrid(a["pos"], a["sdf"], a["tets"], a["sorted_edges"], a["coeff"], a["disc"],
                                                a["msdf_sign"], a["occ"])
            except Exception as e:                                              # record how the reference behaves on no surface
                print(name, "reference raised", type(e).__name__, e)
                continue
            va, fa, _, _, tng, v, gidx, m_aug, m = out
            rec = dict(a, verts_aug=va, faces_aug=fa, v_tng_aug=tng, verts=v, valid_tet_gidx=gidx, msdf_aug=m_aug, msdf=m)
            path = os.path.join(HERE, f"auggrid_{name}.npz")
            np.savez_compressed(path, **{k: x.numpy() for k, x in rec.items()})
            print(name, "Vw", v.shape[0], "Va", va.shape[0], "Fa", fa.shape[0], os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
