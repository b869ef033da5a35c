"""Extraction kernels for `ncu --set full`: one marching-tets forward + backward at the "256" grid (BCC N=103, random SDF/mSDF)
and one G-FlexiCubes forward + backward at 80^3.
usage (GPU box): ncu --set full --clock-control none --import-source on -k regex:'k_(tet|edge|occ|vertex|bwd|scan|case|raw|quads|dual|cut|boundary)' -c 40 -o gpurun_out/extract python profiles/prof_extract.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gshell_b200.geometry.gshell_flexicubes import GShellFlexiCubes   # noqa: E402
from gshell_b200.geometry.gshell_tets import GShell_Tets              # noqa: E402
from gshell_b200.grids import bcc_tet_grid                             # noqa: E402

dev = torch.device("cuda:0")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 103
v, t = bcc_tet_grid(N)
g = torch.Generator().manual_seed(0)
pos = (torch.tensor(v) - 0.5).to(dev).requires_grad_()
sdf = (torch.rand(v.shape[0], generator=g) - 0.1).to(dev).requires_grad_()
msdf = (torch.rand(v.shape[0], generator=g) - 0.01).clamp(-1, 1).to(dev).requires_grad_()
tets = torch.tensor(t).to(dev)
mt = GShell_Tets(index_dtype=torch.int32, with_tangents=False)
mt(pos, sdf, msdf, tets)                      # builds the static tables (not profiled: warm-up, filtered by -s if wanted)
torch.cuda.synchronize()
print("PROFILE_START mt")
va, fa, _, _, _, ex = mt(pos, sdf, msdf, tets)
(va.sum() + ex["msdf"].sum()).backward()
torch.cuda.synchronize()
print("mt", va.shape, fa.shape)
res = 80
fc = GShellFlexiCubes(device=dev, index_dtype=torch.int32)
verts, cubes = fc.construct_voxel_grid(res)
x = (verts + 0.2 / res * (torch.rand(verts.shape[0], 3, generator=g).to(dev) - 0.5)).requires_grad_()
s = (verts.norm(dim=1) - 0.35).requires_grad_()
nu = (verts[:, 1] + 0.15).requires_grad_()
w = (torch.randn(cubes.shape[0], 21, generator=g) * 0.5).to(dev).requires_grad_()
fc(x, s, nu, cubes, res, w[:, :12], w[:, 12:20], w[:, 20])
torch.cuda.synchronize()
print("PROFILE_START flex")
vo, ff, L, e2 = fc(x, s, nu, cubes, res, w[:, :12], w[:, 12:20], w[:, 20])
(vo.sum() + L.sum()).backward()
torch.cuda.synchronize()
print("flex", vo.shape, ff.shape)
