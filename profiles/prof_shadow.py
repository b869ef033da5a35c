"""Shadow-ray throughput probe: G-buffer points sampled on an extracted G-Shell mesh, env_shade fwd with / without the
occluder.  usage: python profiles/prof_shadow.py [grid_N] [n_samples] [res]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gshell_b200.geometry.gshell_tets import GShell_Tets   # noqa: E402
from gshell_b200.grids import bcc_tet_grid                   # noqa: E402
from gshell_b200.render import light, optixutils as ou      # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 52
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
R = int(sys.argv[3]) if len(sys.argv) > 3 else 512
kind = sys.argv[4] if len(sys.argv) > 4 else "rand"
dev = torch.device("cuda:0")
v, t = bcc_tet_grid(N)
g = torch.Generator().manual_seed(0)
pos3 = ((torch.tensor(v) - 0.5) * 2).to(dev)
if kind == "rand":
    sdf = (torch.rand(v.shape[0], generator=g) - 0.1).to(dev)
else:
    sdf = (pos3.norm(dim=1) - 0.6)
msdf = (torch.rand(v.shape[0], generator=g) - 0.01).clamp(-1, 1).to(dev)
va, fa, _, _, _, _ = GShell_Tets(index_dtype=torch.int32)(pos3, sdf, msdf, torch.tensor(t).to(dev))
print("mesh", va.shape[0], fa.shape[0])
# render-like G-buffer: rasterise the mesh from one camera so that points are the VISIBLE surface
from gshell_b200 import synthetic
from gshell_b200.render import raster, renderutils as ru, mesh as meshmod
import numpy as np
mvp, campos = synthetic.random_cameras(1, (R, R), dev, np.random.RandomState(0))
clip = ru.xfm_points(va[None], mvp)
rast, _ = raster.rasterize(clip, fa, (R, R))
pos, _ = raster.interpolate(va[None], rast, fa)
nrm_v = meshmod.vertex_normals(va, fa)
nrm, _ = raster.interpolate(nrm_v[None], rast, fa)
view = campos.view(1, 1, 1, 3)
flip = ((view - pos) * nrm).sum(-1, keepdim=True) < 0
nrm = torch.nn.functional.normalize(torch.where(flip, -nrm, nrm), dim=-1)
mask = (rast[..., 3] > 0).float()
B, H, W = mask.shape
kd = torch.rand(B, H, W, 3, device=dev)
ks = torch.stack([torch.zeros(B, H, W), 0.3 + 0.6 * torch.rand(B, H, W), torch.rand(B, H, W)], -1).to(dev)
lgt = light.create_trainable_env_rnd(256, device=dev)
cov = int(mask.sum())
import ctypes
from gshell_b200 import _lib
from gshell_b200.render.optixutils import ops as _ops
print("lib", os.path.basename(_lib.LIB_PATH))
with torch.no_grad():
    for it in range(2):
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
        d, s = ou.optix_env_shade(None, mask, pos + nrm * 0.001, pos, nrm, view, kd, ks, lgt.base, lgt._pdf, lgt.rows[:, 0], lgt.cols, BSDF="pbr", n_samples_x=n, rnd_seed=it, shadow_scale=0.0)
        e1.record(); torch.cuda.synchronize()
print("no-shadow ms", e0.elapsed_time(e1), "covered", cov)
# cells per face: the occluder resolution knob (R = cbrt(cpf * F)); several values in one process
for cpf in [float(x) for x in os.environ.get("GSB_CPF_LIST", str(_ops.OCCLUDER_CELLS_PER_FACE)).split(",")]:
    _ops.OCCLUDER_CELLS_PER_FACE = cpf
    ctx = ou.OptiXContext()
    torch.cuda.synchronize(); t0 = time.time(); ou.optix_build_bvh(ctx, va, fa, 1); torch.cuda.synchronize()
    t0 = time.time(); ou.optix_build_bvh(ctx, va, fa, 1); torch.cuda.synchronize(); tb = time.time() - t0
    with torch.no_grad():
        for it in range(3):
            if it == 2:
                _lib.lib.gsb_trace_timing(1)
            torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
            d, s = ou.optix_env_shade(ctx, mask, pos + nrm * 0.001, pos, nrm, view, kd, ks, lgt.base, lgt._pdf, lgt.rows[:, 0], lgt.cols, BSDF="pbr", n_samples_x=n, rnd_seed=it, shadow_scale=1.0)
            e1.record(); torch.cuda.synchronize()
    tr = float(_lib.lib.gsb_trace_timing(0))
    ms = e0.elapsed_time(e1)
    print(f"shadow cpf {cpf} R {ctx.grid_res} entries/tri {ctx.n_entries / fa.shape[0]:.2f} build_ms {tb * 1e3:.1f} env_shade_ms {ms:.2f} trace_ms {tr:.2f} "
          f"mean_diff {float(d.mean()):.6f}")
    st = (ctypes.c_uint64 * 16)()
    _lib.lib.gsb_trace_stats(st, 1)
    rays = int(_lib.lib.gsb_trace_ray_count(1))
    if st[0]:
        print(f"  per ray ({rays} rays): tri tests {st[0] / rays:.1f} cell steps {st[1] / rays:.1f} sub-voxel steps {st[4] / rays:.1f} "
              f"cells descended {st[2] / rays:.2f} cells tested {st[5] / rays:.2f} hit fraction {st[3] / rays:.3f}")
        if st[8] + st[10] + st[12]:
            names = ["search", "desc", "test", "refill"]
            print("  pool blocks per ray / lanes: " + "  ".join(f"{names[i]} {st[8 + 2 * i] / rays:.3f} / {st[9 + 2 * i] / max(st[8 + 2 * i], 1):.1f}" for i in range(4)))
    else:
        print(f"  rays/launch-set {rays / 3:.0f}  G rays/s (trace only) {rays / 3 / max(tr, 1e-6) / 1e6:.2f}")
