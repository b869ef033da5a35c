"""Work per shadow ray of the pooled trace kernel (csrc/occluder.cu, csrc/trace_core.cuh), counted by the kernel's own counters:
hardware-independent numbers behind DESIGN 4.1 -- the three-level hierarchy (bricks -> cells -> 64 sub-voxel bits per cell) must
keep rejecting most entered cells without fetching a triangle.  The counters exist only in builds with -DGSB_TRACE_STATS (global
atomics: profiling builds); on the host emulator that is one more -D of the host build (GSB_HOST_DEFINES=GSB_TRACE_STATS), so this
file runs there (tests/test_emulated_gpu_suite_cpu.py) and skips on a library without counters."""
import ctypes

import numpy as np
import pytest
import torch

from _device import DEVICE, device      # cuda:0, or the CPU under the host emulator (tests/_device.py)

pytestmark = pytest.mark.gpu


def _stats(reset):
    from gshell_b200 import _lib
    out = (ctypes.c_uint64 * 16)()
    _lib.lib.gsb_trace_stats(out, 1 if reset else 0)
    return np.array(list(out), dtype=np.float64)


@pytest.mark.parametrize("kind", ["soup", "sphere"])
def test_work_per_ray_of_the_bench_fields(kind):
    """The two fields of the benchmark on a small grid: the random SDF / mSDF ("soup": a third of the tets active, the headline
    configuration) and the reference's sphere_init shell."""
    import gshell_b200.render.optixutils as ou
    from gshell_b200.geometry.gshell_tets import GShell_Tets
    from gshell_b200.grids import bcc_tet_grid
    from oracle import shade_oracle as so
    d = device()
    v, t = bcc_tet_grid(12)
    g = torch.Generator().manual_seed(5)
    p = (torch.tensor(v) - 0.5) * 2
    if kind == "soup":
        sdf, msdf = torch.rand(v.shape[0], generator=g) - 0.1, (torch.rand(v.shape[0], generator=g) - 0.01).clamp(-1, 1)
    else:
        sdf, msdf = p.norm(dim=1) - 0.5, torch.ones(v.shape[0])
    va, fa, _, _, _, _ = GShell_Tets(index_dtype=torch.int32, with_tangents=False)(p.to(d), sdf.to(d), msdf.to(d), torch.tensor(t).to(d))
    va_c, fa_c = va.cpu(), fa.cpu().long()
    B, H, W, n = 1, 24, 24, 3
    sel = torch.randint(0, fa_c.shape[0], (B * H * W,), generator=g)
    bary = torch.rand(B * H * W, 3, generator=g)
    bary = bary / bary.sum(-1, keepdim=True)
    tri = va_c[fa_c[sel]]
    pos = (tri * bary[..., None]).sum(1).view(B, H, W, 3)
    fn = torch.nn.functional.normalize(torch.linalg.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]), dim=-1).view(B, H, W, 3)
    view = torch.tensor([0.0, 0.0, 3.0]).view(1, 1, 1, 3)
    nrm = torch.where(((view - pos) * fn).sum(-1, keepdim=True) > 0, fn, -fn)
    kd = torch.rand(B, H, W, 3, generator=g)
    ks = torch.stack([torch.zeros(B, H, W), 0.4 + 0.5 * torch.rand(B, H, W, generator=g), torch.rand(B, H, W, generator=g)], -1)
    light = torch.rand(16, 32, 3, generator=g) + 0.2
    pdf, rows, cols = so.light_pdf_tables(light)
    ctx = ou.OptiXContext()
    ou.optix_build_bvh(ctx, va, fa, rebuild=1)
    _stats(reset=True)
    from gshell_b200 import _lib
    rays0 = int(_lib.lib.gsb_trace_ray_count(1))
    ou.optix_env_shade(ctx, torch.ones(B, H, W, device=d), (pos + 0.001 * nrm).to(d), pos.to(d), nrm.to(d), view.to(d), kd.to(d), ks.to(d),
                       light.to(d), pdf.to(d), rows.to(d), cols.to(d), BSDF="pbr", n_samples_x=n, rnd_seed=7, shadow_scale=1.0)
    s = _stats(reset=True)
    rays = int(_lib.lib.gsb_trace_ray_count(1))
    if s.sum() == 0:
        pytest.skip("library built without -DGSB_TRACE_STATS")
    assert rays > 0.3 * 2 * n * n * B * H * W, (rays, rays0)
    tri_tests, cell_steps, descended, hits, fine_steps, cells_tested = (s[k] / rays for k in range(6))
    print(f"{kind}: rays {rays}, per ray: cell steps {cell_steps:.2f}, cells entered {descended:.2f}, cells tested {cells_tested:.2f}, "
          f"sub-voxel steps {fine_steps:.2f}, triangle tests {tri_tests:.2f}, hits {hits:.3f}; faces {fa.shape[0]}; "
          f"tests a per-triangle sub-voxel mask would keep: {s[6] / rays:.2f} (ray's whole path in the cell), {s[7] / rays:.2f} (first occupied sub-voxel only)")
    # every ray that hits stops at its first hit; a ray tests triangles only in cells where it touched an occupied sub-voxel
    assert hits <= 1.0 and cells_tested <= descended
    # the sub-voxel bits reject a large share of the entered cells without a triangle fetch (B200, "256" grid: 64 %)
    assert cells_tested <= 0.8 * descended, (cells_tested, descended)          # 0.66 / 0.59 on these two small scenes
    # lanes per executed block of the pooled scheduler (SEARCH / DESC / TEST): the pool keeps most of the 32 lanes busy
    for name, k in (("search", 8), ("descend", 10), ("test", 12)):
        if s[k] > 0:
            assert s[k + 1] / s[k] >= 12.0, (name, s[k + 1] / s[k])
