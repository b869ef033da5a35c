"""Shadow-ray traversal core (gshell_b200/csrc/trace_core.cuh) on the CPU: the three-level walk (bricks -> cells ->
sub-voxel bits) must find exactly the rays that brute-force Moeller-Trumbore finds, on the random-SDF "fog" the benchmark
uses, on a closed surface, and on an arbitrary triangle soup.  The host driver (tests/native/trace_host.cu) is compiled
here with nvcc as plain host code; the kernel's warp-level glue is covered by the -m gpu tests."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "native", "trace_host.cu")
OUT = os.path.join(HERE, "native", "_build", "libtrace_host.so")


@pytest.fixture(scope="module")
def lib():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    core = os.path.join(os.path.dirname(HERE), "gshell_b200", "csrc", "trace_core.cuh")
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(SRC), os.path.getmtime(core)):
        subprocess.run([nvcc, "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-shared", "-Xcompiler", "-fPIC", SRC, "-o", OUT],
                       check=True)
    L = ctypes.CDLL(OUT)
    L.trace_host.restype = ctypes.c_int
    L.trace_host.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p,
                             ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    return L


def trace(L, verts, tris, R, rays, sub=1):
    verts = np.ascontiguousarray(verts, np.float32); tris = np.ascontiguousarray(tris, np.int32)
    rays = np.ascontiguousarray(rays, np.float32)
    vis = np.zeros(rays.shape[0], np.uint8); stats = np.zeros(6, np.int64)
    L.trace_host(verts.ctypes.data, verts.shape[0], tris.ctypes.data, tris.shape[0], R, rays.ctypes.data, rays.shape[0], sub,
                 vis.ctypes.data, stats.ctypes.data)
    return vis, stats


def brute_force(verts, tris, rays, chunk=2048):
    """float32 Moeller-Trumbore against every triangle (same formulation as tests/test_shade_gpu.py's oracle stand-in)."""
    v0 = verts[tris[:, 0]]; e1 = verts[tris[:, 1]] - v0; e2 = verts[tris[:, 2]] - v0
    out = np.ones(rays.shape[0], np.uint8)
    margin = np.full(rays.shape[0], np.inf, np.float32)
    for i in range(0, rays.shape[0], chunk):
        o = rays[i:i + chunk, None, :3]; d = rays[i:i + chunk, None, 3:]
        p = np.cross(d, e2[None]); det = (e1[None] * p).sum(-1)
        ok = det != 0
        inv = 1.0 / np.where(ok, det, 1).astype(np.float32)
        t_ = o - v0[None]
        u = (t_ * p).sum(-1) * inv
        q = np.cross(t_, e1[None])
        v = (d * q).sum(-1) * inv
        t = (e2[None] * q).sum(-1) * inv
        hit = ok & (u >= 0) & (u <= 1) & (v >= 0) & (u + v <= 1) & (t > 0) & (t < 1e16)
        out[i:i + chunk] = ~hit.any(1)
        # distance of the closest call to an edge of any triangle in front of the ray (for diagnosing grazing mismatches)
        m = np.minimum(np.minimum(np.abs(u), np.abs(v)), np.abs(1 - u - v))
        m = np.where(ok & (t > 0), m, np.inf)
        margin[i:i + chunk] = m.min(1)
    return out, margin


def fog_mesh(N, seed, kind="rand"):
    from gshell_b200.grids import bcc_tet_grid
    from oracle import mt_oracle
    v, t = bcc_tet_grid(N)
    g = torch.Generator().manual_seed(seed)
    pos = (torch.tensor(v) - 0.5) * 2
    if kind == "rand":
        sdf = torch.rand(v.shape[0], generator=g) - 0.1
    else:
        sdf = pos.norm(dim=1) - 0.6 + 0.05 * (torch.rand(v.shape[0], generator=g) - 0.5)
    msdf = (torch.rand(v.shape[0], generator=g) - 0.01).clamp(-1, 1)
    out = mt_oracle.gshell_marching_tets(pos, sdf, msdf, torch.tensor(t), unique_mode="packed", with_tangents=False)
    return out[0].numpy().astype(np.float32), out[1].numpy().astype(np.int32)


def surface_rays(verts, tris, n, rs):
    """Shadow-ray-like rays: origins on the mesh, pushed 1e-3 along the face normal, directions uniform on the sphere."""
    sel = rs.randint(0, tris.shape[0], n)
    b = rs.rand(n, 3).astype(np.float32); b /= b.sum(-1, keepdims=True)
    tri = verts[tris[sel]]
    p = (tri * b[..., None]).sum(1)
    nrm = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    nrm /= np.maximum(np.linalg.norm(nrm, axis=-1, keepdims=True), 1e-20)
    d = rs.randn(n, 3).astype(np.float32); d /= np.linalg.norm(d, axis=-1, keepdims=True)
    return np.concatenate([p + 1e-3 * nrm, d], -1).astype(np.float32)


def check(L, verts, tris, R, rays, max_bad=2e-3):
    want, margin = brute_force(verts, tris, rays)
    for sub in (0, 1):
        got, stats = trace(L, verts, tris, R, rays, sub)
        bad = np.nonzero(got != want)[0]
        # the only legitimate disagreements are rays grazing a triangle edge (different fp32 rounding of the same test)
        assert len(bad) <= max_bad * len(rays), (sub, len(bad), len(rays))
        assert np.all(margin[bad] < 1e-3), (sub, margin[bad])
    return stats


@pytest.mark.parametrize("N,R,kind", [(6, 13, "rand"), (8, 17, "rand"), (8, 23, "sphere"), (5, 4, "rand"), (7, 31, "rand")])
def test_walk_matches_brute_force_on_extracted_meshes(lib, N, R, kind):
    verts, tris = fog_mesh(N, seed=N, kind=kind)
    assert tris.shape[0] > 50
    rs = np.random.RandomState(N + R)
    rays = surface_rays(verts, tris, 3000, rs)
    stats = check(lib, verts, tris, R, rays)
    assert stats[5] <= stats[4]          # cells tested <= cells descended into: the sub-voxel bits only ever cull


def test_walk_axis_parallel_outside_and_degenerate_rays(lib):
    verts, tris = fog_mesh(6, seed=3)
    rs = np.random.RandomState(0)
    n = 1500
    o = (rs.rand(n, 3).astype(np.float32) - 0.5) * 4.0            # many origins outside the grid
    d = rs.randn(n, 3).astype(np.float32)
    d[: n // 3, rs.randint(0, 3)] = 0.0                           # a zero component
    d[n // 3: n // 2] = np.eye(3, dtype=np.float32)[rs.randint(0, 3, n // 2 - n // 3)] * rs.choice([-1, 1], (n // 2 - n // 3, 1))
    d /= np.maximum(np.linalg.norm(d, axis=-1, keepdims=True), 1e-20)
    check(lib, verts, tris, 13, np.concatenate([o, d], -1))


def test_walk_random_soup_various_resolutions(lib):
    rs = np.random.RandomState(5)
    c = rs.rand(400, 1, 3).astype(np.float32) * 2 - 1
    verts = (c + 0.15 * rs.randn(400, 3, 3).astype(np.float32)).reshape(-1, 3)
    tris = np.arange(1200, dtype=np.int32).reshape(-1, 3)
    o = rs.rand(2000, 3).astype(np.float32) * 2 - 1
    d = rs.randn(2000, 3).astype(np.float32); d /= np.linalg.norm(d, axis=-1, keepdims=True)
    for R in (1, 3, 8, 21):
        check(lib, verts, tris, R, np.concatenate([o, d], -1))


def test_subvoxel_bits_cull_most_cells_on_the_fog(lib):
    """The design claim behind level 2: on the random-SDF soup about half of the occupied cells a ray
    crosses are rejected by their sub-voxel bits without fetching a triangle."""
    verts, tris = fog_mesh(10, seed=1)
    rs = np.random.RandomState(2)
    rays = surface_rays(verts, tris, 4000, rs)
    R = int(round((2.0 * tris.shape[0]) ** (1 / 3)))
    _, s0 = trace(lib, verts, tris, R, rays, 0)
    _, s1 = trace(lib, verts, tris, R, rays, 1)
    assert s1[1] < 0.75 * s0[1] and s1[5] < 0.7 * s1[4], (s0, s1)


def test_bounded_subvoxel_walk_resumes_to_the_same_result(lib):
    """GSB_TRACE_FINE_CAP: cutting the sub-voxel walk into pieces and resuming from the saved position changes nothing -- same
    visibility, same number of sub-voxel steps, same cells tested."""
    verts, tris = fog_mesh(10, 5)
    rs = np.random.RandomState(11)
    rays = surface_rays(verts, tris, 4000, rs)
    vis1, st1 = trace(lib, verts, tris, 24, rays, sub=1)
    for cap in (1, 2, 3):
        vis, st = trace(lib, verts, tris, 24, rays, sub=1 + cap)
        assert np.array_equal(vis, vis1)
        assert st[3] == st1[3] and st[5] == st1[5] and st[1] == st1[1]
