"""Synthetic grids in the reference's on-disk formats.

The reference loads `data/tets/{res}_tets.npz` (keys `vertices` float[Nv,3], `indices` int[T,4];
reference geometry/gshell_tets_geometry.py:58-67, written by data/tets/generate_tets.py:47).  Those
files are Google-Drive downloads and are not available offline, so benchmarks and tests use a
body-centred-cubic (BCC) tetrahedral lattice generated here: N cells per axis give
(N+1)^3 + N^3 vertices and 12*N^2*(N-1) tetrahedra.  Name -> N mapping used throughout the repo
(matched on vertex count to the reference's quartet grids):
"64" -> N=26, "128" -> N=52, "256" -> N=103.
"""
import numpy as np

GRID_NAME_TO_N = {64: 26, 128: 52, 256: 103}


def bcc_tet_grid(n_cells: int):
    """Return (vertices float32 [Nv,3] in [0,1]^3, indices int64 [T,4]) of a BCC tet lattice.

    Vertex ids: corners first, (i,j,k) row-major over (N+1)^3, then cell centres row-major over N^3.
    Every tet joins the centres of two face-adjacent cells with one edge of their shared face.
    All tets are positively oriented (det[v1-v0, v2-v0, v3-v0] > 0).
    """
    n = int(n_cells)
    assert n >= 2
    g = np.arange(n + 1, dtype=np.float32) / n
    corners = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    c = (np.arange(n, dtype=np.float32) + 0.5) / n
    centres = np.stack(np.meshgrid(c, c, c, indexing="ij"), -1).reshape(-1, 3)
    vertices = np.concatenate([corners, centres], 0).astype(np.float32)
    n_corner = (n + 1) ** 3

    def corner_id(i, j, k):
        return (i * (n + 1) + j) * (n + 1) + k

    def centre_id(i, j, k):
        return n_corner + (i * n + j) * n + k

    tets = []
    for axis in range(3):
        # cells (a, b, c): `a` along `axis` in [0, n-2]; the shared face sits at a+1
        a, b, cc = np.meshgrid(np.arange(n - 1), np.arange(n), np.arange(n), indexing="ij")
        a, b, cc = a.ravel(), b.ravel(), cc.ravel()

        def perm(x, y, z):
            # place the coordinate that runs along `axis` first
            out = [None, None, None]
            out[axis] = x
            out[(axis + 1) % 3] = y
            out[(axis + 2) % 3] = z
            return out

        m0 = centre_id(*perm(a, b, cc))
        m1 = centre_id(*perm(a + 1, b, cc))
        ring = [(0, 0), (1, 0), (1, 1), (0, 1)]
        for e in range(4):
            (u0, w0), (u1, w1) = ring[e], ring[(e + 1) % 4]
            p = corner_id(*perm(a + 1, b + u0, cc + w0))
            q = corner_id(*perm(a + 1, b + u1, cc + w1))
            tets.append(np.stack([m0, m1, p, q], -1))
    indices = np.concatenate(tets, 0).astype(np.int64)
    v = vertices[indices]
    vol = np.einsum("ij,ij->i", np.cross(v[:, 1] - v[:, 0], v[:, 2] - v[:, 0]), v[:, 3] - v[:, 0])
    flip = vol < 0
    indices[flip, 2], indices[flip, 3] = indices[flip, 3].copy(), indices[flip, 2].copy()
    return vertices, indices


def save_tets_npz(path: str, n_cells: int):
    """Write the grid in the format `GShellTetsGeometry(..., tet_init_file=path)` loads."""
    v, t = bcc_tet_grid(n_cells)
    np.savez_compressed(path, vertices=v, indices=t)
    return v, t
