"""Static tables of a FlexiCubes voxel grid (built once per grid with torch ops) + the DMC look-up tables.

The reference rebuilds `torch.unique(all_edges, dim=0, return_inverse, return_counts)` over the surface cubes' 12
ORIENTED edges every step (geometry/gshell_flexicubes.py:316-317) and later groups the cubes around each edge with a
stable sort (:496).  Both orders are properties of the grid alone, so they are computed once here."""
import os

import numpy as np
import torch

# (first, second) corner of the 12 cube edges, reference `cube_edges` (:86-87)
CUBE_EDGES = (0, 1, 1, 5, 4, 5, 0, 4, 2, 3, 3, 7, 6, 7, 2, 6, 2, 0, 3, 1, 7, 5, 6, 4)

_LUT = {}


def luts(device):
    """check_table int16[256,5], num_vd_table int8[256], dmc_table int8[256,4,7], cut tables int8 -- on `device`."""
    key = str(device)
    if key not in _LUT:
        z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "flexicubes_tables.npz"))
        _LUT[key] = {
            "check": torch.from_numpy(z["check_table"].astype(np.int16)).contiguous().to(device),
            "num_vd": torch.from_numpy(z["num_vd_table"].astype(np.int8)).contiguous().to(device),
            "dmc": torch.from_numpy(z["dmc_table"].astype(np.int8)).contiguous().to(device),
            "ntri": torch.from_numpy(z["gflex_num_triangles_table"].astype(np.int8)).contiguous().to(device),
            "conf": torch.from_numpy(z["gflex_configuration_table"].astype(np.int8)).contiguous().to(device),
        }
    return _LUT[key]


class FlexTables:
    def __init__(self, cube_fx8: torch.Tensor, n_verts: int):
        c = cube_fx8.long()
        self.n_cubes = int(c.shape[0])
        self.n_verts = int(n_verts)
        pairs = c[:, list(CUBE_EDGES)].reshape(-1, 2)
        key, inverse, counts = torch.unique(pairs[:, 0] * n_verts + pairs[:, 1], return_inverse=True, return_counts=True)
        self.n_edges = int(key.shape[0])
        assert int(counts.max()) <= 4, "more than 4 cubes around an edge: not a regular voxel grid"
        self.cube_v = c.to(torch.int32).contiguous()
        self.cube_e = inverse.reshape(-1, 12).to(torch.int32).contiguous()
        self.edge_v = torch.stack([key // n_verts, key % n_verts], -1).to(torch.int32).contiguous()
        self.edge_cnt = counts.to(torch.uint8).contiguous()
        order = torch.argsort(inverse, stable=True)               # slots (cube*12 + local edge) grouped by edge, ascending
        start = torch.cumsum(counts, 0) - counts
        j = torch.arange(4, device=c.device).view(1, 4)
        idx = (start.view(-1, 1) + j).clamp(max=order.shape[0] - 1)
        slots = torch.where(j < counts.view(-1, 1), order[idx], torch.zeros_like(idx))
        self.edge_slots = slots.to(torch.int32).contiguous()
        self.device = cube_fx8.device


_CACHE = {}


def tables_for(cube_fx8, n_verts):
    key = (cube_fx8.data_ptr(), tuple(cube_fx8.shape), str(cube_fx8.device), cube_fx8._version, int(n_verts))
    t = _CACHE.get(key)
    if t is None:
        if len(_CACHE) >= 4:
            _CACHE.pop(next(iter(_CACHE)))
        t = _CACHE[key] = FlexTables(cube_fx8, n_verts)
        t.source = cube_fx8             # keeps the keyed storage alive: its address cannot be recycled while cached
    return t
