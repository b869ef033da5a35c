"""Marching-tets look-up tables as plain data (values: reference geometry/gshell_tets.py:82-181).  Single source for the
generated CUDA header (csrc/gen_mt_tables.py -> csrc/mt_tables.cuh) and for host-side torch code that needs them."""
import torch

# watertight triangles per SDF case (tet-local edge ids), `triangle_table` :83-100
TRI = [[-1]*6,[1,0,2,-1,-1,-1],[4,0,3,-1,-1,-1],[1,4,2,1,3,4],[3,1,5,-1,-1,-1],[2,3,0,2,5,3],
       [1,4,0,1,5,4],[4,2,5,-1,-1,-1],[4,5,2,-1,-1,-1],[4,1,0,4,5,1],[3,2,0,3,5,2],[1,3,5,-1,-1,-1],
       [4,1,2,4,3,1],[3,0,4,-1,-1,-1],[2,0,1,-1,-1,-1],[-1]*6]
# polygon loop per SDF case, `mesh_edge_table` :102-119 (first 4 entries; the loop closes on entry 0)
LOOP = [[-1]*4,[1,0,2,1],[4,0,3,4],[1,3,4,2],[3,1,5,3],[2,5,3,0],[1,5,4,0],[4,2,5,4],
        [4,5,2,4],[4,5,1,0],[3,5,2,0],[1,3,5,1],[4,3,1,2],[3,0,4,3],[2,0,1,2],[-1]*4]
NTRI = [0,1,1,2,1,2,2,1,1,2,2,1,2,1,1,0]
# mSDF cut of a triangle / quad polygon, `triangle_table_tri` :122-139, `triangle_table_quad` :141-176
CUT3 = [[-1]*6,[4,2,5,-1,-1,-1],[3,1,4,-1,-1,-1],[3,1,2,3,2,5],[0,3,5,-1,-1,-1],[0,3,4,0,4,2],
        [0,1,4,0,4,5],[0,1,2,-1,-1,-1]]
_m = -1
CUT4 = [[_m]*12,[6,3,7]+[_m]*9,[5,2,6]+[_m]*9,[5,2,7,3,7,2]+[_m]*6,[4,1,5]+[_m]*9,
        [4,1,5,4,5,7,5,6,7,7,6,3],[4,1,2,6,4,2]+[_m]*6,[4,1,2,7,4,2,7,2,3]+[_m]*3,[0,4,7]+[_m]*9,
        [0,4,6,3,0,6]+[_m]*6,[0,4,5,0,5,2,0,2,6,0,6,7],[0,4,5,0,5,2,0,2,3]+[_m]*3,
        [0,1,5,7,0,5]+[_m]*6,[0,1,5,0,5,6,0,6,3]+[_m]*3,[0,1,2,0,2,6,0,6,7]+[_m]*3,[0,1,2,0,2,3]+[_m]*6]
NCUT3 = [0,1,1,2,1,2,2,1]
NCUT4 = [0,1,1,2,1,4,2,3,1,2,4,3,2,3,3,2]

_CACHE = {}


def luts(device):
    """The tables as int64 tensors on `device` (negative entries clamped to 0: they are never selected)."""
    key = str(device)
    if key not in _CACHE:
        t = lambda x: torch.tensor(x, dtype=torch.long, device=device).clamp(min=0)     # noqa: E731
        _CACHE[key] = dict(tri=t(TRI), loop=t(LOOP), ntri=t(NTRI), cut3=t(CUT3), cut4=t(CUT4), ncut3=t(NCUT3), ncut4=t(NCUT4))
    return _CACHE[key]


_CACHE_I32 = {}


def luts_i32(device):
    """The tables in the order the decode kernels take them (csrc/auggrid.cu, `luts7_host`): tri, loop, ntri, cut3, cut4, ncut3,
    ncut4 as contiguous int32 tensors on `device`, negative entries clamped to 0."""
    key = str(device)
    if key not in _CACHE_I32:
        d = luts(device)
        _CACHE_I32[key] = [d[k].int().contiguous() for k in ("tri", "loop", "ntri", "cut3", "cut4", "ncut3", "ncut4")]
    return _CACHE_I32[key]
