"""Static topology tables of a tet grid (built once per grid; pure torch ops, CPU or CUDA).

The reference rediscovers the surface edges every step with `torch.unique(all_edges, dim=0)` over the
edges of the valid tets (geometry/gshell_tets.py:266-268).  Its vertex numbering is the rank of each
sign-crossing edge in the lexicographically sorted unique (lo, hi) edge list; restricted to crossing
edges that rank is the same whether the list holds the valid tets' edges or ALL grid edges, so the
sorted list can be computed once and per-step work reduces to a scan over a crossing flag.
"""
import torch

# local endpoints of the 6 tet edges, reference `base_tet_edges` (gshell_tets.py:178)
TET_EDGE_ENDS = (0, 1, 0, 2, 0, 3, 1, 2, 1, 3, 2, 3)


class TetTables:
    """tet_v int32[T,4], tet_e int32[T,6] (edge id per local edge), edge_v int32[E,2] sorted (lo,hi)."""

    def __init__(self, tet_fx4: torch.Tensor, n_verts: int):
        assert tet_fx4.dim() == 2 and tet_fx4.shape[1] == 4
        t = tet_fx4.long()
        n_verts = int(n_verts)
        ends = t[:, list(TET_EDGE_ENDS)].reshape(-1, 2)
        lo = torch.minimum(ends[:, 0], ends[:, 1])
        hi = torch.maximum(ends[:, 0], ends[:, 1])
        key, inverse = torch.unique(lo * n_verts + hi, return_inverse=True)
        self.n_verts = n_verts
        self.n_tets = int(t.shape[0])
        self.n_edges = int(key.shape[0])
        assert n_verts < 2 ** 31 and self.n_edges < 2 ** 31 and 4 * self.n_tets + self.n_edges < 2 ** 31
        self.tet_v = t.to(torch.int32).contiguous()
        self.tet_e = inverse.reshape(-1, 6).to(torch.int32).contiguous()
        self.edge_v = torch.stack([key // n_verts, key % n_verts], -1).to(torch.int32).contiguous()
        self.device = tet_fx4.device
        self._workspace = None
        self._counts = None
        self._counts_host = None

    # per-grid scratch reused by every extraction on this grid (sized by the library)
    def workspace(self, nbytes: int):
        if self._workspace is None or self._workspace.numel() < nbytes:
            self._workspace = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        return self._workspace

    def counts_buffers(self, n: int):
        if self._counts is None:
            self._counts = torch.zeros(n, dtype=torch.int32, device=self.device)
            self._counts_host = torch.zeros(n, dtype=torch.int32).pin_memory() if self.device.type == "cuda" \
                else torch.zeros(n, dtype=torch.int32)
        return self._counts, self._counts_host


_CACHE = {}


def tables_for(tet_fx4: torch.Tensor, n_verts: int) -> TetTables:
    """Cache keyed on the identity of the index tensor (the reference passes the same `self.indices`
    every iteration, gshell_tets_geometry.py:206-207)."""
    key = (tet_fx4.data_ptr(), tuple(tet_fx4.shape), str(tet_fx4.device), tet_fx4._version, int(n_verts))
    tab = _CACHE.get(key)
    if tab is None:
        if len(_CACHE) >= 4:
            _CACHE.pop(next(iter(_CACHE)))
        tab = _CACHE[key] = TetTables(tet_fx4, n_verts)
        tab.source = tet_fx4            # keeps the keyed storage alive: its address cannot be recycled while cached
    return tab
