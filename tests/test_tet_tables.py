"""Host logic: the static edge tables reproduce the reference's per-step unique() numbering."""
import numpy as np
import torch

from gshell_b200.geometry.tet_tables import TetTables
from gshell_b200.grids import bcc_tet_grid
from oracle.mt_oracle import crossing_edges


def test_static_tables_rank_equals_reference_unique():
    v, t = bcc_tet_grid(7)
    tets = torch.tensor(t)
    tab = TetTables(tets, v.shape[0])
    assert tab.tet_v.dtype == torch.int32 and tab.tet_e.shape == (t.shape[0], 6)
    # edge_v sorted lexicographically, unique
    key = tab.edge_v[:, 0].long() * v.shape[0] + tab.edge_v[:, 1].long()
    assert torch.all(key[1:] > key[:-1])
    # each tet edge id points at the right (lo,hi)
    ends = tets[:, [0, 1, 0, 2, 0, 3, 1, 2, 1, 3, 2, 3]].reshape(-1, 6, 2)
    lo, hi = ends.min(-1).values, ends.max(-1).values
    got = tab.edge_v[tab.tet_e.long()]
    assert torch.equal(got[..., 0].long(), lo) and torch.equal(got[..., 1].long(), hi)
    # vertex numbering: k-th crossing edge of the static list == reference numbering
    torch.manual_seed(0)
    sdf = torch.rand(v.shape[0]) - 0.3
    _, _, _, ref_edges = crossing_edges(sdf, tets, "rows")
    occ = sdf > 0
    cross = occ[tab.edge_v[:, 0].long()] != occ[tab.edge_v[:, 1].long()]
    assert torch.equal(tab.edge_v[cross].long(), ref_edges)


def test_bcc_counts():
    for n in (2, 3, 5):
        v, t = bcc_tet_grid(n)
        assert v.shape[0] == (n + 1) ** 3 + n ** 3 and t.shape[0] == 12 * n * n * (n - 1)
        assert len(np.unique(t)) <= v.shape[0]
