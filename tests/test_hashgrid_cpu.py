"""Hash-grid encoding of the material field: level layout of the product against the oracle's, and the oracle's own properties
(PARITY UNPINNED against tiny-cuda-nn, which is absent: oracle/hashgrid_oracle.py)."""
import numpy as np
import torch

from gshell_b200.render import mlptexture
from oracle import hashgrid_oracle as ho


def test_level_layout_matches_oracle():
    for kw in ({}, {"n_levels": 8, "base_resolution": 4, "desired_resolution": 256, "log2_hashmap_size": 12}):
        offs, ress, scales = mlptexture.hashgrid_levels(**kw)
        o, r, s = ho.levels(**kw)
        assert offs.tolist() == o and ress.tolist() == r
        assert np.array_equal(scales, np.asarray(s, dtype=np.float32))
    offs, ress, scales = mlptexture.hashgrid_levels()
    assert len(ress) == 16 and ress[0] == 16 and ress[-1] == 4096 and scales[-1] == 4095.0
    assert all(int(offs[i + 1] - offs[i]) <= 1 << 19 for i in range(16))


def test_oracle_interpolates_lattice_values_and_partitions_unity():
    o, r, s = ho.levels(n_levels=3, base_resolution=4, desired_resolution=16, log2_hashmap_size=19)      # all three levels dense
    g = torch.Generator().manual_seed(0)
    table = torch.randn(o[-1], 2, generator=g)
    # a lattice point of level 0: pos = x * scale + 0.5 = k exactly -> the encoding is that entry
    k = torch.tensor([[1, 2, 3], [0, 0, 0], [3, 1, 2]])
    x = (k.float() - 0.5) / s[0]
    x = x.clamp(0, 1)
    enc = ho.encode(x[:1], table, o, r, s)
    e = int(k[0, 0] + k[0, 1] * r[0] + k[0, 2] * r[0] * r[0])
    assert torch.allclose(enc[0, :2], table[e], atol=1e-6)
    ones = torch.ones(o[-1], 2)
    pts = torch.rand(64, 3, generator=g)
    assert torch.allclose(ho.encode(pts, ones, o, r, s), torch.ones(64, 6), atol=1e-6)


def test_oracle_hashes_fine_levels_in_uint32():
    o, r, s = ho.levels()
    lvl = 15
    size = o[lvl + 1] - o[lvl]
    assert size == 1 << 19 and r[lvl] ** 2 > size
    ix, iy, iz = torch.tensor([4095]), torch.tensor([4096]), torch.tensor([17])
    want = ((4095 * 1) ^ ((4096 * 2654435761) & 0xFFFFFFFF) ^ ((17 * 805459861) & 0xFFFFFFFF)) % size
    assert int(ho.entry_index(ix, iy, iz, r[lvl], size)) == want
    # a level whose x-y plane still fits but whose cube does not is hashed as well
    lvl = 6
    size = o[lvl + 1] - o[lvl]
    assert r[lvl] ** 2 <= size < r[lvl] ** 3
    got = int(ho.entry_index(torch.tensor([3]), torch.tensor([5]), torch.tensor([7]), r[lvl], size))
    assert got == ((3 * 1) ^ ((5 * 2654435761) & 0xFFFFFFFF) ^ ((7 * 805459861) & 0xFFFFFFFF)) % size
