"""ORACLE (test infrastructure): PyTorch restatement of the multiresolution hash-grid encoding behind the reference's
learned material field.

PARITY UNPINNED: the reference takes the encoding from tiny-cuda-nn (`tcnn.Encoding(3, {"otype": "HashGrid", "n_levels": 16,
"n_features_per_level": 2, "log2_hashmap_size": 19, "base_resolution": 16, "per_level_scale": exp(log(4096 / 16) / 15)})`,
render/mlptexture.py:59-73; `pip install git+https://github.com/NVlabs/tiny-cuda-nn`, unpinned, reference README.md:39), which is
absent from this environment, and the reference holds no test or golden vector for it.  This file restates the published algorithm
(Mueller et al. 2022, "Instant neural graphics primitives", section 3; tiny-cuda-nn's grid encoding): per level
scale = base * s^l - 1, resolution = ceil(scale) + 1, entries = min(round_up(resolution^3, 8), 2^log2_hashmap_size) (tiny-cuda-nn
evaluates the scale with float32 exp2f/log2f, which moves it by an ulp or two: one more thing that is unpinned); a point
x in [0,1]^3 maps to pos = x * scale + 0.5, its 8 surrounding lattice points are looked up by the dense index
x + y res + z res^2 while the stride stays within the level's entries and by the hash x ^ y * 2654435761 ^ z * 805459861 otherwise
(both modulo the entries), and their features are blended trilinearly.  Only tests/ may import this module.
"""
import math

import torch

PRIMES = (1, 2654435761, 805459861)


def levels(n_levels=16, base_resolution=16, per_level_scale=None, log2_hashmap_size=19, desired_resolution=4096):
    """-> (offsets [L+1], resolutions [L], scales [L]); scales are float32-representable Python floats."""
    if per_level_scale is None:
        per_level_scale = math.exp(math.log(desired_resolution / base_resolution) / (n_levels - 1))
    offs, ress, scales = [0], [], []
    for lvl in range(n_levels):
        scale = base_resolution * math.pow(per_level_scale, lvl) - 1.0
        if abs(scale - round(scale)) < 1e-9 * max(1.0, abs(scale)):      # 16 * 256 - 1 must not become 4095.000000000001
            scale = float(round(scale))
        scale = float(torch.tensor(scale, dtype=torch.float32))
        res = int(math.ceil(scale)) + 1
        n = min((res ** 3 + 7) // 8 * 8, 1 << log2_hashmap_size)
        offs.append(offs[-1] + n)
        ress.append(res)
        scales.append(scale)
    return offs, ress, scales


def entry_index(ix, iy, iz, res, size):
    """int64 tensors of lattice coordinates -> entry index inside the level (uint32 arithmetic)."""
    M = 0xFFFFFFFF
    stride, idx = 1, ix.clone()
    stride *= res
    if stride <= size:
        idx = idx + iy * stride
        stride *= res
    if stride <= size:
        idx = idx + iz * stride
        stride *= res
    if size < stride:
        idx = (ix * PRIMES[0] & M) ^ (iy * PRIMES[1] & M) ^ (iz * PRIMES[2] & M)
    return (idx & M) % size


def encode(x01, table, offs, ress, scales):
    """x01 [N,3] in [0,1], table [E,2] (all levels) -> [N, 2 L]; differentiable w.r.t. x01 and table through torch autograd."""
    outs = []
    for lvl, (res, scale) in enumerate(zip(ress, scales)):
        size = offs[lvl + 1] - offs[lvl]
        tab = table[offs[lvl]:offs[lvl + 1]]
        pos = x01 * scale + 0.5
        cell = torch.floor(pos).detach()
        w = pos - cell
        c = cell.to(torch.int64)
        acc = 0
        for corner in range(8):
            d = [(corner >> k) & 1 for k in range(3)]
            wgt = 1
            for k in range(3):
                wgt = wgt * (w[:, k] if d[k] else 1 - w[:, k])
            e = entry_index(c[:, 0] + d[0], c[:, 1] + d[1], c[:, 2] + d[2], res, size)
            acc = acc + wgt[:, None] * tab[e]
        outs.append(acc)
    return torch.cat(outs, -1)
