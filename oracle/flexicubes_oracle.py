"""ORACLE (test infrastructure, not product code): CPU restatement of G-FlexiCubes extraction.

Restates `GShellFlexiCubes.__call__` of the reference (geometry/gshell_flexicubes.py:136-230 and helpers
:242-591; tables geometry/flexicubes_table.py) with plain PyTorch CPU ops; autograd provides the gradient oracle.
Only tests/, bench.py's cpu_baseline leg and smoke() may import this file.

Parity pin: tests/golden/flex_*.npz hold outputs of the UNMODIFIED reference run on CPU through
tests/golden/_ref_shim.py (generator tests/golden/make_golden_flex.py); tests/test_oracle_flex.py checks faces
bit-for-bit and positions / mSDF values exactly (same IEEE op order: per-dual-vertex sums are accumulated in the
reference's `index_add_` order, i.e. slot 0..6 of each dmc_table row).

Quirks of the reference that are restated on purpose:
  * `nu_d` is mutated in place at :476-477, so the returned nu_d = S1/beta + S2 and nu_d_stopvgd = (S1/beta + S2)/beta;
  * mSDF occupancy is `>= 0` (:556) (tets use `> 0`);
  * when no face is fully inside the mSDF the watertight mesh is returned unchanged (:566-567).
"""
import os

import numpy as np
import torch

_T = None


def tables():
    """dmc_table [256,4,7], num_vd_table [256], check_table [256,5], cut tables -- data file shipped with the product."""
    global _T
    if _T is None:
        here = os.path.dirname(os.path.abspath(__file__))
        z = np.load(os.path.join(here, "..", "gshell_b200", "geometry", "flexicubes_tables.npz"))
        _T = {k: torch.from_numpy(z[k]).long() for k in z.files}
    return _T


CUBE_EDGES = [0, 1, 1, 5, 4, 5, 0, 4, 2, 3, 3, 7, 6, 7, 2, 6, 2, 0, 3, 1, 7, 5, 6, 4]     # :86-87 (oriented pairs)


def voxel_grid(res):
    """construct_voxel_grid (:103-134): vertices (x-major lexicographic) - 0.5 and cube corner ids (corner k = x+2y+4z)."""
    g = torch.arange(res + 1, dtype=torch.float32) / res
    verts = torch.stack(torch.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    verts = torch.round(verts * 10 ** 5) / (10 ** 5)
    i, j, k = torch.meshgrid(torch.arange(res), torch.arange(res), torch.arange(res), indexing="ij")
    corners = []
    for c in range(8):
        dx, dy, dz = c & 1, (c >> 1) & 1, (c >> 2) & 1
        corners.append(((i + dx) * (res + 1) + (j + dy)) * (res + 1) + (k + dz))
    return verts - 0.5, torch.stack(corners, -1).reshape(-1, 8)


def _lerp0(w, x):
    """_linear_interp (:345-355): zero crossing of w along the pair axis (dim -2)."""
    ww = torch.cat([w[..., 1:2, :], -w[..., 0:1, :]], -2)
    return (x * ww).sum(-2) / ww.sum(-2)


def _lerp0_nonan(w, x):
    """_linear_interp_nonan (:357-371)."""
    ww = torch.cat([w[:, 1:2], -w[:, 0:1]], 1)
    den = ww.sum(1, keepdim=True).expand(-1, 2, 1)
    ok = (den.abs() > 0).detach()
    scale = torch.where(ok, ww / torch.where(ok, den, torch.ones_like(den)), torch.zeros_like(ww))
    return (x * scale).sum(1)


def gflexicubes(x, s, nu, cubes, res, beta=None, alpha=None, gamma=None, weight_scale=0.99):
    """-> (vertices_open, faces_open, L_dev, extra) as the reference (:214-224)."""
    T = tables()
    occ = s < 0
    occ8 = occ[cubes]
    n_in = occ8.sum(-1)
    surf = (n_in > 0) & (n_in < 8)
    if int(surf.sum()) == 0:
        return torch.zeros((0, 3)), torch.zeros((0, 3), dtype=torch.long), torch.zeros((0,)), None
    n_cubes = cubes.shape[0]
    beta = (torch.tanh(beta) * weight_scale + 1) if beta is not None else torch.ones((n_cubes, 12))
    alpha = (torch.tanh(alpha) * weight_scale + 1) if alpha is not None else torch.ones((n_cubes, 8))
    gamma = (torch.sigmoid(gamma) * weight_scale + (1 - weight_scale) / 2) if gamma is not None else torch.ones((n_cubes,))
    beta, alpha, gamma = beta[surf], alpha[surf], gamma[surf]

    with torch.no_grad():
        # ---- case ids with the C16/C19 ambiguity fix (:266-306) -------------------------------------------------
        case = (occ8[surf].long() * (2 ** torch.arange(8))).sum(-1)
        chk = T["check_table"][case]
        coords = torch.stack(torch.meshgrid(torch.arange(res), torch.arange(res), torch.arange(res), indexing="ij"), -1).reshape(-1, 3)
        sc = coords[surf]
        flagged = torch.zeros(res, res, res, dtype=torch.bool)
        amb = chk[:, 0] == 1
        flagged[sc[amb, 0], sc[amb, 1], sc[amb, 2]] = True
        adj = sc + chk[:, 1:4]
        inside = ((adj >= 0) & (adj < res)).all(-1)
        adjc = adj.clamp(0, res - 1)
        invert = amb & inside & flagged[adjc[:, 0], adjc[:, 1], adjc[:, 2]]
        case = torch.where(invert, chk[:, 4], case)

        # ---- surface edges (:309-331) ------------------------------------------------------------------------------
        sc_cubes = cubes[surf]
        pairs = sc_cubes[:, CUBE_EDGES].reshape(-1, 2)
        uniq, inv, counts = torch.unique(pairs, dim=0, return_inverse=True, return_counts=True)
        crossing = occ[uniq].sum(-1) == 1
        eid = torch.full((uniq.shape[0],), -1, dtype=torch.long)
        eid[crossing] = torch.arange(int(crossing.sum()))
        idx_map = eid[inv].reshape(-1, 12)                    # per surf cube, per local edge: crossing-edge id or -1
        slot_cross = crossing[inv]
        slot_count = counts[inv]
        surf_edges = uniq[crossing]

        # ---- dual-vertex numbering (:398-421): groups by num_vd ascending, cube order, k ---------------------------
        num_vd = T["num_vd_table"][case]
        n_sc = case.shape[0]
        vd_base = torch.zeros(n_sc, dtype=torch.long)
        total = 0
        for num in torch.unique(num_vd).tolist():
            sel = num_vd == num
            cnt = int(sel.sum())
            vd_base[sel] = total + torch.arange(cnt) * num
            total += cnt * num
        n_vd = total
        dmc = T["dmc_table"][case]                             # [n_sc, 4, 7]
        k_ids = torch.arange(4).view(1, 4).expand(n_sc, 4)
        vd_of = vd_base[:, None] + k_ids                       # [n_sc,4]
        live = k_ids < num_vd[:, None]
        vd_cube = torch.arange(n_sc).view(-1, 1).expand(n_sc, 4)[live]
        order = torch.argsort(vd_of[live])
        vd_cube = vd_cube[order]                               # surf-cube rank of each dual vertex
        vd_edges = dmc[live][order]                            # [n_vd, 7] local edge ids or -1
        vd_slots = vd_edges >= 0
        vd_num_edges = vd_slots.sum(-1, keepdim=True)
        le = vd_edges.clamp(min=0)
        ce = idx_map[vd_cube[:, None].expand(-1, 7), le]       # crossing-edge id of each slot
        vd_gamma = gamma[vd_cube]
        # vd_idx_map (:480-483)
        vd_idx_map = torch.zeros(n_sc * 12, dtype=torch.long)
        flat = (vd_cube[:, None] * 12 + le)[vd_slots]
        vd_idx_map[flat] = torch.arange(n_vd).view(-1, 1).expand(-1, 7)[vd_slots]

    # ---- dual vertices (:391-396, :452-478) --------------------------------------------------------------------------
    ex = x[surf_edges]                                         # [E,2,3]
    es = s[surf_edges].unsqueeze(-1)                           # [E,2,1]
    enu = nu[surf_edges].unsqueeze(-1)
    zero_crossing = _lerp0(es, ex)                             # [E,3]
    alpha_pairs = alpha[:, CUBE_EDGES].reshape(-1, 12, 2)
    ces = ce.clamp(min=0)
    a_slot = alpha_pairs[vd_cube[:, None].expand(-1, 7), le].unsqueeze(-1)       # [n_vd,7,2,1]
    coeff = es[ces] * a_slot
    ue = _lerp0(coeff, ex[ces])                                # [n_vd,7,3]
    nue = _lerp0(coeff, enu[ces])                              # [n_vd,7,1]
    nue_sg = _lerp0(coeff.detach(), enu[ces])
    b_slot = beta[vd_cube[:, None].expand(-1, 7), le].unsqueeze(-1)             # [n_vd,7,1]
    m = vd_slots.unsqueeze(-1)
    beta_sum = torch.zeros(n_vd, 1)
    acc_v = torch.zeros(n_vd, 3)
    s1 = torch.zeros(n_vd, 1)
    s2 = torch.zeros(n_vd, 1)
    for j in range(7):                                         # index_add_ order of the reference
        mj = m[:, j]
        beta_sum = beta_sum + torch.where(mj, b_slot[:, j], torch.zeros_like(beta_sum))
        acc_v = acc_v + torch.where(mj, ue[:, j] * b_slot[:, j], torch.zeros_like(acc_v))
        s1 = s1 + torch.where(mj, nue[:, j] * b_slot[:, j], torch.zeros_like(s1))
    vd = acc_v / beta_sum
    nu_d = s1 / beta_sum
    for j in range(7):                                         # in-place aliasing quirk (:476-477)
        mj = m[:, j]
        nu_d = nu_d + torch.where(mj, nue_sg[:, j] * b_slot[:, j].detach(), torch.zeros_like(nu_d))
    nu_d_sg = nu_d / beta_sum.detach()
    # L_dev (:232-240)
    dist = (zero_crossing[ces] - vd[:, None, :]).norm(dim=-1)                    # [n_vd,7]
    mean_l2 = torch.zeros(n_vd)
    for j in range(7):
        mean_l2 = mean_l2 + torch.where(vd_slots[:, j], dist[:, j], torch.zeros_like(mean_l2))
    mean_l2 = mean_l2 / vd_num_edges.squeeze(1).float()
    L_dev = (dist - mean_l2[:, None]).abs()[vd_slots]

    # ---- quads -> triangles (:487-552, non-training split) ---------------------------------------------------------------
    with torch.no_grad():
        gm = ((slot_count == 4) & slot_cross)
        group = idx_map.reshape(-1)[gm]
        vdi = vd_idx_map[gm]
        e_sorted, perm = torch.sort(group, stable=True)
        quad = vdi[perm].reshape(-1, 4)
        s_e = s[surf_edges[e_sorted.reshape(-1, 4)[:, 0]]]
        flip = s_e[:, 0] > 0
        quad = torch.cat([quad[flip][:, [0, 1, 3, 2]], quad[~flip][:, [2, 3, 1, 0]]])
    qg = vd_gamma[quad]
    g02, g13 = qg[:, 0] * qg[:, 2], qg[:, 1] * qg[:, 3]
    with torch.no_grad():
        first = g02 > g13
        faces = torch.where(first[:, None], quad[:, [0, 1, 2, 0, 2, 3]], quad[:, [0, 1, 3, 3, 1, 2]]).reshape(-1, 3)

    # ---- open-surface cut (:554-591) ----------------------------------------------------------------------------------------
    nus, nus_sg = nu_d, nu_d_sg
    with torch.no_grad():
        mocc = (nus >= 0)[faces.reshape(-1), 0].reshape(-1, 3)
        msum = mocc.sum(-1)
        uncut, cut = msum == 3, (msum < 3) & (msum > 0)
    if int(uncut.sum()) == 0:
        extra = {"n_verts_watertight": vd.shape[0], "vertices_watertight": vd, "faces_watertight": faces, "msdf": nus,
                 "msdf_watertight": nus, "msdf_boundary": nus[:1].detach() * 0.0}
        return vd, faces, L_dev, extra
    cf = faces[cut]
    pair_idx = cf[:, [0, 1, 1, 2, 2, 0]].reshape(-1)
    pv = vd[pair_idx].view(-1, 2, 3)
    pn = nus[pair_idx].view(-1, 2, 1)
    pn_sg = nus_sg[pair_idx].view(-1, 2, 1)
    bverts = _lerp0_nonan(pn, pv)
    bnu_sg = _lerp0_nonan(pn_sg.detach(), pn_sg)
    vertices_open = torch.cat([vd, bverts], 0)
    nus_open_sg = torch.cat([nus_sg, bnu_sg], 0)
    with torch.no_grad():
        code = (mocc[cut].long() * torch.tensor([4, 2, 1])).sum(-1)
        ids = torch.cat([cf, vd.shape[0] + torch.arange(cf.shape[0] * 3).view(-1, 3)], -1)
        ntri = T["gflex_num_triangles_table"][code]
        conf = T["gflex_configuration_table"]
        one, two = ntri == 1, ntri == 2
        faces_open = torch.cat([faces[uncut],
                                torch.gather(ids[one], 1, conf[code[one]][:, :3]).view(-1, 3),
                                torch.gather(ids[two], 1, conf[code[two]][:, :6]).view(-1, 3)])
    extra = {"n_verts_watertight": vd.shape[0], "vertices_watertight": vd, "faces_watertight": faces,
             "msdf": nus_open_sg, "msdf_watertight": nus, "msdf_boundary": bnu_sg}
    return vertices_open, faces_open, L_dev, extra
