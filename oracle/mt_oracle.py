"""ORACLE (test infrastructure, not product code): CPU restatement of G-Shell marching tetrahedra.

Restates `GShell_Tets.__call__` of the reference (geometry/gshell_tets.py:245-443) with plain
PyTorch CPU ops so that autograd provides the gradient oracle.  Only `tests/`, `bench.py`'s
cpu_baseline / `--impl reference` legs and `__graft_entry__.smoke()` may import this file; the
product path (`gshell_b200/`) never does.

Parity pin: `tests/golden/mt_*.npz` hold outputs of the UNMODIFIED reference run in the build
container through `tests/golden/_ref_shim.py` (generator: `tests/golden/make_golden_mt.py`);
`tests/test_oracle_mt.py` checks this restatement against them bit-for-bit (faces) and to 0 ulp
(vertex positions / mSDF values; same IEEE op order).

Stage map (reference line numbers):
  crossing_edges      :250-276   valid tets, unique sorted edges, crossing-edge ids
  lerp_on_sdf         :277-290   zero-crossing vertices and interpolated mSDF (+ stop-grad twin)
  watertight_faces    :293-316   case index, 1-triangle group then 2-triangle group
  tangent_frame       :9-78,210-239,318-319   uv atlas, smooth normals, tangents
  polygon_loops       :323-331   per-tet polygon edges and mSDF occupancy
  boundary_weights    :345-365   mSDF zero-crossing weights on polygon edges
  cut_faces           :394-416   six face groups via the tri/quad cut tables
"""
import math

import torch

# ---- look-up tables (values are the reference's: gshell_tets.py:82-181) -------------------------
TRI_TABLE = [[-1] * 6, [1, 0, 2, -1, -1, -1], [4, 0, 3, -1, -1, -1], [1, 4, 2, 1, 3, 4],
             [3, 1, 5, -1, -1, -1], [2, 3, 0, 2, 5, 3], [1, 4, 0, 1, 5, 4], [4, 2, 5, -1, -1, -1],
             [4, 5, 2, -1, -1, -1], [4, 1, 0, 4, 5, 1], [3, 2, 0, 3, 5, 2], [1, 3, 5, -1, -1, -1],
             [4, 1, 2, 4, 3, 1], [3, 0, 4, -1, -1, -1], [2, 0, 1, -1, -1, -1], [-1] * 6]
LOOP_TABLE = [[-1] * 6, [1, 0, 2, 1, -1, -1], [4, 0, 3, 4, -1, -1], [1, 3, 4, 2, 1, -1],
              [3, 1, 5, 3, -1, -1], [2, 5, 3, 0, 2, -1], [1, 5, 4, 0, 1, -1], [4, 2, 5, 4, -1, -1],
              [4, 5, 2, 4, -1, -1], [4, 5, 1, 0, 4, -1], [3, 5, 2, 0, 3, -1], [1, 3, 5, 1, -1, -1],
              [4, 3, 1, 2, 4, -1], [3, 0, 4, 3, -1, -1], [2, 0, 1, 2, -1, -1], [-1] * 6]
CUT_TRI = [[-1] * 6, [4, 2, 5, -1, -1, -1], [3, 1, 4, -1, -1, -1], [3, 1, 2, 3, 2, 5],
           [0, 3, 5, -1, -1, -1], [0, 3, 4, 0, 4, 2], [0, 1, 4, 0, 4, 5], [0, 1, 2, -1, -1, -1]]
_m = -1
CUT_QUAD = [[_m] * 12,
            [6, 3, 7] + [_m] * 9, [5, 2, 6] + [_m] * 9, [5, 2, 7, 3, 7, 2] + [_m] * 6,
            [4, 1, 5] + [_m] * 9, [4, 1, 5, 4, 5, 7, 5, 6, 7, 7, 6, 3], [4, 1, 2, 6, 4, 2] + [_m] * 6,
            [4, 1, 2, 7, 4, 2, 7, 2, 3] + [_m] * 3, [0, 4, 7] + [_m] * 9, [0, 4, 6, 3, 0, 6] + [_m] * 6,
            [0, 4, 5, 0, 5, 2, 0, 2, 6, 0, 6, 7], [0, 4, 5, 0, 5, 2, 0, 2, 3] + [_m] * 3,
            [0, 1, 5, 7, 0, 5] + [_m] * 6, [0, 1, 5, 0, 5, 6, 0, 6, 3] + [_m] * 3,
            [0, 1, 2, 0, 2, 6, 0, 6, 7] + [_m] * 3, [0, 1, 2, 0, 2, 3] + [_m] * 6]
N_TRI = [0, 1, 1, 2, 1, 2, 2, 1, 1, 2, 2, 1, 2, 1, 1, 0]
N_CUT_TRI = [0, 1, 1, 2, 1, 2, 2, 1]
N_CUT_QUAD = [0, 1, 1, 2, 1, 4, 2, 3, 1, 2, 4, 3, 2, 3, 3, 2]
TET_EDGE_ENDS = [0, 1, 0, 2, 0, 3, 1, 2, 1, 3, 2, 3]


def _t(x):
    return torch.tensor(x, dtype=torch.long)


def crossing_edges(sdf, tets, unique_mode="rows", msdf=None):
    """-> (valid mask [T], case index [Tv], edge->vertex id map [Tv,6], crossing edges [Vw,2]).
    msdf given = output_watertight_template=False (reference :260-263): tets without a positive mSDF corner are dropped too."""
    inside = sdf > 0
    corner_in = inside[tets]                                  # [T,4]
    n_in = corner_in.sum(-1)
    valid = (n_in > 0) & (n_in < 4)
    if msdf is not None:
        valid = valid & ((msdf > 0)[tets].sum(-1) > 0)
    ends = tets[valid][:, _t(TET_EDGE_ENDS)].reshape(-1, 2)
    lo = torch.minimum(ends[:, 0], ends[:, 1])
    hi = torch.maximum(ends[:, 0], ends[:, 1])
    if unique_mode == "rows":       # what the reference executes: row-wise unique (gshell_tets.py:268)
        uniq, inverse = torch.unique(torch.stack([lo, hi], -1), dim=0, return_inverse=True)
    else:                           # same order (lexicographic), via packed 64-bit keys
        nv = int(sdf.shape[0])
        key, inverse = torch.unique(lo * nv + hi, return_inverse=True)
        uniq = torch.stack([key // nv, key % nv], -1)
    crosses = inside[uniq].sum(-1) == 1
    vert_of_edge = torch.full((uniq.shape[0],), -1, dtype=torch.long)
    vert_of_edge[crosses] = torch.arange(int(crosses.sum()))
    case = (corner_in[valid].long() * _t([1, 2, 4, 8])).sum(-1)
    return valid, case, vert_of_edge[inverse].reshape(-1, 6), uniq[crosses]


def lerp_on_sdf(pos, sdf, msdf, edge_lo_hi):
    """Zero crossing of the SDF on each edge; same weights applied to mSDF (with / without grad)."""
    p = pos[edge_lo_hi]                                        # [Vw,2,3]
    s = sdf[edge_lo_hi].unsqueeze(-1) * torch.tensor([1.0, -1.0]).view(1, 2, 1)   # (s_lo, -s_hi)
    den = s.sum(1, keepdim=True)
    den = torch.sign(den) * (den.abs() + 1e-12)
    den = torch.where(den == 0, torch.full_like(den, 1e-12), den)
    w = torch.flip(s, [1]) / den                               # (-s_hi/den, s_lo/den)
    verts = (p * w).sum(1)
    m = msdf[edge_lo_hi]
    m_vert = (m * w.squeeze(-1)).sum(1)
    m_vert_sg = (m * w.squeeze(-1).detach()).sum(1)
    return verts, m_vert, m_vert_sg


def watertight_faces(case, vmap):
    one = _t(N_TRI)[case] == 1
    two = _t(N_TRI)[case] == 2
    tri = _t(TRI_TABLE)
    f1 = torch.gather(vmap[one], 1, tri[case[one]][:, :3]).reshape(-1, 3)
    f2 = torch.gather(vmap[two], 1, tri[case[two]][:, :6]).reshape(-1, 3)
    return torch.cat([f1, f2], 0), one, two


def _dot(a, b):
    return (a * b).sum(-1, keepdim=True)


def _unit(x, eps=1e-20):
    return x / torch.sqrt(torch.clamp(_dot(x, x), min=eps))


def smooth_normals(v, f):
    """gshell_tets.py:9-34."""
    p0, p1, p2 = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]
    fn = torch.linalg.cross(p1 - p0, p2 - p0)
    n = torch.zeros_like(v)
    for c in range(3):
        n = n.index_add(0, f[:, c], fn)
    n = torch.where(_dot(n, n) > 1e-20, n, torch.tensor([0.0, 0.0, 1.0]))
    return _unit(n)


def tangent_frame(verts, faces, valid, one, two, n_tets):
    """uv atlas (map_uv :210-239) + MikkT-style tangents (compute_tangents :40-78)."""
    tet_id = torch.arange(n_tets)[valid]
    gid = torch.cat([tet_id[one] * 2, torch.stack([tet_id[two] * 2, tet_id[two] * 2 + 1], -1).view(-1)])
    n = int(math.ceil(math.sqrt((2 * n_tets + 1) // 2)))
    lin = torch.linspace(0, 1 - (1 / n), n, dtype=torch.float32)
    ty, tx = torch.meshgrid(lin, lin, indexing="ij")
    pad = 0.9 / n
    uvs = torch.stack([tx, ty, tx + pad, ty, tx + pad, ty + pad, tx, ty + pad], -1).view(-1, 2)
    # NB the reference builds per-face uv indices (uv_idx_pre, :232-237) but then calls
    # compute_tangents(verts, uvs_pre, v_nrm, faces, faces, faces) (:319): the uv table is indexed
    # with the *vertex* ids of each face, not with uv_idx_pre.  That is the behaviour restated here.
    del gid
    nrm = smooth_normals(verts, faces)
    p = [verts[faces[:, i]] for i in range(3)]
    t = [uvs[faces[:, i]] for i in range(3)]
    du1, du2 = t[1] - t[0], t[2] - t[0]
    dp1, dp2 = p[1] - p[0], p[2] - p[0]
    nom = dp1 * du2[:, 1:2] - dp2 * du1[:, 1:2]
    den = du1[:, 0:1] * du2[:, 1:2] - du1[:, 1:2] * du2[:, 0:1]
    tang = nom / torch.where(den > 0, torch.clamp(den, min=1e-6), torch.clamp(den, max=-1e-6))
    acc = torch.zeros_like(nrm)
    cnt = torch.zeros_like(nrm)
    for i in range(3):
        acc = acc.index_add(0, faces[:, i], tang)
        cnt = cnt.index_add(0, faces[:, i], torch.ones_like(tang))
    tng = _unit(acc / cnt)
    return _unit(tng - _dot(tng, nrm) * nrm)


def tangent_conditioning(verts, faces, n_tets):
    """Test helper (not part of the restatement): per watertight vertex |sum x_f| / sum scale_f over its faces f, for the face
    normals (scale = product of the two edge lengths) and for the per-face tangents (scale = size of the two terms of the
    numerator over the denominator).  A ratio near 0 means the summands cancel: the normalised sum -- the vertex normal the
    tangent is orthogonalised against, or the tangent itself -- is then rounding residue that depends on the summation order in
    ANY fp32 implementation (the reference's own result moves by O(1) against its fp64 evaluation on such rows,
    tests/test_oracle_tangent_conditioning.py).  On the fixtures the ratios are either < 1e-7 or > 0.2."""
    f = faces.long()
    n = int(math.ceil(math.sqrt((2 * n_tets + 1) // 2)))
    lin = torch.linspace(0, 1 - (1 / n), n, dtype=torch.float32)
    ty, tx = torch.meshgrid(lin, lin, indexing="ij")
    pad = 0.9 / n
    uvs = torch.stack([tx, ty, tx + pad, ty, tx + pad, ty + pad, tx, ty + pad], -1).view(-1, 2)
    p = [verts[f[:, i]] for i in range(3)]
    t = [uvs[f[:, i]] for i in range(3)]
    du1, du2 = t[1] - t[0], t[2] - t[0]
    dp1, dp2 = p[1] - p[0], p[2] - p[0]
    den = du1[:, 0:1] * du2[:, 1:2] - du1[:, 1:2] * du2[:, 0:1]
    denc = torch.where(den > 0, torch.clamp(den, min=1e-6), torch.clamp(den, max=-1e-6))
    tang = (dp1 * du2[:, 1:2] - dp2 * du1[:, 1:2]) / denc
    l1, l2 = dp1.norm(dim=-1), dp2.norm(dim=-1)
    # scale of the terms each face contributes BEFORE any cancellation inside the face: a zero-area face (two coinciding
    # vertices) has a normal of pure rounding residue although it is the only face of its vertex
    scale_n = l1 * l2
    scale_t = (l1 * du2[:, 1].abs() + l2 * du1[:, 1].abs()) / denc[:, 0].abs()
    ratios = []
    for x, scale in ((torch.linalg.cross(dp1, dp2), scale_n), (tang, scale_t)):
        vec, mag = torch.zeros_like(verts), torch.zeros(verts.shape[0])
        for i in range(3):
            vec = vec.index_add(0, f[:, i], x)
            mag = mag.index_add(0, f[:, i], scale)
        ratios.append(vec.norm(dim=-1) / mag.clamp(min=1e-30))
    return ratios[0], ratios[1]


def determined_tangent_rows(verts, faces, tri_loop, quad_loop, n_tets, thresh=1e-4):
    """bool [Va]: rows of v_tng_aug whose value is determined by the inputs rather than by the summation order -- watertight rows
    whose normal sum does not cancel and whose tangent sum is either exactly zero (-> the zero vector everywhere) or does not
    cancel; boundary rows whose two end points are such rows."""
    rn, rt = tangent_conditioning(verts.detach(), faces, n_tets)
    ok = (rn > thresh) & ((rt == 0) | (rt > thresh))
    parts = [ok]
    for loop in (tri_loop, quad_loop):
        parts.append((ok[loop[:, :, 0]] & ok[loop[:, :, 1]]).reshape(-1))
    return torch.cat(parts)


def polygon_loops(case, vmap, one, two):
    loop = _t(LOOP_TABLE)
    tri = torch.gather(vmap[one], 1, loop[case[one]][:, _t([0, 1, 1, 2, 2, 0])]).view(-1, 3, 2)
    quad = torch.gather(vmap[two], 1, loop[case[two]][:, _t([0, 1, 1, 2, 2, 3, 3, 0])]).view(-1, 4, 2)
    return tri, quad


def boundary_weights(m_pair):
    """m_pair [...,2] = interpolated mSDF at the two ends of a polygon edge -> lerp weights [...,2]."""
    ma, mb = m_pair[..., 0], m_pair[..., 1]
    straddles = (torch.sign(ma) + torch.sign(mb)).abs() != 2
    den = ma + (-mb)
    ok = straddles & (den.abs() > 1e-12)
    safe = torch.where(ok, den, torch.ones_like(den))
    w = torch.stack([-mb / safe, ma / safe], -1)
    return torch.where(ok.unsqueeze(-1), w, torch.zeros_like(w))


def cut_faces(m_vert, tri_loop, quad_loop, n_wt):
    occ3 = (m_vert[tri_loop[:, :, 0]] > 0).long()
    occ4 = (m_vert[quad_loop[:, :, 0]] > 0).long()
    code3 = (occ3 * _t([4, 2, 1])).sum(-1)
    code4 = (occ4 * _t([8, 4, 2, 1])).sum(-1)
    nt, nq = tri_loop.shape[0], quad_loop.shape[0]
    ids3 = torch.cat([tri_loop[:, :, 0], n_wt + torch.arange(nt * 3).view(-1, 3)], -1)
    ids4 = torch.cat([quad_loop[:, :, 0], n_wt + nt * 3 + torch.arange(nq * 4).view(-1, 4)], -1)
    groups = []
    for ids, code, table, counts, kmax in ((ids3, code3, _t(CUT_TRI), _t(N_CUT_TRI), 2),
                                           (ids4, code4, _t(CUT_QUAD), _t(N_CUT_QUAD), 4)):
        for k in range(1, kmax + 1):
            sel = counts[code] == k
            groups.append(torch.gather(ids[sel], 1, table[code[sel]][:, :3 * k]).view(-1, 3))
    return torch.cat(groups, 0)


def gshell_marching_tets(pos, sdf, msdf, tets, unique_mode="rows", with_tangents=True, output_watertight_template=True):
    """Same contract as reference `GShell_Tets.__call__(pos_nx3, sdf_n, msdf_n, tet_fx4, output_watertight_template)`.

    Returns (verts_aug, faces_aug, None, None, v_tng_aug, extra) — see gshell_tets.py:426-443.
    """
    sdf = sdf.float().reshape(-1)
    msdf = msdf.reshape(-1)
    with torch.no_grad():
        valid, case, vmap, edge_lo_hi = crossing_edges(sdf, tets, unique_mode, None if output_watertight_template else msdf)
    verts, m_vert, m_vert_sg = lerp_on_sdf(pos, sdf, msdf, edge_lo_hi)
    n_wt = verts.shape[0]
    with torch.no_grad():
        faces, one, two = watertight_faces(case, vmap)
        tri_loop, quad_loop = polygon_loops(case, vmap, one, two)
    v_tng = tangent_frame(verts, faces, valid, one, two, tets.shape[0]) if with_tangents else None

    parts_v, parts_t, parts_m = [verts], [v_tng], [m_vert_sg]
    for loop in (tri_loop, quad_loop):
        w = boundary_weights(m_vert[loop])                     # [P,k,2], grads flow to msdf and sdf
        parts_v.append((verts[loop] * w.unsqueeze(-1)).sum(2).reshape(-1, 3))
        if with_tangents:
            parts_t.append((v_tng[loop] * w.unsqueeze(-1)).sum(2).reshape(-1, 3))
        parts_m.append((m_vert_sg[loop] * w.detach()).sum(2).reshape(-1))
    verts_aug = torch.cat(parts_v, 0)
    v_tng_aug = torch.cat(parts_t, 0) if with_tangents else None
    m_aug_sg = torch.cat(parts_m, 0)

    with torch.no_grad():
        faces_aug = cut_faces(m_vert, tri_loop, quad_loop, n_wt)
        used = torch.zeros(verts_aug.shape[0], dtype=torch.bool)
        used[faces_aug.reshape(-1)] = True
    verts_aug = torch.where(used.unsqueeze(-1), verts_aug, torch.zeros_like(verts_aug))

    extra = {
        "n_verts_watertight": n_wt,
        "vertices_watertight": verts,
        "faces_watertight": faces,
        "v_tng_watertight": v_tng,
        "msdf": m_aug_sg,
        "msdf_watertight": m_vert_sg,
        "msdf_boundary": m_aug_sg[n_wt:],
    }
    if not output_watertight_template:            # reference :435-441: only the mSDF entries
        extra = {k: extra[k] for k in ("msdf", "msdf_watertight", "msdf_boundary")}
    return verts_aug, faces_aug, None, None, v_tng_aug, extra


def gshell_marching_from_auggrid(pos, sdf, tets, sorted_tet_edges, coeff_grid, verts_discretized, msdf_sign_grid, occgrid,
                                 with_tangents=True):
    """Same contract as reference `GShell_Tets.marching_from_auggrid` (gshell_tets.py:446-629, the decode path of generated
    grids; runs under no_grad there).  Differences to `__call__`: crossing vertices sit at `coeff_grid[midpoint]` along their
    edge (:479-482), their mSDF is `msdf_sign_grid[midpoint]` (:484), and EVERY polygon edge gets a boundary vertex whose
    weight comes from `occgrid` at the edge's canonical midpoint, ordered by the lexicographic sign of the canonical edge
    direction (:543-575).  Returns the reference's 9-tuple."""
    with torch.no_grad():
        sdf = sdf.float().reshape(-1)
        n_tets = tets.shape[0]
        inside = sdf > 0
        corner_in = inside[tets]
        n_in = corner_in.sum(-1)
        valid = (n_in > 0) & (n_in < 4)
        ends = sorted_tet_edges.reshape(-1, 6, 2)[valid].reshape(-1, 1, 2)
        uniq, inverse = torch.unique(ends, dim=0, return_inverse=True)                   # :467
        uniq = uniq.long().reshape(-1, 2)
        crosses = inside[uniq].sum(-1) == 1
        vert_of_edge = torch.full((uniq.shape[0],), -1, dtype=torch.long)
        vert_of_edge[crosses] = torch.arange(int(crosses.sum()))
        vmap = vert_of_edge[inverse].reshape(-1, 6)
        edge = uniq[crosses]
        case = (corner_in[valid].long() * _t([1, 2, 4, 8])).sum(-1)

        p = pos[edge]
        cano = verts_discretized[edge].float()
        verts_cano = (cano[:, 0] + cano[:, 1]) / 2.0
        mid = cano.mean(1).long()
        c = coeff_grid[mid[:, 0], mid[:, 1], mid[:, 2]].view(-1, 1).clamp(0, 1)
        verts = p[:, 1] * c + p[:, 0] * (1 - c)
        m_vert = msdf_sign_grid[mid[:, 0], mid[:, 1], mid[:, 2]]
        n_wt = verts.shape[0]

        faces, one, two = watertight_faces(case, vmap)
        tet_gidx = torch.arange(n_tets)[valid]
        valid_tet_gidx = torch.cat([tet_gidx[one], tet_gidx[two]])
        v_tng = tangent_frame(verts, faces, valid, one, two, n_tets) if with_tangents else None
        tri_loop, quad_loop = polygon_loops(case, vmap, one, two)

        parts_v, parts_t = [verts], [v_tng]
        for loop in (tri_loop, quad_loop):                                              # [P,k,2] watertight vertex ids
            e_cano = verts_cano[loop]                                                   # [P,k,2,3]
            loc = (e_cano.mean(2) * 2.0).long()
            co = occgrid[loc[..., 0], loc[..., 1], loc[..., 2]] * 0.5 + 0.5
            co = torch.stack([co, 1 - co], -1)
            k = (torch.sign(e_cano[:, :, 0] - e_cano[:, :, 1]) * torch.tensor([16.0, 4.0, 1.0])).sum(-1)
            order = torch.stack([k, -k], -1).sort(-1, descending=True)[1]
            w = torch.gather(co, -1, order).unsqueeze(-1)                               # [P,k,2,1]
            parts_v.append((verts[loop] * w).sum(2).reshape(-1, 3))
            if with_tangents:
                parts_t.append((v_tng[loop] * w).sum(2).reshape(-1, 3))
        verts_aug = torch.cat(parts_v, 0)
        v_tng_aug = torch.cat(parts_t, 0) if with_tangents else None
        m_aug = torch.cat([m_vert, torch.zeros(verts_aug.shape[0] - n_wt)])
        faces_aug = cut_faces(m_vert, tri_loop, quad_loop, n_wt)
    return verts_aug, faces_aug, None, None, v_tng_aug, verts, valid_tet_gidx, m_aug, m_vert
