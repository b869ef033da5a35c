"""ORACLE (test infrastructure): brute-force PyTorch restatement of the rasterise / interpolate operators.

PARITY UNPINNED: the reference obtains these from nvdiffrast (third-party, not vendored, unpinned --
`pip install git+https://github.com/NVlabs/nvdiffrast/`, reference README.md:38; call sites
render/render.py:26,240-275,306,377-383), which is absent from this environment, and the reference holds no
test or golden vector for them.  This file restates nvdiffrast's documented output conventions
(rast = (u, v, z/w, triangle_id+1), perspective-correct barycentrics of vertices 0/1, row j <-> NDC y =
(j+0.5)/H*2-1, interpolate = barycentric blend, analytic barycentric gradients) together with the exact
coverage rule of csrc/raster.cu (vertices snapped to 1/256 px, int64 edge functions, top-left ties, nearest
z/w then lowest id), so CUDA-vs-oracle parity is bit-exact on ids and ~1e-6 on floats.
"""
import torch

SUB = 256


def rasterize(clip, tris, H, W):
    """clip [B,V,4] fp32, tris long [F,3] -> rast [B,H,W,4] with autograd on (u,v) w.r.t. clip (x,y,w)."""
    B = clip.shape[0]
    out = []
    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    cx = (xs * SUB + SUB // 2).reshape(-1, 1).to(torch.int64)
    cy = (ys * SUB + SUB // 2).reshape(-1, 1).to(torch.int64)
    for b in range(B):
        c = clip[b]
        p = c[tris]                                             # [F,3,4]
        w = p[..., 3]
        ok = (w > 1e-8).all(-1)
        wsafe = torch.where(w > 1e-8, w, torch.ones_like(w)).detach()
        pd = p.detach()
        sx = (pd[..., 0] / wsafe * 0.5 + 0.5) * float(W) * float(SUB)
        sy = (pd[..., 1] / wsafe * 0.5 + 0.5) * float(H) * float(SUB)
        ok = ok & (sx.abs() < 1e9).all(-1) & (sy.abs() < 1e9).all(-1)
        X = torch.round(sx).to(torch.int64)                     # round-half-even == llrintf
        Y = torch.round(sy).to(torch.int64)
        zw = pd[..., 2] / wsafe
        area = (X[:, 1] - X[:, 0]) * (Y[:, 2] - Y[:, 0]) - (Y[:, 1] - Y[:, 0]) * (X[:, 2] - X[:, 0])
        ok = ok & (area != 0)
        flip = area < 0
        order = torch.where(flip[:, None], torch.tensor([[0, 2, 1]]), torch.tensor([[0, 1, 2]]))
        Xn, Yn = torch.gather(X, 1, order), torch.gather(Y, 1, order)
        zwn, wn = torch.gather(zw, 1, order), torch.gather(wsafe, 1, order)
        an = area.abs()

        def edge(i, j):
            dx, dy = Xn[:, j] - Xn[:, i], Yn[:, j] - Yn[:, i]
            e = dx[None] * (cy - Yn[None, :, i]) - dy[None] * (cx - Xn[None, :, i])      # [P,F]
            own = (dy > 0) | ((dy == 0) & (dx < 0))
            return e, own
        e0, o0 = edge(1, 2)
        e1, o1 = edge(2, 0)
        e2, o2 = edge(0, 1)
        inside = (e0 >= 0) & (e1 >= 0) & (e2 >= 0) & ok[None]
        inside &= ((e0 != 0) | o0[None]) & ((e1 != 0) | o1[None]) & ((e2 != 0) | o2[None])
        inv = 1.0 / an.to(torch.float32)
        b0, b1, b2 = e0.float() * inv, e1.float() * inv, e2.float() * inv
        z = (b0 * zwn[None, :, 0] + b1 * zwn[None, :, 1]) + b2 * zwn[None, :, 2]
        inside &= (z >= -1) & (z <= 1)
        depth_key = (z * 0.5 + 0.5)
        big = torch.where(inside, depth_key, torch.full_like(depth_key, 3.0))
        zmin = big.min(dim=1, keepdim=True).values
        cand = inside & (big == zmin)
        fid = torch.where(cand.any(1), cand.float().argmax(1), torch.full((H * W,), -1))
        hit = fid >= 0
        f = fid.clamp(min=0)
        idx = torch.arange(H * W)
        q0, q1, q2 = b0[idx, f] / wn[f, 0], b1[idx, f] / wn[f, 1], b2[idx, f] / wn[f, 2]
        S = q0 + q1 + q2
        un, vn = q0 / S, q1 / S
        v_snap = torch.where(flip[f], 1.0 - un - vn, vn)
        # differentiable twin: 2-D homogeneous barycentrics (gradient only; value replaced by the snapped one)
        P = torch.stack([p[f][..., 0], p[f][..., 1], p[f][..., 3]], -1)           # [P,3,3] (x,y,w)
        s = torch.stack([(xs.reshape(-1).float() + 0.5) / W * 2 - 1, (ys.reshape(-1).float() + 0.5) / H * 2 - 1,
                         torch.ones(H * W)], -1)
        h0 = (s * torch.linalg.cross(P[:, 1], P[:, 2])).sum(-1)
        h1 = (s * torch.linalg.cross(P[:, 2], P[:, 0])).sum(-1)
        h2 = (s * torch.linalg.cross(P[:, 0], P[:, 1])).sum(-1)
        hs = h0 + h1 + h2
        hs = torch.where(hit, hs, torch.ones_like(hs))
        uh, vh = h0 / hs, h1 / hs
        u = un.detach() + (uh - uh.detach())
        v = v_snap.detach() + (vh - vh.detach())
        zsel = z[idx, f]
        r = torch.stack([u, v, zsel, (f + 1).float()], -1)
        r = torch.where(hit[:, None], r, torch.zeros_like(r))
        out.append(r.view(H, W, 4))
    return torch.stack(out, 0)


def interpolate(attr, rast, tris):
    """attr [1|B,V,C], rast [B,H,W,4], tris long [F,3] -> [B,H,W,C]."""
    B, H, W, _ = rast.shape
    f = rast[..., 3].long() - 1
    hit = f >= 0
    vi = tris[f.clamp(min=0)]                                     # [B,H,W,3]
    a = attr.expand(B, -1, -1)
    bidx = torch.arange(B).view(B, 1, 1, 1).expand(B, H, W, 3)
    corners = a[bidx, vi]                                         # [B,H,W,3,C]
    u, v = rast[..., 0:1], rast[..., 1:2]
    out = corners[..., 0, :] * u + corners[..., 1, :] * v + corners[..., 2, :] * (1.0 - u - v)
    return torch.where(hit[..., None], out, torch.zeros_like(out))
