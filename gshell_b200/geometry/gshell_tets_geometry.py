"""`GShellTetsGeometry`: parameters, `getMesh()`, `render()`, `tick()` -- same surface as the reference's
geometry/gshell_tets_geometry.py (:45-384), on this repo's CUDA operators.

Differences: faces are kept int32 end to end (the reference widens to int64 and narrows again at every
consumer); `all_edges` for the SDF regulariser reuses the static sorted edge table of the extraction; kaolin's
`sample_points` (Eikonal term, only with use_sdf_mlp) is a small torch function here; `FLAGS` may be any object
with the attributes read below (defaults filled by `default_flags`).
"""
import types

import numpy as np
import torch

from .. import losses, timing
from ..render import mesh, render
from ..render import optixutils as ou
from .gshell_tets import GShell_Tets
from .tet_tables import tables_for
from ..distributed import allreduce_or_mask_


def default_flags(**kw):
    """Hot-path flags with the reference's defaults (train_gshelltet_deepfashion.py:539-594)."""
    d = dict(boxscale=[1.0, 1.0, 1.0], use_sdf_mlp=False, use_msdf_mlp=False, sphere_init=False, sphere_init_norm=0.5,
             use_tanh_deform=False, visualize_watertight=False, n_samples=8, decorrelated=False, denoiser_demodulate=True,
             use_img_2nd_layer=False, use_depth=False, use_eikonal=False, eikonal_scale=None, use_mesh_msdf_reg=True,
             msdf_reg_open_scale=1e-6, msdf_reg_close_scale=3e-6, sdf_regularizer=0.2, lambda_diffuse=0.15,
             lambda_specular=0.0025, lambda_kd=0.1, lambda_ks=0.05, lambda_nrm=0.025, lambda_chroma=0.0, iter=5000,
             skip_in=[3], n_freq=6, n_hidden=6, d_hidden=256, use_float16=False, sdf_mlp_pretrain_steps=500)
    d.update(kw)
    return types.SimpleNamespace(**d)


def compute_sdf_reg_loss(sdf, all_edges):
    """Reference :33-39: BCE between the SDF values at the two ends of every sign-changing grid edge -- one reduction kernel over
    the static edge table and one adjoint kernel (gshell_b200/losses.py)."""
    return losses.sdf_reg_loss(sdf, all_edges if all_edges.dtype == torch.int32 else all_edges.int().contiguous())


def sample_points(v_pos, faces, n):
    """Area-weighted surface samples (stands in for kaolin.ops.mesh.sample_points, reference :236)."""
    f = faces.long()
    p0, p1, p2 = v_pos[f[:, 0]], v_pos[f[:, 1]], v_pos[f[:, 2]]
    area = torch.linalg.cross(p1 - p0, p2 - p0).norm(dim=-1)
    idx = torch.multinomial(area.clamp(min=1e-20), n, replacement=True)
    u = torch.rand(n, 2, device=v_pos.device)
    su = u[:, 0:1].sqrt()
    return (1 - su) * p0[idx] + su * (1 - u[:, 1:2]) * p1[idx] + su * u[:, 1:2] * p2[idx]


class GShellTetsGeometry(torch.nn.Module):
    def __init__(self, grid_res, scale, FLAGS, offset=None, tet_init_file=None, extract_from_generative=False, device="cuda"):
        super().__init__()
        self.FLAGS = FLAGS
        self.grid_res = grid_res
        self.gshell_tets = GShell_Tets(index_dtype=torch.int32, with_tangents=False)   # getMesh discards v_tng (reference :206-214)
        self.scale = scale
        self.boxscale = torch.tensor(FLAGS.boxscale, dtype=torch.float32).view(1, 3).to(device)
        with torch.no_grad():
            self.optix_ctx = ou.OptiXContext()
            tets = np.load("data/tets/{}_tets.npz".format(grid_res) if tet_init_file is None else tet_init_file)
            self.verts = torch.tensor(tets["vertices"], dtype=torch.float32, device=device)
            self.verts = self.verts - self.verts.mean(dim=0)
            self.verts = self.verts * scale * self.boxscale
            self.indices = torch.tensor(tets["indices"], dtype=torch.long, device=device)
            self.generate_edges()
            self.original_verts = None
            if extract_from_generative:
                # reference :64,70-78: the lattice on which generated grids store their per-edge / per-tet features
                raw = torch.tensor(tets["vertices"], dtype=torch.float32, device=device)
                self.original_verts = raw.clone()
                if "tet_edges" in tets:
                    self.sorted_tetedges = torch.tensor(tets["tet_edges"], dtype=torch.long, device=device)
                else:               # the synthetic grids of gshell_b200.grids carry no edge table: same content, built here
                    from .tet_tables import TET_EDGE_ENDS
                    ends = self.indices[:, list(TET_EDGE_ENDS)].reshape(-1, 6, 2)
                    self.sorted_tetedges = torch.sort(ends, dim=-1)[0]
                uniq = raw.view(-1).unique()
                dx = (uniq[1] - uniq[0]) / 2.0
                self.verts_discretized = ((raw - raw.min()) / dx).long().float()
            self.offset = 0.0 if offset is None else torch.tensor(offset, dtype=torch.float32, device=device).view(1, 3)

        if FLAGS.use_sdf_mlp:
            from .mlp import MLP
            self.sdf = torch.nn.Parameter(torch.zeros_like(self.verts[:, 0]), requires_grad=True)
            self.sdf_net = MLP(skip_in=FLAGS.skip_in, n_freq=FLAGS.n_freq, n_hidden=FLAGS.n_hidden, d_hidden=FLAGS.d_hidden,
                               use_float16=FLAGS.use_float16).to(device)
            opt = torch.optim.Adam(self.sdf_net.parameters(), lr=1e-3)
            for _ in range(FLAGS.sdf_mlp_pretrain_steps):
                target = (self.verts / self.boxscale).norm(dim=1, keepdim=True) - FLAGS.sphere_init_norm
                loss = (self.sdf_net(self.verts) - target).pow(2).mean()
                opt.zero_grad()
                loss.backward()
                opt.step()
        else:
            if not FLAGS.sphere_init:
                sdf = torch.rand_like(self.verts[:, 0]) - 0.1
            else:
                sdf = (self.verts / self.boxscale).norm(dim=1) - 0.5
            self.sdf = torch.nn.Parameter(sdf.clone().detach(), requires_grad=True)
        if FLAGS.use_msdf_mlp:
            # reference :118-136: a placeholder parameter plus a field MLP fitted to the constant 0.1 (plain PyTorch, as the SDF field)
            from .mlp import MLP
            self.msdf = torch.nn.Parameter(torch.zeros_like(self.verts[:, 0]), requires_grad=True)
            self.msdf_net = MLP(skip_in=FLAGS.skip_in, n_freq=FLAGS.n_freq, n_hidden=FLAGS.n_hidden, d_hidden=FLAGS.d_hidden,
                                use_float16=FLAGS.use_float16).to(device)
            opt = torch.optim.Adam(self.msdf_net.parameters(), lr=1e-3)
            for _ in range(100):
                loss = (self.msdf_net(self.verts) - 0.1).pow(2).mean()
                opt.zero_grad()
                loss.backward()
                opt.step()
        else:
            msdf = (torch.rand_like(self.verts[:, 0]) - 0.01).clamp(-1, 1)
            self.msdf = torch.nn.Parameter(msdf.clone().detach(), requires_grad=True)
        self.deform = torch.nn.Parameter(torch.zeros_like(self.verts), requires_grad=True)
        self.clamp_deform()

    @torch.no_grad()
    def generate_edges(self):
        # the extraction's static table already holds the sorted unique (lo,hi) grid edges the reference rebuilds
        # with sort + unique(dim=0) (:149-155)
        self.all_edges = tables_for(self.indices, self.verts.shape[0]).edge_v
        self.max_displacement = 1.0 / self.grid_res * self.scale / 2.1

    def getMesh_from_augmented_grid_withocc(self, material, sdf_sign, sdf_coeff, msdf_sign, occgrid):
        """Reference :166-189: decode one generated augmented grid into a mesh (no BVH build, no gradients)."""
        if self.original_verts is None:
            raise RuntimeError("construct the geometry with extract_from_generative=True")
        v_deformed = self.verts + self.max_displacement * self.deform
        sdf = self.sdf_net(v_deformed) if self.FLAGS.use_sdf_mlp else self.sdf
        verts, faces, _, _, v_tng, _, _, v_msdf, _ = GShell_Tets(index_dtype=torch.int32, with_tangents=True).marching_from_auggrid(
            v_deformed, sdf_sign, self.indices, self.sorted_tetedges, sdf_coeff, self.verts_discretized, msdf_sign, occgrid)
        imesh = mesh.auto_normals(mesh.Mesh(verts, faces, material=material))
        imesh = mesh.compute_tangents(imesh, v_tng=v_tng)
        return {"imesh": imesh, "sdf": sdf, "v_msdf": v_msdf}

    @torch.no_grad()
    def getAABB(self):
        return torch.min(self.verts, dim=0).values, torch.max(self.verts, dim=0).values

    @torch.no_grad()
    def clamp_deform(self):
        if not self.FLAGS.use_tanh_deform:
            self.deform.data[:] = self.deform.clamp(-1.0, 1.0)
        self.msdf.data[:] = self.msdf.clamp(-2.0, 2.0)

    def getMesh(self, material):
        v_deformed = self.verts + self.max_displacement * self.deform
        sdf = self.sdf_net(v_deformed) if self.FLAGS.use_sdf_mlp else self.sdf
        msdf = self.msdf_net(v_deformed) if self.FLAGS.use_msdf_mlp else self.msdf      # reference :199-202
        v_deformed = v_deformed + self.offset
        with timing.stage("extraction"):
            verts, faces, uvs, uv_idx, v_tng, extra = self.gshell_tets(v_deformed, sdf, msdf, self.indices)
        imesh = mesh.Mesh(verts, faces, v_tex=uvs, t_tex_idx=uv_idx, material=material)
        with torch.no_grad(), timing.stage("occluder_build"):
            ou.optix_build_bvh(self.optix_ctx, imesh.v_pos.contiguous(), imesh.t_pos_idx.int(), rebuild=1)
        with timing.stage("vertex_normals"):
            imesh = mesh.auto_normals(imesh)
        out = {"imesh": imesh, "sdf": sdf, "msdf": extra["msdf"], "msdf_watertight": extra["msdf_watertight"],
               "msdf_boundary": extra["msdf_boundary"], "n_verts_watertight": extra["n_verts_watertight"]}
        if self.FLAGS.visualize_watertight:
            wt = mesh.Mesh(extra["vertices_watertight"], extra["faces_watertight"], material=material)
            out["imesh_watertight"] = mesh.auto_normals(wt)
        return out

    def render(self, glctx, target, lgt, opt_material, bsdf=None, denoiser=None, shadow_scale=1.0, use_uv=False):
        d = self.getMesh(opt_material)
        opt_mesh = d["imesh"]
        if opt_mesh.v_pos.size(0) != 0 and self.FLAGS.use_sdf_mlp and self.FLAGS.use_eikonal:
            d["sampled_pts"] = sample_points(opt_mesh.v_pos, opt_mesh.t_pos_idx, 50000)
        else:
            d["sampled_pts"] = None
        extra_dict = {"msdf": d["msdf"]}
        d["buffers"] = render.render_mesh(self.FLAGS, glctx, opt_mesh, target["mvp"], target["campos"], lgt,
                                          target["resolution"], spp=target["spp"], msaa=True, background=target["background"],
                                          bsdf=bsdf, use_uv=use_uv, optix_ctx=self.optix_ctx, denoiser=denoiser,
                                          shadow_scale=shadow_scale, extra_dict=extra_dict)
        if self.FLAGS.visualize_watertight:
            d["buffers_watertight"] = render.render_mesh(self.FLAGS, glctx, d["imesh_watertight"], target["mvp"],
                                                         target["campos"], lgt, target["resolution"], spp=target["spp"],
                                                         msaa=True, background=target["background"], bsdf=bsdf, use_uv=use_uv,
                                                         optix_ctx=self.optix_ctx, denoiser=denoiser,
                                                         shadow_scale=shadow_scale, extra_dict=extra_dict)
        return d

    def tick(self, glctx, target, lgt, opt_material, loss_fn, iteration, denoiser):
        """Reference :257-384: render + loss assembly.  Returns (img_loss, depth_loss, reg_loss)."""
        FLAGS = self.FLAGS
        t_iter = iteration / FLAGS.iter
        shadow_ramp = min(iteration / 1000, 1.0)
        if denoiser is not None:
            denoiser.set_influence(shadow_ramp)
        d = self.render(glctx, target, lgt, opt_material, denoiser=denoiser, shadow_scale=shadow_ramp)
        buffers = d["buffers"]
        with torch.no_grad():
            color_ref = target["img"]
        # Every image-space term in ONE pass over the composited buffers (reference :283-290 alpha MSE + the two mSDF-image L1
        # terms; render/regularizer.py shading / material-smoothness / chroma terms): csrc/tick_ops.cu.  The reference's default
        # flags never reach the unsupported second-layer / depth terms; ask for them and tick() says so.
        for flag in ("use_img_2nd_layer", "use_depth", "use_depth_2nd_layer"):
            if getattr(FLAGS, flag, False):
                raise NotImplementedError(f"FLAGS.{flag}: second-layer / depth losses are outside the single-layer hot path")
        t_losses = timing.stage("losses")
        t_losses.__enter__()
        have_light = "diffuse_light" in buffers
        terms = losses.T_ALPHA | losses.T_MSDF | losses.T_SMOOTH | (losses.T_SHADING if have_light else 0) | \
            (losses.T_CHROMA if FLAGS.lambda_chroma != 0 else 0)
        it = losses.image_terms(color_ref, terms,
                                (FLAGS.lambda_chroma, FLAGS.lambda_diffuse, FLAGS.lambda_specular, FLAGS.lambda_kd, FLAGS.lambda_ks,
                                 FLAGS.lambda_nrm),
                                shaded=buffers["shaded"], msdf_img=buffers["msdf_image"], kd=buffers["kd"], kd_grad=buffers["kd_grad"],
                                ks_grad=buffers["ks_grad"], nrm_grad=buffers["normal_grad"],
                                diffuse=buffers["diffuse_light"] if have_light else None,
                                specular=buffers["specular_light"] if have_light else None)
        img_loss = it[0] + loss_fn(buffers["shaded"][..., 0:3] * color_ref[..., 3:], color_ref[..., 0:3] * color_ref[..., 3:])
        depth_loss = torch.zeros((), device=img_loss.device)

        eik_loss = torch.zeros((), device=img_loss.device)
        if FLAGS.use_sdf_mlp and FLAGS.use_eikonal and d["sampled_pts"] is not None:
            v = d["sampled_pts"].detach().requires_grad_(True)
            sdf_eik = self.sdf_net(v)
            if FLAGS.eikonal_scale is None:
                eik_coeff = 3e-1 if iteration < 500 else (1e-1 if iteration < 2000 else 1e-2)
            else:
                eik_coeff = FLAGS.eikonal_scale
            gnorm = torch.autograd.grad(sdf_eik.sum(), v, create_graph=True)[0].pow(2).sum(dim=-1).sqrt()
            eik_loss = eik_coeff * (gnorm - 1).pow(2).mean()

        mesh_msdf_reg_loss = torch.zeros((), device=img_loss.device)
        if FLAGS.use_mesh_msdf_reg and d["msdf"].numel() > 0:
            regscale = (64 / self.grid_res) ** 3
            bmask = None
            if FLAGS.msdf_reg_close_scale != 0:
                # boundary vertices of any visible triangle (reference :343-356 sorts the ids twice with `unique`; a flag scatter
                # gives the same set), OR-ed over the ranks when the views of the batch are sharded
                bmask = losses.visible_boundary_mask(d["imesh"].t_pos_idx, buffers["visible_triangles"], d["n_verts_watertight"],
                                                     d["msdf_boundary"].size(0))
                allreduce_or_mask_(bmask)
            mesh_msdf_reg_loss = losses.msdf_reg_loss(d["msdf"] if FLAGS.msdf_reg_open_scale > 0 else None, d["msdf_boundary"], bmask,
                                                      FLAGS.msdf_reg_open_scale * regscale, FLAGS.msdf_reg_close_scale * regscale)

        sdf_weight = FLAGS.sdf_regularizer - (FLAGS.sdf_regularizer - 0.01) * min(1.0, 4.0 * t_iter)
        sdf_reg_loss = compute_sdf_reg_loss(d["sdf"].reshape(-1), self.all_edges) * sdf_weight
        reg_loss = sdf_reg_loss + eik_loss + mesh_msdf_reg_loss + it[1]
        t_losses.__exit__(None, None, None)
        return img_loss, depth_loss, reg_loss
