"""`GShellFlexiCubesGeometry` -- same surface as the reference's geometry/gshell_flexicubes_geometry.py (:44-364):
parameters (sdf, msdf, deform, per-cube weights [C,21]), `getMesh()`, `render()`, `tick()`.  It shares render() and the
loss assembly with the tet geometry (the reference's two tick() bodies differ only by the FlexiCubes L_dev term, :358-360)."""
import torch

from ..render import mesh
from ..render import optixutils as ou
from ..render import util
from .flex_tables import tables_for
from .gshell_flexicubes import GShellFlexiCubes
from .gshell_tets_geometry import GShellTetsGeometry


class GShellFlexiCubesGeometry(GShellTetsGeometry):
    def __init__(self, grid_res, scale, FLAGS, device="cuda"):
        torch.nn.Module.__init__(self)
        self.FLAGS = FLAGS
        self.grid_res = grid_res
        self.scale = scale
        self.gflexicubes = GShellFlexiCubes(device=device, index_dtype=torch.int32)
        verts, indices = self.gflexicubes.construct_voxel_grid(grid_res)
        self.boxscale = torch.tensor(FLAGS.boxscale, dtype=torch.float32).view(1, 3).to(device)
        with torch.no_grad():
            self.optix_ctx = ou.OptiXContext()
        self.verts = verts * scale * self.boxscale
        self.indices = indices
        self.offset = 0.0
        self.generate_edges()
        if FLAGS.use_sdf_mlp:
            # reference :66-85 (its default for this geometry): the SDF is a coordinate MLP pre-fitted to a sphere; `sdf` stays a
            # placeholder parameter.  The field itself is upstream of the hot path: plain torch / cuBLAS.
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
        self.per_cube_weights = torch.nn.Parameter(torch.ones((indices.shape[0], 21), dtype=torch.float, device=device), requires_grad=True)
        self.register_parameter("weight", self.per_cube_weights)     # second name of the same parameter, as the reference (:97): checkpoint keys
        msdf = (torch.rand_like(self.verts[:, 0]) - 0.01).clamp(-1, 1)
        self.msdf = torch.nn.Parameter(msdf.clone().detach(), requires_grad=True)
        self.deform = torch.nn.Parameter(torch.zeros_like(self.verts), requires_grad=True)
        self.gflexi_reg_loss = torch.zeros((), device=device)
        self.clamp_deform()

    @torch.no_grad()
    def generate_edges(self):
        # unordered unique grid edges (reference :104-110 sorts the pairs and uniques them); the static oriented edge table
        # of the extraction lists every grid edge exactly once already
        tab = tables_for(self.indices, self.verts.shape[0])
        self.all_edges = tab.edge_v
        e = tab.edge_v.long()
        self.max_displacement = util.length(self.verts[e[:, 0]] - self.verts[e[:, 1]]).mean() / 4

    def getMesh(self, material, _training=False):
        v_deformed = self.verts + self.max_displacement * self.deform
        sdf = self.sdf_net(v_deformed).reshape(-1) if self.FLAGS.use_sdf_mlp else self.sdf        # reference :172-175
        msdf = self.msdf
        w = self.per_cube_weights
        # NB the reference never forwards `_training`: the non-training quad split is always used
        verts, faces, reg_loss, extra = self.gflexicubes(v_deformed, sdf, msdf, self.indices, self.grid_res, w[:, :12], w[:, 12:20],
                                                         w[:, 20], training=False)
        self.gflexi_reg_loss = reg_loss.mean() if reg_loss.numel() else torch.zeros((), device=verts.device)
        if extra is None:
            z = torch.zeros((0,), device=verts.device)
            extra = {"msdf": z, "msdf_watertight": z, "msdf_boundary": z, "n_verts_watertight": 0,
                     "vertices_watertight": verts, "faces_watertight": faces}
        imesh = mesh.Mesh(verts, faces, material=material)
        with torch.no_grad():
            ou.optix_build_bvh(self.optix_ctx, imesh.v_pos.contiguous(), imesh.t_pos_idx.int(), rebuild=1)
        imesh = mesh.auto_normals(imesh)
        out = {"imesh": imesh, "sdf": sdf, "msdf": extra["msdf"].reshape(-1), "msdf_watertight": extra["msdf_watertight"].reshape(-1),
               "msdf_boundary": extra["msdf_boundary"].reshape(-1), "n_verts_watertight": extra["n_verts_watertight"]}
        if self.FLAGS.visualize_watertight:
            wt = mesh.Mesh(extra["vertices_watertight"], extra["faces_watertight"], material=material)
            out["imesh_watertight"] = mesh.auto_normals(wt)
        return out

    def tick(self, glctx, target, lgt, opt_material, loss_fn, iteration, denoiser):
        img_loss, depth_loss, reg_loss = super().tick(glctx, target, lgt, opt_material, loss_fn, iteration, denoiser)
        return img_loss, depth_loss, reg_loss + self.gflexi_reg_loss * 0.25          # reference :358-360
