"""Deferred renderer: `render_mesh -> render_layer -> shade`, same signatures and buffer dictionary as the
reference's render/render.py (:31-191 shade, :199-317 render_layer, :325-444 render_mesh), on this repo's CUDA
operators:  ru.xfm_points / ru.prepare_shading_normal (gbuffer_ops.cu), raster.rasterize / interpolate
(raster.cu, standing in for nvdiffrast), ou.optix_env_shade (env_shade.cu), the bilateral denoiser
(denoise.cu).  Differences from the reference, all deliberate and listed in DESIGN.md:
  * `ctx` (an nvdiffrast GL/CUDA context there) is accepted and ignored;
  * diffuse and specular light are denoised in one fused pass when a BilateralDenoiser is given;
  * antialias() is the identity (no silhouette-edge gradient yet);
  * the single D2H sync of `visible_triangles.unique()` is kept (the caller indexes with it, tick :344).
"""
import torch

from . import light, raster, util
from . import optixutils as ou
from . import renderutils as ru

rnd_seed = 0


def interpolate(attr, rast, attr_idx, rast_db=None):
    out, d = raster.interpolate(attr.contiguous(), rast, attr_idx, rast_db=rast_db)
    return out, (d if rast_db is not None else None)


# ==============================================================================================
#  pixel shader
# ==============================================================================================
def shade(FLAGS, rast, gb_depth, gb_pos, gb_geometric_normal, gb_normal, gb_tangent, gb_texc, gb_texc_deriv, view_pos,
          lgt, material, optix_ctx, mesh, bsdf, denoiser, shadow_scale, use_uv=True, finetune_normal=True, xfm_lgt=None,
          shade_data=False):
    dev = gb_pos.device
    B, H, W = gb_depth.shape[0], gb_depth.shape[1], gb_depth.shape[2]
    offset = torch.normal(mean=0, std=0.005, size=(B, H, W, 2), device=dev)
    jitter = (util.pixel_grid(W, H, device=dev)[None, ...] + offset).contiguous()

    mask = (rast[..., -1:] > 0).float()
    mask_tap = util.bilinear_tap(mask, jitter)
    grad_weight = mask * mask_tap

    # ---- material lookups (reference :65-101) -------------------------------------------------------
    perturbed_nrm = None
    if "kd_ks" in material:
        all_tex_jitter = material["kd_ks"].sample(gb_pos + torch.normal(mean=0, std=0.01, size=gb_pos.shape, device=dev))
        all_tex = material["kd_ks"].sample(gb_pos)
        assert all_tex.shape[-1] == 6, "Combined kd_ks must be 6 channels"
        kd, ks = all_tex[..., 0:3], all_tex[..., 3:6]
        kd_grad = torch.abs(all_tex_jitter[..., 0:3] - kd)
        ks_grad = torch.abs(all_tex_jitter[..., 3:6] - ks) * torch.tensor([0, 1, 1], dtype=torch.float32, device=dev)[None, None, None, :]
    else:
        raise NotImplementedError("only the combined 'kd_ks' material field of the G-Shell training path is supported")

    alpha = kd[..., 3:4] if kd.shape[-1] == 4 else torch.ones_like(kd[..., 0:1])
    kd = kd[..., 0:3]

    # ---- normal regulariser tap + shading normal (reference :106-118) ---------------------------------
    nrm_jitter = util.bilinear_tap(gb_normal, jitter)
    nrm_grad = torch.abs(nrm_jitter - gb_normal) * grad_weight
    gb_normal = ru.prepare_shading_normal(gb_pos, view_pos, perturbed_nrm, gb_normal, gb_tangent, gb_geometric_normal,
                                          two_sided_shading=True, opengl=True)

    # ---- BSDF (reference :124-162) ----------------------------------------------------------------------
    assert "bsdf" in material or bsdf is not None, "Material must specify a BSDF type"
    bsdf = material["bsdf"] if bsdf is None else bsdf
    diffuse_accum = specular_accum = None
    if bsdf in ("pbr", "diffuse", "white"):
        kd = torch.ones_like(kd) if bsdf == "white" else kd
        assert isinstance(lgt, light.EnvironmentLight) and optix_ctx is not None
        ro = gb_pos + gb_normal * 0.001
        global rnd_seed
        diffuse_accum, specular_accum = ou.optix_env_shade(
            optix_ctx, rast[..., -1], ro, gb_pos, gb_normal, view_pos, kd, ks, lgt.base, lgt._pdf, lgt.rows[:, 0], lgt.cols,
            BSDF=bsdf, n_samples_x=FLAGS.n_samples, rnd_seed=None if FLAGS.decorrelated else rnd_seed,
            shadow_scale=shadow_scale)
        rnd_seed += 1
        if denoiser is not None and FLAGS.denoiser_demodulate:
            if hasattr(denoiser, "forward_pair"):
                diffuse_accum, specular_accum = denoiser.forward_pair(diffuse_accum, specular_accum, gb_normal, gb_depth)
            else:
                diffuse_accum = denoiser.forward(torch.cat((diffuse_accum, gb_normal, gb_depth), dim=-1))
                specular_accum = denoiser.forward(torch.cat((specular_accum, gb_normal, gb_depth), dim=-1))
        if bsdf in ("white", "diffuse"):
            shaded_col = diffuse_accum * kd
        else:
            kd = kd * (1.0 - ks[..., 2:3])
            shaded_col = diffuse_accum * kd + specular_accum
        if denoiser is not None and not FLAGS.denoiser_demodulate:
            shaded_col = denoiser.forward(torch.cat((shaded_col, gb_normal, gb_depth), dim=-1))
    elif bsdf == "normal":
        shaded_col = (gb_normal + 1.0) * 0.5
    elif bsdf == "tangent":
        shaded_col = (gb_tangent + 1.0) * 0.5
    elif bsdf == "kd":
        shaded_col = kd
    elif bsdf == "ks":
        shaded_col = ks
    else:
        assert False, "Invalid BSDF '%s'" % bsdf

    buffers = {
        "shaded": torch.cat((shaded_col, alpha), dim=-1),
        "z_grad": torch.cat((gb_depth, torch.zeros_like(alpha), alpha), dim=-1),
        "normal": torch.cat((gb_normal, alpha), dim=-1),
        "geometric_normal": torch.cat((gb_geometric_normal, alpha), dim=-1),
        "kd": torch.cat((kd, alpha), dim=-1),
        "ks": torch.cat((ks, alpha), dim=-1),
        "kd_grad": torch.cat((kd_grad, alpha), dim=-1),
        "ks_grad": torch.cat((ks_grad, alpha), dim=-1),
        "normal_grad": torch.cat((nrm_grad, alpha), dim=-1),
    }
    if diffuse_accum is not None:
        buffers["diffuse_light"] = torch.cat((diffuse_accum, alpha), dim=-1)
    if specular_accum is not None:
        buffers["specular_light"] = torch.cat((specular_accum, alpha), dim=-1)
    return buffers


# ==============================================================================================
#  Render a depth slice of the mesh
# ==============================================================================================
def render_layer(FLAGS, v_pos_clip, rast, rast_deriv, mesh, view_pos, lgt, resolution, spp, msaa, optix_ctx, bsdf,
                 denoiser, shadow_scale, use_uv=True, finetune_normal=True, extra_dict=None, xfm_lgt=None, shade_data=False):
    full_res = [resolution[0] * spp, resolution[1] * spp]
    if spp > 1 and msaa:
        rast_out_s = util.scale_img_nhwc(rast, resolution, mag="nearest", min="nearest")
        rast_out_deriv_s = util.scale_img_nhwc(rast_deriv, resolution, mag="nearest", min="nearest") * spp
    else:
        rast_out_s, rast_out_deriv_s = rast, rast_deriv
    tri = mesh.t_pos_idx.int()

    gb_pos, _ = interpolate(mesh.v_pos[None, ...], rast_out_s, tri)

    # geometric (face) normals, constant per triangle (reference :243-248)
    v0 = mesh.v_pos[mesh.t_pos_idx[:, 0].long(), :]
    v1 = mesh.v_pos[mesh.t_pos_idx[:, 1].long(), :]
    v2 = mesh.v_pos[mesh.t_pos_idx[:, 2].long(), :]
    face_normals = util.safe_normalize(torch.linalg.cross(v1 - v0, v2 - v0))
    face_normal_indices = torch.arange(0, face_normals.shape[0], dtype=torch.int32, device=face_normals.device)[:, None].repeat(1, 3)
    gb_geometric_normal, _ = interpolate(face_normals[None, ...], rast_out_s, face_normal_indices)

    if use_uv:
        raise NotImplementedError("use_uv=True (texture-space materials) is outside the G-Shell training path")
    assert mesh.v_nrm is not None
    gb_normal, _ = interpolate(mesh.v_nrm[None, ...], rast_out_s, mesh.t_nrm_idx.int())
    with torch.no_grad():
        noise = torch.randn_like(gb_normal)
        noise = noise / noise.norm(dim=-1, keepdim=True)
    gb_tangent = torch.linalg.cross(noise, gb_normal)       # only used to add isotropic noise (reference :264-267)
    gb_texc, gb_texc_deriv = None, None

    with torch.no_grad():
        eps = 0.00001
        clip_pos, clip_pos_deriv = interpolate(v_pos_clip, rast_out_s, tri, rast_db=rast_out_deriv_s)
        z0 = torch.clamp(clip_pos[..., 2:3], min=eps) / torch.clamp(clip_pos[..., 3:4], min=eps)
        z1 = torch.clamp(clip_pos[..., 2:3] + torch.abs(clip_pos_deriv[..., 2:3]), min=eps) / \
            torch.clamp(clip_pos[..., 3:4] + torch.abs(clip_pos_deriv[..., 3:4]), min=eps)
        z_grad = torch.abs(z1 - z0)
        gb_depth = torch.cat((z0, z_grad), dim=-1)

    buffers = shade(FLAGS, rast_out_s, gb_depth, gb_pos, gb_geometric_normal, gb_normal, gb_tangent, gb_texc,
                    gb_texc_deriv, view_pos, lgt, mesh.material, optix_ctx, mesh, bsdf, denoiser, shadow_scale,
                    use_uv=use_uv, finetune_normal=finetune_normal, xfm_lgt=xfm_lgt, shade_data=shade_data)

    if extra_dict is not None:
        for key in extra_dict:
            if key == "msdf" and extra_dict[key] is not None:
                assert extra_dict[key].dim() == 1 or (extra_dict[key].dim() == 2 and extra_dict[key].size(1) == 1)
                buffers["msdf_image"], _ = interpolate(extra_dict[key].reshape(-1)[None, :, None], rast_out_s, tri)
            elif key == "msdf_watertight" and extra_dict[key] is not None:
                buffers["msdf_watertight_image"], _ = interpolate(extra_dict[key].reshape(-1)[None, :, None],
                                                                  rast_out_s.detach(), tri)
    if spp > 1 and msaa:
        for key in buffers.keys():
            buffers[key] = util.scale_img_nhwc(buffers[key], full_res, mag="nearest", min="nearest")
    return buffers


# ==============================================================================================
#  Render a mesh (single layer)
# ==============================================================================================
def render_mesh(FLAGS, ctx, mesh, mtx_in, view_pos, lgt, resolution, spp=1, num_layers=1, msaa=False, background=None,
                optix_ctx=None, bsdf=None, denoiser=None, shadow_scale=1.0, use_uv=True, finetune_normal=True,
                extra_dict=None, xfm_lgt=None, shade_data=False):
    dev = mesh.v_pos.device

    def prepare_input_vector(x):
        x = torch.tensor(x, dtype=torch.float32, device=dev) if not torch.is_tensor(x) else x
        return x[:, None, None, :] if len(x.shape) == 2 else x

    def composite_buffer(key, layers, background, antialias):
        accum = background
        for buffers, rast in reversed(layers):
            alpha = (rast[..., -1:] > 0).float() * buffers[key][..., -1:]
            accum = torch.lerp(accum, torch.cat((buffers[key][..., :-1], torch.ones_like(buffers[key][..., -1:])), dim=-1), alpha)
            if antialias:
                accum = raster.antialias(accum.contiguous(), rast, v_pos_clip, mesh.t_pos_idx.int())
        return accum

    full_res = [resolution[0] * spp, resolution[1] * spp]
    mtx_in = torch.tensor(mtx_in, dtype=torch.float32, device=dev) if not torch.is_tensor(mtx_in) else mtx_in
    view_pos = prepare_input_vector(view_pos)

    v_pos_clip = ru.xfm_points(mesh.v_pos[None, ...], mtx_in)
    assert num_layers == 1
    rast, db = raster.rasterize(v_pos_clip, mesh.t_pos_idx.int(), full_res)
    # sorted unique ids of the visible triangles (reference :380-383 sorts 8M pixel ids with unique(); a flag
    # scatter + nonzero yields the same sorted list without the sort)
    with torch.no_grad():
        seen = torch.zeros(mesh.t_pos_idx.shape[0] + 1, dtype=torch.bool, device=dev)
        seen[rast[..., -1].reshape(-1).long()] = True
        visible_triangles = seen[1:].nonzero()[:, 0]
    layers = [(render_layer(FLAGS, v_pos_clip, rast, db, mesh, view_pos, lgt, resolution, spp, msaa, optix_ctx, bsdf,
                            denoiser, shadow_scale, use_uv=use_uv, finetune_normal=finetune_normal, extra_dict=extra_dict,
                            xfm_lgt=xfm_lgt, shade_data=shade_data), rast)]

    if background is not None:
        if spp > 1:
            background = util.scale_img_nhwc(background, full_res, mag="nearest", min="nearest")
        background = torch.cat((background, torch.zeros_like(background[..., 0:1])), dim=-1)
    else:
        background = torch.zeros(1, full_res[0], full_res[1], 4, dtype=torch.float32, device=dev)

    out_buffers = {"visible_triangles": visible_triangles}
    for key in layers[0][0].keys():
        if layers[0][0][key] is None:
            out_buffers[key] = None
            continue
        if key == "shaded":
            accum = composite_buffer(key, layers, background, True)
        else:
            accum = composite_buffer(key, layers, torch.zeros_like(layers[0][0][key]), True)
        out_buffers[key] = util.avg_pool_nhwc(accum, spp) if spp > 1 else accum
    return out_buffers
