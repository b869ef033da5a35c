"""ctypes binding of libgshell_b200.so (C ABI declared in include/gshell_b200.h).

There is NO fallback: if the CUDA library is missing or fails to load, importing this module raises.
Build it with `python -m gshell_b200.build` (or `__graft_entry__.build()`).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# GSB_LIB_PATH: profiling builds of the same library (profiles/build_variants.py); there is still no non-CUDA path
LIB_PATH = os.environ.get("GSB_LIB_PATH") or os.path.join(_HERE, "libgshell_b200.so")

_P = ctypes.c_void_p
_I64 = ctypes.c_int64
_I32 = ctypes.c_int
_SZ = ctypes.c_size_t
_F32 = ctypes.c_float
_U32 = ctypes.c_uint32

# name -> (restype, argtypes); must list every symbol declared in include/gshell_b200.h
SIGNATURES = {
    "gsb_abi_version": (_I32, []),
    "gsb_compiled_arch": (_I32, []),
    "gsb_mt_workspace_bytes": (_SZ, [_I64, _I64]),
    "gsb_mt_count": (_I32, [_P, _P, _P, _P, _P, _P, _I64, _I64, _I64, _P, _SZ, _I32, _P, _P]),
    "gsb_mt_emit": (_I32, [_P, _P, _P, _P, _I64, _I64, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "gsb_xfm_points_fwd": (_I32, [_P, _P, _I64, _I64, _I32, _P, _P]),
    "gsb_xfm_points_bwd": (_I32, [_P, _P, _I64, _I64, _I32, _P, _P]),
    "gsb_shading_normal_fwd": (_I32, [_P, _P, _I64, _I64, _I64, _I32, _I32, _P, _P]),
    "gsb_shading_normal_bwd": (_I32, [_P, _P, _I64, _I64, _I64, _I32, _I32, _P, _P, _P]),
    "gsb_image_loss_partials": (_I64, [_I64]),
    "gsb_image_loss_fwd": (_I32, [_P, _P, _I64, _I32, _I32, _P, _P]),
    "gsb_image_loss_bwd": (_I32, [_P, _P, _I64, _I32, _I32, _P, _F32, _P, _P, _P]),
    "gsb_trace_timing": (_F32, [_I32]),
    "gsb_trace_ray_count": (ctypes.c_uint64, [_I32]),
    "gsb_trace_stats": (None, [_P, _I32]),
    "gsb_trace_launches": (_I32, []),
    "gsb_env_shade_dropped_rays": (_U32, [_I32]),
    "gsb_env_shade_scratch_bytes": (_SZ, [_I64, _I64, _I64, _I64, _I32, _SZ]),
    "gsb_env_shade_chunks": (_I32, [_I64, _I64, _I64, _I64, _I32, _SZ]),
    "gsb_trace_shadow_rays": (_I32, [_P, _P, _P, _I64, _P, _P, _P]),
    "gsb_hashgrid_fwd": (_I32, [_P, _I64, _P, _P, _P, _P, _I32, _P, _P]),
    "gsb_hashgrid_bwd": (_I32, [_P, _I64, _P, _P, _P, _P, _I32, _P, _P, _P, _P]),
    "gsb_field_infer": (_I32, [_P, _I64, _P, _P, _P, _P, _I32, _P, _P, _P, _I32, _P, _P, _P, _P]),
    "gsb_env_shade_fwd": (_I32, [_P] * 14 + [_I64] * 6 + [_I32, _I32, _U32, _F32, _P, _P, _SZ, _I64, _P, _P, _P, _P, _P]),
    "gsb_env_shade_bwd": (_I32, [_P] * 14 + [_I64] * 6 + [_I32, _I32, _U32, _F32, _P, _P, _SZ, _I64, _P] + [_P] * 7 + [_P, _P]),
    "gsb_bilateral_fwd": (_I32, [_P, _P, _P, _P, _I64, _I64, _I64, _I64, _I64, _I64, _F32, _P, _P, _P, _P, _P]),
    "gsb_bilateral_bwd": (_I32, [_P, _P, _P, _P, _P, _P, _I64, _I64, _I64, _I64, _I64, _F32, _P, _P, _P, _P]),
    "gsb_rasterize_fwd": (_I32, [_P, _P, _I64, _I64, _I64, _I32, _I64, _I64, _P, _P, _P, _P]),
    "gsb_rasterize_bwd": (_I32, [_P, _P, _P, _P, _I64, _I64, _I32, _I64, _I64, _P, _P]),
    "gsb_interpolate_fwd": (_I32, [_P, _P, _P, _P, _I64, _I64, _I64, _I32, _I64, _I64, _P, _P, _P]),
    "gsb_interpolate_bwd": (_I32, [_P, _P, _P, _P, _I64, _I64, _I64, _I32, _I64, _I64, _P, _P, _P]),
    "gsb_vertex_normals_fwd": (_I32, [_P, _P, _I64, _I64, _P, _P, _P]),
    "gsb_vertex_normals_bwd": (_I32, [_P, _P, _P, _P, _I64, _I64, _P, _P, _P]),
    "gsb_tangents_fwd": (_I32, [_P] * 6 + [_I64] * 4 + [_I32, _F32, _P, _P, _P]),
    "gsb_tangents_bwd": (_I32, [_P] * 5 + [_I64] * 4 + [_I32, _F32] + [_P] * 8),
    "gsb_auggrid_edge_flags": (_I32, [_P, _P, _I64, _P, _P]),
    "gsb_auggrid_vertices": (_I32, [_P] * 5 + [_I64, _P, _P, _I32, _I32, _I32, _P, _P, _P, _P]),
    "gsb_auggrid_classify": (_I32, [_P] * 6 + [_I64, _P, _P, _P]),
    "gsb_auggrid_emit": (_I32, [_P] * 9 + [_I64, _P, _P, _I32, _I32, _I32, _I64, _P] + [_P] * 6 + [_P]),
    "gsb_auggrid_boundary_attr": (_I32, [_P, _P, _P, _I64, _P, _P]),
    "gsb_occluder_struct_bytes": (_SZ, []),
    "gsb_occluder_scan_ws_ints": (_I64, [_I64]),
    "gsb_occluder_brick_words": (_I64, [_I32]),
    "gsb_occluder_cells": (_I64, [_I32]),
    "gsb_occluder_build_count": (_I32, [_P, _P, _I64, _P, _P, _I32, _P, _P, _P, _P, _P, _P]),
    "gsb_occluder_build_fill": (_I32, [_P, _P, _I64, _I32, _P, _P, _P, _P, _P, _P]),
    "gsb_fc_blocks": (_I64, [_I64]),
    "gsb_fc_count": (_I32, [_P] * 6 + [_I64, _I64, _I32] + [_P] * 6),
    "gsb_fc_emit": (_I32, [_P] * 7 + [_I64, _I64] + [_P] * 13 + [_I64, _P]),
    "gsb_fc_cut_count": (_I32, [_P, _P, _I64, _P, _P, _P, _P]),
    "gsb_fc_cut_emit": (_I32, [_P, _P, _I64, _P, _P, _P, _P, _I64, _P, _P, _P]),
    "gsb_fc_dual_fwd": (_I32, [_P] * 10 + [_I64] + [_P] * 5),
    "gsb_fc_dual_bwd": (_I32, [_P] * 10 + [_I64] + [_P] * 10),
    "gsb_fc_boundary_fwd": (_I32, [_P, _I64, _P, _P, _P, _P, _P, _P]),
    "gsb_fc_boundary_bwd": (_I32, [_P, _I64, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "gsb_light_pdf": (_I32, [_P, _I64, _I64, _P, _P, _P, _P, _P]),
    "gsb_sdf_reg_fwd": (_I32, [_P, _P, _I64, _P, _P]),
    "gsb_sdf_reg_bwd": (_I32, [_P, _P, _I64, _P, _P, _F32, _P, _P]),
    "gsb_mark_visible_boundary": (_I32, [_P, _P, _I64, _I64, _P, _P]),
    "gsb_msdf_reg_fwd": (_I32, [_P, _I64, _P, _P, _I64, _F32, _P, _P]),
    "gsb_msdf_reg_bwd": (_I32, [_P, _I64, _P, _P, _I64, _F32, _P, _F32, _F32, _P, _P, _P]),
    "gsb_image_terms_accumulators": (_I32, []),
    "gsb_image_terms_reduce": (_I32, [_P] * 9 + [_I64, _I32, _I32, _P, _P, _P]),
    "gsb_image_terms_finish": (_I32, [_P, _P, _P, _I64, _I32, _I32, _P, _P, _F32, _P, _P]),
    "gsb_image_terms_bwd": (_I32, [_P] * 9 + [_I64, _I32, _I32, _P, _P, _F32, _P] + [_P] * 8 + [_P]),
    "gsb_gbuffer_fwd": (_I32, [_P] * 7 + [_I64] * 4 + [_P] * 5 + [_P]),
    "gsb_gbuffer_bwd": (_I32, [_P] * 5 + [_I64] * 4 + [_P] * 8 + [_P]),
    "gsb_compose_fwd": (_I32, [_P, _P, _I32, _I64, _I64, _I64, _I32, _I32, _P, _P]),
    "gsb_compose_bwd": (_I32, [_P, _I64, _I64, _I64, _I32, _I32, _P, _P, _P]),
    "gsb_antialias_hash_slots": (_I64, [_I64]),
    "gsb_antialias_item_bytes": (_SZ, []),
    "gsb_antialias_analyse": (_I32, [_P, _P, _P, _I64, _I64, _I64, _I64, _I64, _P, _P, _P, _I64, _P]),
    "gsb_antialias_fwd": (_I32, [_P, _P, _P, _I64, _I64, _P, _P]),
    "gsb_antialias_bwd": (_I32, [_P, _P, _P, _P, _I64, _I64, _P, _I64, _I64, _I64, _P, _P, _P]),
    "gsb_mt_backward": (_I32, [_P, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I64, _P, _P, _P, _P, _P, _P, _P, _P]),
    "gsb_fresnel_shlick_fwd": (_I32, [_P, _P, _P, _I64, _P, _P]),
    "gsb_fresnel_shlick_bwd": (_I32, [_P, _P, _P, _P, _I64, _P, _P, _P, _P]),
    "gsb_ndf_ggx_fwd": (_I32, [_P, _P, _I64, _P, _P]),
    "gsb_ndf_ggx_bwd": (_I32, [_P, _P, _P, _I64, _P, _P, _P]),
    "gsb_lambda_ggx_fwd": (_I32, [_P, _P, _I64, _P, _P]),
    "gsb_lambda_ggx_bwd": (_I32, [_P, _P, _P, _I64, _P, _P, _P]),
    "gsb_masking_smith_fwd": (_I32, [_P, _P, _P, _I64, _P, _P]),
    "gsb_masking_smith_bwd": (_I32, [_P, _P, _P, _P, _I64, _P, _P, _P, _P]),
    "gsb_lambert_fwd": (_I32, [_P, _P, _I64, _P, _P]),
    "gsb_lambert_bwd": (_I32, [_P, _P, _P, _I64, _P, _P, _P]),
    "gsb_frostbite_fwd": (_I32, [_P, _P, _P, _P, _I64, _P, _P]),
    "gsb_frostbite_bwd": (_I32, [_P, _P, _P, _P, _P, _I64, _P, _P, _P, _P, _P]),
    "gsb_pbr_specular_fwd": (_I32, [_P, _P, _P, _P, _P, _F32, _I64, _P, _P]),
    "gsb_pbr_specular_bwd": (_I32, [_P, _P, _P, _P, _P, _F32, _P, _I64, _P, _P, _P, _P, _P, _P]),
    "gsb_pbr_bsdf_fwd": (_I32, [_P, _F32, _I32, _I64, _P, _P]),
    "gsb_pbr_bsdf_bwd": (_I32, [_P, _F32, _I32, _P, _I64, _P, _P]),
    "gsb_xfm_vectors_fwd": (_I32, [_P, _P, _I64, _I64, _I32, _P, _P]),
    "gsb_xfm_vectors_bwd": (_I32, [_P, _P, _I64, _I64, _I32, _P, _P]),
}


class GshellB200LibraryError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise GshellB200LibraryError(
            f"{LIB_PATH} not found: the CUDA library is required (no CPU fallback). "
            "Build it with `python -m gshell_b200.build`.")
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as e:  # pragma: no cover
        raise GshellB200LibraryError(f"cannot load {LIB_PATH}: {e}") from e
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise GshellB200LibraryError(f"{LIB_PATH} does not export {name}; rebuild it") from e
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


# kernels launched by each C-ABI entry point (memsets not counted); used for the `gpu_launches` bench claim
KERNELS_PER_CALL = {"gsb_mt_count": 6, "gsb_mt_emit": 2, "gsb_mt_backward": 2, "gsb_vertex_normals_fwd": 2,
                    "gsb_vertex_normals_bwd": 2, "gsb_rasterize_fwd": 3, "gsb_occluder_build_count": 6,
                    "gsb_occluder_build_fill": 4, "gsb_bilateral_bwd": 3, "gsb_fc_count": 5, "gsb_fc_emit": 3,
                    "gsb_fc_cut_count": 2, "gsb_light_pdf": 2, "gsb_antialias_analyse": 3, "gsb_tangents_fwd": 3, "gsb_tangents_bwd": 4}
launch_count = 0


def check(err: int, what: str, kernels: int = None):
    global launch_count
    if err != 0:
        raise RuntimeError(f"{what} failed with cudaError {err}")
    launch_count += KERNELS_PER_CALL.get(what, 1) if kernels is None else kernels


def require_cuda(t, what):
    """The product has no CPU path: every operator refuses non-CUDA tensors here.  (One function so that the CPU test harness,
    which binds a HOST build of the same kernel source in place of this module, can lift exactly this check: tests/native.)"""
    if not t.is_cuda:
        raise RuntimeError(f"{what}: gshell_b200 runs on CUDA tensors only (no CPU path)")


def synchronize(device=None):
    """Wait for the current stream of `device` (the host reads a count the kernels wrote)."""
    import torch
    torch.cuda.current_stream(device).synchronize()


def ptr(t):
    """Device (or host) address of a contiguous tensor, or None."""
    if t is None:
        return None
    assert t.is_contiguous(), "gshell_b200 kernels take contiguous tensors"
    return t.data_ptr()


def current_stream(device=None):
    """Stream handle for a launch on `device`.  The C entry points launch on the CUDA device that is current in the calling
    thread, so a tensor on another device is an error here rather than a cross-device launch."""
    import torch
    if device is not None:
        d = torch.device(device)
        if d.type == "cuda" and d.index is not None and d.index != torch.cuda.current_device():
            raise RuntimeError(f"gshell_b200: tensors live on {d} but cuda:{torch.cuda.current_device()} is current; "
                               f"wrap the call in torch.cuda.device({d.index})")
    return torch.cuda.current_stream(device).cuda_stream
