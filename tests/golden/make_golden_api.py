"""Generate tests/golden/api_signatures.json: parameter lists of the reference's callables at the drop-in boundary,
read from the reference SOURCE with `ast` (no import: most of these modules need OptiX / nvdiffrast / tiny-cuda-nn).
Run in the build container only:   python tests/golden/make_golden_api.py"""
import ast
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _ref_shim import REFERENCE_ROOT   # noqa: E402

# reference file -> callables ("Class.method" or "function") that train scripts / geometry modules call
BOUNDARY = {
    "geometry/gshell_tets.py": ["GShell_Tets.__init__", "GShell_Tets.__call__", "GShell_Tets.marching_from_auggrid"],
    "geometry/gshell_flexicubes.py": ["GShellFlexiCubes.__init__", "GShellFlexiCubes.construct_voxel_grid", "GShellFlexiCubes.__call__"],
    "geometry/gshell_tets_geometry.py": ["GShellTetsGeometry.__init__", "GShellTetsGeometry.getMesh", "GShellTetsGeometry.render",
                                         "GShellTetsGeometry.tick", "GShellTetsGeometry.getAABB", "GShellTetsGeometry.clamp_deform",
                                         "GShellTetsGeometry.getMesh_from_augmented_grid_withocc", "compute_sdf_reg_loss"],
    "geometry/gshell_flexicubes_geometry.py": ["GShellFlexiCubesGeometry.__init__", "GShellFlexiCubesGeometry.getMesh",
                                               "GShellFlexiCubesGeometry.tick"],
    "render/render.py": ["shade", "render_layer", "render_mesh", "render_uv"],
    "render/renderutils/ops.py": ["xfm_points", "xfm_vectors", "prepare_shading_normal", "image_loss", "lambert", "frostbite_diffuse", "pbr_specular",
                                  "pbr_bsdf", "_fresnel_shlick", "_ndf_ggx", "_lambda_ggx", "_masking_smith"],
    "render/optixutils/ops.py": ["optix_build_bvh", "optix_env_shade", "bilateral_denoiser"],
    "render/light.py": ["EnvironmentLight.__init__", "EnvironmentLight.update_pdf", "EnvironmentLight.clamp_", "EnvironmentLight.generate_image",
                        "EnvironmentLight.xfm", "EnvironmentLight.parameters", "EnvironmentLight.clone", "create_trainable_env_rnd", "load_env",
                        "save_env_map"],
    "render/mesh.py": ["Mesh.__init__", "auto_normals", "compute_tangents"],
    "render/regularizer.py": ["chroma_loss", "shading_loss", "material_smoothness_grad"],
    "render/mlptexture.py": ["MLPTexture3D.__init__", "MLPTexture3D.sample", "MLPTexture3D.clamp_", "MLPTexture3D.cleanup"],
    "denoiser/denoiser.py": ["BilateralDenoiser.__init__", "BilateralDenoiser.set_influence", "BilateralDenoiser.forward"],
}


def signature(fn):
    a = fn.args
    names = [x.arg for x in a.posonlyargs + a.args]
    defaults = [None] * (len(names) - len(a.defaults)) + [ast.unparse(d) for d in a.defaults]
    out = [[n, d] for n, d in zip(names, defaults)]
    if a.vararg:
        out.append(["*" + a.vararg.arg, None])
    out += [[k.arg, None if d is None else ast.unparse(d)] for k, d in zip(a.kwonlyargs, a.kw_defaults)]
    if a.kwarg:
        out.append(["**" + a.kwarg.arg, None])
    return out


def main():
    result = {}
    for rel, wanted in BOUNDARY.items():
        tree = ast.parse(open(os.path.join(REFERENCE_ROOT, rel)).read())
        found = {}
        for node in tree.body:
            if isinstance(node, ast.FunctionDef):
                found[node.name] = node
            elif isinstance(node, ast.ClassDef):
                for sub in node.body:
                    if isinstance(sub, ast.FunctionDef):
                        found[f"{node.name}.{sub.name}"] = sub
        for name in wanted:
            if name not in found:
                print("not in reference:", rel, name)
                continue
            result[f"{rel}::{name}"] = signature(found[name])
    path = os.path.join(HERE, "api_signatures.json")
    json.dump(result, open(path, "w"), indent=1, sort_keys=True)
    print(len(result), "signatures ->", path)


if __name__ == "__main__":
    main()
