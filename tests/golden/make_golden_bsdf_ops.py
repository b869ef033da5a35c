"""Generate tests/golden/bsdf_ops.npz from the reference's own PyTorch statements of its pointwise BSDF operators
(render/renderutils/ops.py with use_python=True -> render/renderutils/bsdf.py:57-151) and of xfm_vectors (ops.py:552), run
unmodified on the CPU through _ref_shim.  Input distributions follow the reference's tests/test_bsdf.py (uniform [0,1) vectors,
cosines stretched past both clamps) on a 12 x 12 image; every case stores inputs, output, the weights w of the scalar
sum(out * w) and autograd's gradients of that scalar.
Run in the build container only:   python tests/golden/make_golden_bsdf_ops.py"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _ref_shim import reference_on_cpu   # noqa: E402

RES = 12


def main():
    warnings.filterwarnings("ignore")
    g = torch.Generator().manual_seed(11)
    R = lambda *s: torch.rand(*s, generator=g)          # noqa: E731
    N = lambda *s: torch.randn(*s, generator=g)         # noqa: E731
    img3 = lambda: R(1, RES, RES, 3)                    # noqa: E731
    img1 = lambda: R(1, RES, RES, 1)                    # noqa: E731
    store = {}

    def case(tag, fn, names, inputs, **kw):
        leaves = [t.clone().requires_grad_() for t in inputs]
        out = fn(*leaves, **kw)
        w = N(*out.shape)
        gs = torch.autograd.grad((out * w).sum(), leaves)
        for nme, t, gg in zip(names, inputs, gs):
            store[f"{tag}.{nme}"] = t.numpy()
            store[f"{tag}.g_{nme}"] = gg.numpy()
        store[f"{tag}.out"] = out.detach().numpy()
        store[f"{tag}.w"] = w.numpy()
        print(tag, tuple(out.shape), float(out.abs().mean()))

    with reference_on_cpu() as imp:
        ru = imp("render.renderutils")
        py = dict(use_python=True)
        case("fresnel_shlick", ru._fresnel_shlick, ["f0", "f90", "cos"], [img3(), img3(), img1() * 2.0 - 0.5], **py)
        case("ndf_ggx", ru._ndf_ggx, ["alpha_sqr", "cos"], [img1(), img1() * 3.0 - 1.0], **py)
        case("lambda_ggx", ru._lambda_ggx, ["alpha_sqr", "cos"], [img1(), img1() * 3.0 - 1.0], **py)
        case("masking_smith", ru._masking_smith, ["alpha_sqr", "cos_i", "cos_o"], [img1(), img1() * 1.4 - 0.2, img1() * 1.4 - 0.2], **py)
        # vectors: the reference's tests use un-normalised uniform [0,1) vectors; a second set is centred so that back-facing
        # configurations (the masked branches) occur too
        for tag, shift in (("pos", 0.0), ("mixed", 0.5)):
            v = lambda: img3() - shift                  # noqa: E731
            case(f"lambert_{tag}", ru.lambert, ["nrm", "wi"], [v(), v()], **py)
            case(f"frostbite_{tag}", ru.frostbite_diffuse, ["nrm", "wi", "wo", "rough"], [v(), v(), v(), img1()], **py)
            case(f"pbr_specular_{tag}", ru.pbr_specular, ["col", "nrm", "wo", "wi", "alpha"], [img3(), v(), v(), v(), img1()], **py)
            for bsdf in ("lambert", "frostbite"):
                case(f"pbr_bsdf_{bsdf}_{tag}", ru.pbr_bsdf, ["kd", "arm", "pos", "nrm", "view_pos", "light_pos"],
                     [img3(), img3(), v(), v(), v() * 3.0, v() * 3.0], bsdf=bsdf, **py)
        # broadcast operands, as the callers of the reference pass them (camera / light position per batch item)
        case("pbr_bsdf_broadcast", ru.pbr_bsdf, ["kd", "arm", "pos", "nrm", "view_pos", "light_pos"],
             [R(2, RES, RES, 3), R(2, RES, RES, 3), N(2, RES, RES, 3) * 0.3, N(2, RES, RES, 3), N(2, 1, 1, 3) * 3.0, N(1, 1, 1, 3) * 3.0],
             bsdf="lambert", **py)
        case("pbr_specular_minrough", ru.pbr_specular, ["col", "nrm", "wo", "wi", "alpha"],
             [img3(), img3(), img3(), img3(), img1() * 0.1], min_roughness=0.2, **py)
        case("xfm_vectors_shared", ru.xfm_vectors, ["vectors", "matrix"], [N(1, 37, 3), N(3, 4, 4)], **py)
        case("xfm_vectors_batched", ru.xfm_vectors, ["vectors", "matrix"], [N(3, 37, 3), N(3, 4, 4)], **py)
    path = os.path.join(HERE, "bsdf_ops.npz")
    np.savez_compressed(path, **store)
    print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
