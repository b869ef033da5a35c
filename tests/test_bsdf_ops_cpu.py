"""The pointwise BSDF operators of `renderutils` (lambert, frostbite_diffuse, pbr_specular, pbr_bsdf, _fresnel_shlick, _ndf_ggx,
_lambda_ggx, _masking_smith, xfm_vectors; reference render/renderutils/ops.py:91-390, 540-556) against goldens of the reference's
own PyTorch statements of them (tests/golden/bsdf_ops.npz, written by make_golden_bsdf_ops.py from the unmodified
render/renderutils/bsdf.py) -- on the CPU, through the product's own code: csrc/bsdf_ops.cu is "one independent thread per
element" code, compiled unmodified as host code behind the same C ABI (tests/native/host_kernels.py), and the product's Python
layer (broadcasting, autograd nodes, ctypes signatures) runs unchanged on CPU tensors with that library bound in place of
libgshell_b200.so.  Mirrors the reference's tests/test_bsdf.py (operator vs its `use_python=True` twin, values and gradients);
the GPU run of the same comparison is tests/test_zz4_bsdf_ops_gpu.py."""
import os

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))

# golden tag -> (operator name, input names in call order, keyword arguments)
CASES = {
    "fresnel_shlick": ("_fresnel_shlick", ["f0", "f90", "cos"], {}),
    "ndf_ggx": ("_ndf_ggx", ["alpha_sqr", "cos"], {}),
    "lambda_ggx": ("_lambda_ggx", ["alpha_sqr", "cos"], {}),
    "masking_smith": ("_masking_smith", ["alpha_sqr", "cos_i", "cos_o"], {}),
    "pbr_bsdf_broadcast": ("pbr_bsdf", ["kd", "arm", "pos", "nrm", "view_pos", "light_pos"], {"bsdf": "lambert"}),
    "pbr_specular_minrough": ("pbr_specular", ["col", "nrm", "wo", "wi", "alpha"], {"min_roughness": 0.2}),
    "xfm_vectors_shared": ("xfm_vectors", ["vectors", "matrix"], {}),
    "xfm_vectors_batched": ("xfm_vectors", ["vectors", "matrix"], {}),
}
for _tag in ("pos", "mixed"):
    CASES[f"lambert_{_tag}"] = ("lambert", ["nrm", "wi"], {})
    CASES[f"frostbite_{_tag}"] = ("frostbite_diffuse", ["nrm", "wi", "wo", "rough"], {})
    CASES[f"pbr_specular_{_tag}"] = ("pbr_specular", ["col", "nrm", "wo", "wi", "alpha"], {})
    for _b in ("lambert", "frostbite"):
        CASES[f"pbr_bsdf_{_b}_{_tag}"] = ("pbr_bsdf", ["kd", "arm", "pos", "nrm", "view_pos", "light_pos"], {"bsdf": _b})


def check_case(ru, z, tag, device="cpu", rtol=2e-5):
    """Run operator `tag` on the golden inputs; values and gradients against the reference's.  Bound: 2e-5 of the largest
    reference magnitude per array plus 2e-5 relative per element (fp32 chains of ~40 operations; north_star asks 1e-4)."""
    op, names, kw = CASES[tag]
    leaves = [torch.from_numpy(z[f"{tag}.{n}"]).to(device).requires_grad_(n != "matrix") for n in names]
    out = getattr(ru, op)(*leaves, **kw)
    want = torch.from_numpy(z[f"{tag}.out"])
    assert out.shape == want.shape, (tag, out.shape, want.shape)

    def close(got, ref, what):
        tol = rtol * float(ref.abs().max()) + rtol * ref.abs()
        bad = (got.detach().cpu() - ref).abs() > tol
        assert not bool(bad.any()), (tag, what, int(bad.sum()), float((got.detach().cpu() - ref).abs().max()), float(ref.abs().max()))
    close(out, want, "out")
    w = torch.from_numpy(z[f"{tag}.w"]).to(device)
    diff = [t for t in leaves if t.requires_grad]
    grads = torch.autograd.grad((out * w).sum(), diff)
    for n, gg in zip([n for n in names if n != "matrix"], grads):
        ref = torch.from_numpy(z[f"{tag}.g_{n}"])
        assert gg.shape == ref.shape, (tag, n, gg.shape, ref.shape)
        close(gg, ref, "g_" + n)


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(HERE, "golden", "bsdf_ops.npz"))


def test_golden_covers_every_case(golden):
    tags = {k.split(".")[0] for k in golden.files}
    assert tags == set(CASES), tags ^ set(CASES)
    # the masked branches occur in the "mixed" sets and the clamps on both sides in the scalar-term sets
    assert float((golden["pbr_specular_mixed.out"] == 0).mean()) > 0.3 and float((golden["pbr_specular_mixed.out"] != 0).mean()) > 0.05
    c = golden["ndf_ggx.cos"]
    assert (c < 1e-4).any() and (c > 1 - 1e-4).any()


@pytest.mark.parametrize("tag", sorted(CASES))
def test_operator_matches_reference_values_and_gradients(tag, golden, host_kernels_lib, monkeypatch):
    fake, _ = host_kernels_lib
    from gshell_b200.render.renderutils import ops
    from gshell_b200.render import renderutils as ru
    monkeypatch.setattr(ops, "_lib", fake)
    monkeypatch.setattr(ops, "_need_cuda", lambda t, what: None)
    check_case(ru, golden, tag)


def test_operators_refuse_the_python_switch_and_cpu_tensors():
    """No PyTorch fallback inside the product: `use_python=True` and CPU tensors are errors, not silent other paths."""
    from gshell_b200.render import renderutils as ru
    a = torch.rand(1, 2, 2, 3)
    with pytest.raises(NotImplementedError):
        ru.lambert(a, a, use_python=True)
    with pytest.raises(RuntimeError):
        ru.lambert(a, a)
    with pytest.raises(RuntimeError):
        ru.xfm_vectors(torch.rand(1, 4, 3), torch.rand(2, 4, 4))
