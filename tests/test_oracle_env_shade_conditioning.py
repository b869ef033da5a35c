"""How far does the reference's env-light integrator move against ITSELF when only the code generation changes?

north_star asks for 1e-4 relative on shaded pixels and gradients.  tests/test_shade_gpu.py::test_env_shade_vs_oracle asserts
that bar for roughness >= 0.3 and a wider one over the reference's full roughness range (minimum 0.08), arguing that the GGX
lobe amplifies fp32 rounding there.  This file MEASURES that claim on the reference's own code: the unmodified
render/optixutils/c_src/envsampling/kernel.cu is compiled three times for the CPU (oracle/build_ref.py) --

    ""      no contraction, precise libm          (the checker every parity test uses)
    "fma"   mul+add contraction                   (what nvcc does by default, -fmad=true)
    "fast"  contraction + -ffast-math             (the reference JIT-compiles its program with --use_fast_math,
                                                   render/optixutils/c_src/optix_wrapper.cpp:31-41)

-- and the three builds are run on exactly the inputs of the GPU parity test.  Measured here (x86-64, g++ 13):

    roughness >= 0.3 :  no pixel moves by more than 3e-5 relative, gradients agree to 2e-5 relative L2
    roughness >= 0.08:  3.5 - 6.7 % of the covered pixels move by more than 1e-4 in the specular term, the largest by
                        0.7e-3 - 1.9e-3; gradients move by up to 1.2e-3 relative L2 (d_pos)

so 1e-4 per pixel is not a property the reference has against itself below roughness 0.3, and the bounds of the GPU test's
full-range case (8 % of the pixels, 5e-3 largest, 1e-2 relative L2 on gradients) sit within 3x (values) / 8x (gradients) of the reference's own spread.
The assertions below keep both halves of that statement true.

Needs the reference checkout (build container) or prebuilt variant libraries; skipped otherwise (e.g. on the GPU box)."""
import importlib.util
import os

import pytest
import torch

from oracle import build_ref, shade_oracle as so

HERE = os.path.dirname(os.path.abspath(__file__))


def _inputs():
    spec = importlib.util.spec_from_file_location("_shade_gpu_inputs", os.path.join(HERE, "test_shade_gpu.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m._shade_inputs


def _run(ref, variant, a, wd, ws, n, seed):
    with ref.code_generation(variant):
        d, s = ref.env_shade_fwd(*a, bsdf=0, n_samples_x=n, rnd_seed=seed)
        g = ref.env_shade_bwd(*a, wd, ws, bsdf=0, n_samples_x=n, rnd_seed=seed)
    return d, s, g


def _spread(base, other, cov):
    """The statistics test_env_shade_vs_oracle asserts: (fraction of covered pixels off by > 1e-4, largest relative move) over
    both radiance terms, and the largest relative L2 move of a gradient."""
    bad = worst = 0.0
    for got, want in ((other[0], base[0]), (other[1], base[1])):
        floor = 1e-3 * want[cov].abs().mean().clamp(min=1e-8)
        rel = ((got - want).abs() / want.abs().clamp(min=floor))[cov]
        bad, worst = max(bad, float((rel > 1e-4).float().mean())), max(worst, float(rel.max()))
    l2 = max(float((x - y).norm() / y.norm().clamp(min=1e-12)) for x, y in zip(other[2], base[2]))
    return bad, worst, l2


@pytest.mark.parametrize("variant", ["fma", "fast"])
def test_reference_integrator_against_its_own_other_code_generation(variant):
    if build_ref.build() is None or build_ref.build(variant=variant) is None:
        pytest.skip("needs the reference checkout (or prebuilt oracle/_ref variant libraries)")
    from oracle import ref_env_shade as ref
    shade_inputs = _inputs()
    B, H, W, n = 2, 24, 20, 4
    rows_out = []
    for rough_min, seed in ((0.3, 1), (0.08, 4), (0.08, 5), (0.08, 6)):        # (0.3, 1) and (0.08, 4) are the GPU test's cases
        mask, pos, nrm, view, kd, ks, light = shade_inputs(B, H, W, seed, rough_min=rough_min)
        pdf, rows, cols = so.light_pdf_tables(light)
        perms = torch.argsort(torch.rand(32768, n * n, generator=torch.Generator().manual_seed(seed)), dim=-1).int()
        a = (mask, pos + 0.001 * nrm, pos, nrm, view, kd, ks, light, pdf, rows, cols, perms)
        gen = torch.Generator().manual_seed(99)
        wd, ws = torch.randn(B, H, W, 3, generator=gen), torch.randn(B, H, W, 3, generator=gen)
        base = _run(ref, "", a, wd, ws, n, 17 + seed)
        bad, worst, l2 = _spread(base, _run(ref, variant, a, wd, ws, n, 17 + seed), mask > 0)
        rows_out.append((rough_min, seed, bad, worst, l2))
        print(f"reference vs itself ({variant}) rough_min={rough_min} seed={seed}: {bad:.2%} of pixels > 1e-4, max {worst:.2e}, "
              f"gradients rel L2 <= {l2:.2e}")
        if rough_min >= 0.3:
            # well conditioned: the strict bar of the GPU test is one the reference meets against itself
            assert bad == 0.0 and worst < 1e-4 and l2 < 1e-4, rows_out[-1]
        else:
            # ill conditioned: 1e-4 per pixel fails for the reference against itself ...
            assert bad > 0.01 and worst > 3e-4, rows_out[-1]
            # ... and the GPU test's full-range bounds are no looser than a small multiple of this spread
            assert bad < 0.08 and worst < 5e-3 and l2 < 5e-3, rows_out[-1]
