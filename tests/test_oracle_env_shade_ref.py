"""Pins oracle/shade_oracle.py::env_shade (the checker of the CUDA integrator) to the REFERENCE'S OWN integrator: the unmodified
render/optixutils/c_src/envsampling/kernel.cu compiled for the CPU (oracle/build_ref.py; OptiX intrinsics stubbed, shadow
rays answered by a brute-force any-hit test).  Two layers:
  * tests/golden/shade_envshade_ref.npz (generator beside it): values and gradients of the compiled reference, always checked;
  * the compiled library itself on fresh seeds, when it is present (built here from /root/reference; the .so travels to the GPU
    box, where the reference checkout does not exist).
The oracle's gradients come from autograd, the reference's from its hand-written bwd* functions (bsdf.h): agreement pins both."""
import os

import numpy as np
import pytest
import torch

from oracle import shade_oracle as so

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = ["pbr_n4", "pbr_rough008_n3", "diffuse_n4", "white_n2", "pbr_shadow_n3", "pbr_halfshadow_n2"]


def brute_force_visibility(verts, tris):
    """Stand-in for the shadow ray inside the oracle: two-sided Moeller-Trumbore against every triangle, t in (0, 1e16)."""
    verts, tris = verts.double(), tris.long()
    v0 = verts[tris[:, 0]]; e1 = verts[tris[:, 1]] - v0; e2 = verts[tris[:, 2]] - v0

    def vis(o, d):
        o, d = o.detach().double(), d.detach().double()
        p = torch.linalg.cross(d[:, None, :].expand(-1, e2.shape[0], -1), e2[None].expand(d.shape[0], -1, -1))
        det = (e1[None] * p).sum(-1)
        ok = det != 0
        inv = 1.0 / torch.where(ok, det, torch.ones_like(det))
        t_ = o[:, None, :] - v0[None]
        u = (t_ * p).sum(-1) * inv
        q = torch.linalg.cross(t_, e1[None].expand_as(t_))
        v = (d[:, None, :] * q).sum(-1) * inv
        t = (e2[None] * q).sum(-1) * inv
        hit = ok & (u >= 0) & (u <= 1) & (v >= 0) & (u + v <= 1) & (t > 0) & (t < 1e16)
        return (~hit.any(1)).float()[:, None]
    return vis


def run_oracle(c, want_grads=True):
    pdf, rows, cols = so.light_pdf_tables(c["light"])
    leaves = [c[k].clone().requires_grad_() for k in ("pos", "nrm", "kd", "ks", "light")]
    shadow = float(c["shadow"])
    vis = brute_force_visibility(c["verts"], c["tris"]) if shadow > 0 else None
    ro = (leaves[0] + 0.001 * leaves[1]).detach()
    d, s = so.env_shade(c["mask"], ro, leaves[0], leaves[1], c["view"], leaves[2], leaves[3], leaves[4], pdf, rows, cols,
                        c["perms"], bsdf=int(c["bsdf"]), n_samples_x=int(c["n"]), rnd_seed=int(c["seed"]), shadow_scale=shadow,
                        visibility=vis)
    grads = None
    if want_grads:
        grads = torch.autograd.grad((d * c["gd"]).sum() + (s * c["gs"]).sum(), leaves, allow_unused=True)
        grads = [torch.zeros_like(l) if g is None else g for g, l in zip(grads, leaves)]
    return d.detach(), s.detach(), grads


def check(c, d, s, grads, strict):
    cov = c["mask"] > 0
    assert float(d[~cov].abs().max()) == 0 and float(c["diff"][~cov].abs().max()) == 0      # masked pixels stay zero
    for name, got, want in (("diff", d, c["diff"]), ("spec", s, c["spec"])):
        if float(want.abs().max()) == 0:
            assert float(got.abs().max()) == 0, name
            continue
        err = (got - want).abs() / want.abs().max()
        # one of the 2 n^2 samples of a pixel may take the other side of a discrete decision (CDF bin, lobe, texel) on a 1-ulp
        # libm difference; the fixtures contain no such pixel for the strict cases
        assert float(err.max()) < (1e-5 if strict else 1e-4), (name, float(err.max()))
    for name, got, want in zip(("pos", "nrm", "kd", "ks", "light"), grads, (c[f"g_{k}"] for k in ("pos", "nrm", "kd", "ks", "light"))):
        if float(want.abs().max()) == 0:
            assert float(got.abs().max()) == 0, name
            continue
        l2 = float((got - want).norm() / want.norm())
        assert l2 < (5e-5 if strict else 5e-3), (name, l2)


def _golden(case):
    z = np.load(os.path.join(HERE, "golden", "shade_envshade_ref.npz"))
    return {k.split("/", 1)[1]: torch.from_numpy(z[k]) for k in z.files if k.startswith(case + "/")}


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_compiled_reference_golden(case):
    c = _golden(case)
    d, s, grads = run_oracle(c)
    check(c, d, s, grads, strict="rough008" not in case)


def test_oracle_matches_compiled_reference_live():
    """Fresh seeds through the compiled reference itself (skipped when neither the .so nor the reference checkout exists)."""
    from oracle import build_ref
    if build_ref.build() is None:
        pytest.skip("oracle/_ref/libref_env_shade.so not built and no reference checkout")
    from oracle import ref_env_shade as ref
    import importlib.util
    spec = importlib.util.spec_from_file_location("mk", os.path.join(HERE, "golden", "make_golden_envshade_ref.py"))
    mk = importlib.util.module_from_spec(spec); spec.loader.exec_module(mk)
    for seed, bsdf, n, shadow in ((31, 0, 3, 0.0), (32, 0, 2, 1.0), (33, 1, 3, 1.0)):
        sc = mk.scene(seed, 1, 9, 8, 0.3)
        g = torch.Generator().manual_seed(seed)
        perms = torch.argsort(torch.rand(53, n * n, generator=g), dim=-1).int()
        pdf, rows, cols = so.light_pdf_tables(sc["light"])
        ro = sc["pos"] + 0.001 * sc["nrm"]
        args = (sc["mask"], ro, sc["pos"], sc["nrm"], sc["view"], sc["kd"], sc["ks"], sc["light"], pdf, rows, cols, perms)
        kw = dict(bsdf=bsdf, n_samples_x=n, rnd_seed=seed, shadow_scale=shadow, verts=sc["verts"], tris=sc["tris"])
        diff, spec_ = ref.env_shade_fwd(*args, **kw)
        gd, gs = torch.rand(diff.shape, generator=g), torch.rand(diff.shape, generator=g)
        gr = ref.env_shade_bwd(*args, gd, gs, **kw)
        c = dict(sc, perms=perms, diff=diff, spec=spec_, gd=gd, gs=gs, bsdf=torch.tensor(bsdf), n=torch.tensor(n), seed=torch.tensor(seed),
                 shadow=torch.tensor(shadow), **{f"g_{k}": v for k, v in zip(("pos", "nrm", "kd", "ks", "light"), gr)})
        d, s, grads = run_oracle(c)
        check(c, d, s, grads, strict=True)
