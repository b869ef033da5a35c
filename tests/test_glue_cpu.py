"""Loss-assembly glue of tick() against the unmodified reference's values and gradients
(tests/golden/glue_losses.npz, generator tests/golden/make_golden_glue.py).  These pieces are plain torch in the product,
so the comparison runs on CPU."""
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def _load():
    z = np.load(os.path.join(HERE, "golden", "glue_losses.npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def _close(a, b, tol=1e-5):
    return torch.allclose(a, b, rtol=tol, atol=tol * float(b.abs().max()) + 1e-12)


def test_regularizers_match_reference():
    from gshell_b200.render import regularizer as reg
    g = _load()
    leaves = {k: g[k].clone().requires_grad_() for k in ("diff", "spec", "kd", "kd_grad", "ks_grad", "nrm_grad")}
    l_sh = reg.shading_loss(leaves["diff"], leaves["spec"], g["color_ref"], 0.15, 0.0025)
    l_ch = reg.chroma_loss(leaves["kd"], g["color_ref"], 0.3)
    l_ms = reg.material_smoothness_grad(leaves["kd_grad"], leaves["ks_grad"], leaves["nrm_grad"], lambda_kd=0.25, lambda_ks=0.1,
                                        lambda_nrm=0.05)
    assert _close(l_sh, g["shading_loss"]) and _close(l_ch, g["chroma_loss"]) and _close(l_ms, g["material_smoothness"])
    grads = torch.autograd.grad(l_sh + 2.0 * l_ch + 3.0 * l_ms, list(leaves.values()))
    for k, got in zip(leaves, grads):
        assert _close(got, g[f"g_{k}"]), k


def test_sdf_regulariser_matches_reference():
    from gshell_b200.geometry.gshell_tets_geometry import compute_sdf_reg_loss
    g = _load()
    sdf = g["sdf"].clone().requires_grad_()
    loss = compute_sdf_reg_loss(sdf, g["edges"])
    assert _close(loss, g["sdf_reg"])
    assert _close(torch.autograd.grad(loss, sdf)[0], g["g_sdf"])
    assert int((g["sdf"] == 0).sum()) > 0            # the sign(0) edge case is part of the fixture
