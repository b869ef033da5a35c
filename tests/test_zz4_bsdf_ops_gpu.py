"""GPU run of the pointwise BSDF operators of `renderutils` (csrc/bsdf_ops.cu) against goldens of the reference's own PyTorch
statements (tests/golden/bsdf_ops.npz): the same comparison as tests/test_bsdf_ops_cpu.py -- which runs the unmodified kernel
source as host code -- through libgshell_b200.so on the device.  Mirrors the reference's tests/test_bsdf.py (operator vs its
`use_python=True` twin, values and gradients).  Sorts last: the operators are off the training loop and were written after the
round's last GPU run, so this is their first device run; a surprise here must not stop the `-x` run before the established suites."""
import os
import sys

import numpy as np
import pytest
import torch

from _device import DEVICE, device      # cuda:0, or the CPU under the host emulator (tests/_device.py)

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)
from test_bsdf_ops_cpu import CASES, check_case      # noqa: E402  (the case table and the comparison are shared)


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(HERE, "golden", "bsdf_ops.npz"))


@pytest.mark.parametrize("tag", sorted(CASES))
def test_operator_matches_reference_values_and_gradients_on_device(tag, golden):
    from gshell_b200.render import renderutils as ru
    # fp32 with FMA contraction on the device: same bound as on the host (2e-5 of the array's largest magnitude + 2e-5 relative)
    check_case(ru, golden, tag, device=DEVICE)


def test_large_image_runs_and_is_finite():
    """One launch at G-buffer size (8 x 1024^2 elements, grids of 32 768 blocks) -- shapes and finiteness only."""
    from gshell_b200.render import renderutils as ru
    d = device()
    g = torch.Generator(device=d).manual_seed(0)
    v = lambda: torch.rand(8, 1024, 1024, 3, device=d, generator=g)      # noqa: E731
    kd, arm, pos, nrm = v().requires_grad_(), v(), v() - 0.5, v()
    view, light = torch.tensor([[[[0.0, 0.0, 3.0]]]], device=d), torch.tensor([[[[1.0, 2.0, 3.0]]]], device=d)
    out = ru.pbr_bsdf(kd, arm, pos, nrm, view, light, bsdf="frostbite")
    assert out.shape == (8, 1024, 1024, 3) and bool(torch.isfinite(out).all())
    out.sum().backward()
    assert kd.grad.shape == kd.shape and bool(torch.isfinite(kd.grad).all())
