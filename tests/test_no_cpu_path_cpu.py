"""The product has no CPU path: nothing under gshell_b200/ imports the oracle, the test harness or a host build of the kernels; in a
normal process the bound library is the CUDA library; every public operator refuses CPU tensors.  (The host emulator of
tests/native is bound only by tests/conftest.py, in processes started with GSB_HOST_EMULATION=1.)"""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "gshell_b200")


def _sources():
    for dp, _, fs in os.walk(PKG):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                yield os.path.join(dp, f)


def test_product_sources_do_not_reach_into_the_test_infrastructure():
    bad = []
    for path in _sources():
        text = open(path, errors="ignore").read()
        code = "\n".join(ln.split("#")[0] for ln in text.splitlines()) if path.endswith(".py") else re.sub(r"//[^\n]*", "", text)
        for pat in (r"^\s*(from|import)\s+oracle\b", r"^\s*from\s+\.+\s*oracle", r"\bhost_kernels\b", r"GSB_HOST_EMULATION", r"bind_host_library",
                    r"^\s*(from|import)\s+tests\b"):
            if re.search(pat, code, re.M):
                bad.append((os.path.relpath(path, ROOT), pat))
    assert not bad, bad


@pytest.mark.skipif(os.environ.get("GSB_HOST_EMULATION") == "1", reason="this process runs on the host emulator by request")
def test_bound_library_is_the_cuda_library_and_operators_refuse_cpu_tensors():
    from gshell_b200 import _lib
    from gshell_b200.geometry.gshell_flexicubes import GShellFlexiCubes
    from gshell_b200.geometry.gshell_tets import GShell_Tets
    from gshell_b200.render import light, mlptexture, raster
    from gshell_b200.render import optixutils as ou
    from gshell_b200.render import renderutils as ru
    assert os.path.basename(_lib.lib._name) == "libgshell_b200.so" and _lib.lib.gsb_compiled_arch() == 100
    x = torch.rand(1, 4, 4, 3)
    calls = [lambda: GShell_Tets()(torch.rand(5, 3), torch.rand(5), torch.rand(5), torch.tensor([[0, 1, 2, 3]])),
             lambda: GShellFlexiCubes(device="cpu")(torch.rand(8, 3), torch.rand(8), torch.rand(8), torch.arange(8)[None], 1),
             lambda: raster.rasterize(torch.rand(1, 3, 4), torch.tensor([[0, 1, 2]], dtype=torch.int32), (4, 4)),
             lambda: ru.xfm_points(torch.rand(1, 3, 3), torch.rand(1, 4, 4)),
             lambda: ru.image_loss(x, x),
             lambda: ru.lambert(x, x),
             lambda: ou.bilateral_denoiser(x, x, x[..., :2], 1.0),
             lambda: light.EnvironmentLight(torch.rand(16, 16, 3)),
             lambda: mlptexture.HashGridEncoding(device="cpu")(torch.rand(4, 3))]
    for k, call in enumerate(calls):
        with pytest.raises(RuntimeError):
            call()
