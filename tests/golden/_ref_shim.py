"""Import the UNMODIFIED reference (`/root/reference`) on a CPU-only box.

Used only by the `make_golden_*.py` generators in this directory (run in the build container, where
the read-only reference checkout is mounted).  Nothing under tests/ that runs with `-m gpu`, nor
bench.py / smoke(), imports this module: `/root/reference` does not exist on the GPU box.

Obstacles handled:
  * `render/util.py` imports `nvdiffrast.torch` and `imageio` at module scope -> empty stub modules;
  * `device='cuda'` is hard-coded in the geometry modules -> torch factory functions are wrapped so
    that a `device='cuda'` argument is rewritten to 'cpu', and `Tensor.cuda()` is a no-op.
"""
import contextlib
import functools
import importlib
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("GSHELL_REFERENCE", "/root/reference")

_FACTORIES = ["tensor", "arange", "ones", "zeros", "linspace", "rand", "randn", "empty", "full",
              "ones_like", "zeros_like", "rand_like", "randn_like", "eye", "meshgrid", "as_tensor"]


def _cpuify(fn):
    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        dev = kwargs.get("device", None)
        if dev is not None and "cuda" in str(dev):
            kwargs["device"] = "cpu"
        return fn(*args, **kwargs)
    return wrapped


@contextlib.contextmanager
def reference_on_cpu():
    """Context in which `import geometry.gshell_tets` etc. resolve to the reference, on CPU."""
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}")
    saved_modules = {k: sys.modules.get(k) for k in ("nvdiffrast", "nvdiffrast.torch", "imageio")}
    for name in saved_modules:
        if saved_modules[name] is None:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["nvdiffrast"].torch = sys.modules["nvdiffrast.torch"]
    saved_fns = {n: getattr(torch, n) for n in _FACTORIES}
    for n, f in saved_fns.items():
        setattr(torch, n, _cpuify(f))
    saved_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    sys.path.insert(0, REFERENCE_ROOT)
    # make sure `geometry` / `render` resolve to the reference, not to anything of ours
    shadowed = {k: sys.modules.pop(k) for k in list(sys.modules)
                if k.split(".")[0] in ("geometry", "render", "denoiser")}
    try:
        yield importlib.import_module
    finally:
        for k in list(sys.modules):
            if k.split(".")[0] in ("geometry", "render", "denoiser"):
                del sys.modules[k]
        sys.modules.update(shadowed)
        sys.path.remove(REFERENCE_ROOT)
        torch.Tensor.cuda = saved_cuda
        for n, f in saved_fns.items():
            setattr(torch, n, f)
        for name, mod in saved_modules.items():
            if mod is None:
                sys.modules.pop(name, None)
