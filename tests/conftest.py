import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


EMULATED = os.environ.get("GSB_HOST_EMULATION") == "1"


def bind_host_library():
    """GSB_HOST_EMULATION=1: every .cu unit of the product, unmodified, compiled as host code with the thread-block emulator
    (tests/native/host_kernels.py, cuda_host/block_emulator.h) and bound in place of libgshell_b200.so -- the attributes of
    gshell_b200._lib that touch the device are replaced in the module itself, so every product module and every test sees them.
    The `-m gpu` tests then run on CPU tensors (tests/_device.py).  GSB_HOST_SANITIZE=1 adds AddressSanitizer."""
    native = os.path.join(ROOT, "tests", "native")
    sys.path.insert(0, native)
    try:
        import host_kernels
    finally:
        sys.path.remove(native)
    from gshell_b200 import _lib, build
    # GSB_HOST_DEFINES="A B=1": extra -D switches of the build (e.g. GSB_TRACE_STATS: the trace kernel's work counters)
    lib = host_kernels.build(build.sources(), sanitize=os.environ.get("GSB_HOST_SANITIZE") == "1", blocks=True,
                             defines=tuple(os.environ.get("GSB_HOST_DEFINES", "").split()))
    for name, (res, args) in _lib.SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    # GSB_HOST_ORDER_SEED=k: blocks of every launch in a shuffled order (odd k: the threads of a block in descending order too) --
    # results must not depend on the order in which a device happens to schedule blocks
    host_kernels.set_thread_order(lib, int(os.environ.get("GSB_HOST_ORDER_SEED", "0")))
    _lib.lib = lib
    _lib.current_stream = lambda device=None: None
    _lib.require_cuda = lambda t, what: None
    _lib.synchronize = lambda device=None: None
    return lib


def pytest_sessionstart(session):
    if EMULATED:
        bind_host_library()


def pytest_collection_modifyitems(config, items):
    """`pytest tests` in a checkout without a CUDA device skips the gpu-marked tests instead of failing on the driver."""
    if EMULATED:
        return
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


HOST_UNITS = ["mesh_ops.cu", "tangents.cu", "auggrid.cu", "bsdf_ops.cu"]      # "one independent thread per element" translation units


@pytest.fixture(scope="session")
def host_kernels_lib():
    """(stand-in for gshell_b200._lib bound to the HOST build of the kernel source, the builder module).

    tests/native/host_kernels.py compiles the unmodified .cu files above as host code behind the same C ABI; the stand-in carries
    the product's own ctypes signatures, so `monkeypatch.setattr(module, "_lib", stand_in)` makes the product's Python layer run
    its kernels on CPU tensors."""
    import types
    native = os.path.join(ROOT, "tests", "native")
    sys.path.insert(0, native)
    try:
        import host_kernels
    finally:
        sys.path.remove(native)
    from gshell_b200 import _lib
    # GSB_HOST_SANITIZE=1 (with LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0): AddressSanitizer build,
    # every out-of-bounds read or write of a kernel thread aborts the run
    lib = host_kernels.build(HOST_UNITS, sanitize=os.environ.get("GSB_HOST_SANITIZE") == "1")
    for name, (res, args) in _lib.SIGNATURES.items():
        fn = getattr(lib, name, None)
        if fn is not None:
            fn.restype, fn.argtypes = res, args
    stand_in = types.SimpleNamespace(lib=lib, ptr=_lib.ptr, check=_lib.check, current_stream=lambda device=None: None)
    return stand_in, host_kernels
