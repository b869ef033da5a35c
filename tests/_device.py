"""Device of the `-m gpu` tests: cuda:0 -- or, with GSB_HOST_EMULATION=1, the CPU: tests/conftest.py then binds a HOST build of
the unmodified kernel source (tests/native/host_kernels.py with the thread-block emulator) in place of libgshell_b200.so, so the
same test bodies drive the same Python layer and the same kernels without a GPU (tests/test_emulated_gpu_suite_cpu.py)."""
import os

import torch

EMULATED = os.environ.get("GSB_HOST_EMULATION") == "1"
DEVICE = "cpu" if EMULATED else "cuda:0"


def device():
    assert EMULATED or torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device(DEVICE)


def synchronize():
    if not EMULATED:
        torch.cuda.synchronize()
