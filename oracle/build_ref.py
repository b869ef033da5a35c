"""TEST INFRASTRUCTURE: compile the reference's own env-light integrator for the CPU.

    python oracle/build_ref.py        ->  oracle/_ref/libref_env_shade.so   (git-ignored, travels to the GPU box)

Recipe: g++ on the reference's UNMODIFIED `render/optixutils/c_src/envsampling/kernel.cu` (which pulls in params.h, common.h,
math_utils.h, bsdf.h, accessor.h next to it) where it lies under the reference checkout, through the host driver
oracle/ref_env_shade_driver.cpp and the two shim headers in oracle/ref_shim/ (execution-space qualifiers / vector types /
min-max overloads, and the five OptiX device intrinsics the program uses).  No reference source is copied into this repository;
the reference's own build system (torch cpp_extension + NVRTC + OptiX SDK) is not run.  FP contraction is off and libm is the
host's precise one, whereas the reference JIT-compiles with --use_fast_math: expect ~1e-6 relative differences to a GPU run.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("GSHELL_REFERENCE", "/root/reference")
KERNEL = os.path.join(REFERENCE_ROOT, "render", "optixutils", "c_src", "envsampling", "kernel.cu")
OUT_DIR = os.path.join(HERE, "_ref")
OUT = os.path.join(OUT_DIR, "libref_env_shade.so")


def available():
    return os.path.exists(OUT)


def build(force=False):
    """Build if the reference checkout is present; returns the library path or None (GPU box: use the prebuilt file)."""
    if not os.path.isfile(KERNEL):
        return OUT if available() else None
    deps = [KERNEL, os.path.join(HERE, "ref_env_shade_driver.cpp"), os.path.join(HERE, "ref_shim", "host_cuda.h"),
            os.path.join(HERE, "ref_shim", "optix.h")]
    if not force and available() and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")      # <math_constants.h> only
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = [shutil.which("g++") or "g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off",
           "-I", os.path.join(HERE, "ref_shim"), "-I", cuda_inc, "-include", os.path.join(HERE, "ref_shim", "host_cuda.h"),
           f'-DREF_KERNEL_CU="{KERNEL}"', "-x", "c++", os.path.join(HERE, "ref_env_shade_driver.cpp"), "-o", OUT]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
