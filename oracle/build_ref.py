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
# Other code generations of the SAME unmodified source, used only by tests/test_oracle_env_shade_conditioning.py to measure how far
# the reference's integrator moves against ITSELF when the compiler is allowed what nvcc / NVRTC do to it on the GPU:
#   "fma"  = mul+add contraction (nvcc's default -fmad=true);  "fast" = contraction + -ffast-math (the reference JIT-compiles its
#   program with --use_fast_math, render/optixutils/c_src/optix_wrapper.cpp:31-41)
VARIANTS = {"": ["-ffp-contract=off"], "fma": ["-ffp-contract=fast", "-mfma"], "fast": ["-ffp-contract=fast", "-mfma", "-ffast-math"]}


def _out(variant):
    return OUT if not variant else os.path.join(OUT_DIR, f"libref_env_shade_{variant}.so")


def available(variant=""):
    return os.path.exists(_out(variant))


def build(force=False, variant=""):
    """Build if the reference checkout is present; returns the library path or None (GPU box: use the prebuilt file)."""
    OUT = _out(variant)
    if not os.path.isfile(KERNEL):
        return OUT if available(variant) else None
    deps = [KERNEL, os.path.join(HERE, "ref_env_shade_driver.cpp"), os.path.join(HERE, "ref_shim", "host_cuda.h"),
            os.path.join(HERE, "ref_shim", "optix.h")]
    if not force and available(variant) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")      # <math_constants.h> only
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = [shutil.which("g++") or "g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fopenmp", *VARIANTS[variant],
           "-I", os.path.join(HERE, "ref_shim"), "-I", cuda_inc, "-include", os.path.join(HERE, "ref_shim", "host_cuda.h"),
           f'-DREF_KERNEL_CU="{KERNEL}"', "-x", "c++", os.path.join(HERE, "ref_env_shade_driver.cpp"), "-o", OUT]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
