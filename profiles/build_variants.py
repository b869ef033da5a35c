"""Profiling builds of libgshell_b200.so that differ only in the trace kernel's compile-time knobs (csrc/occluder.cu).
Run in the build container; the .so files land in profiles/_variants/ (git-ignored, shipped to the GPU box) and are picked
with GSB_LIB_PATH=... by profiles/prof_shadow.py.   usage: python profiles/build_variants.py name:K=V,K=V ..."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gshell_b200 import build as b   # noqa: E402

OUT = os.path.join(ROOT, "profiles", "_variants")


def one(spec):
    """name:K=V,...  -> occluder.cu rebuilt with -DGSB_TRACE_K=V;  name:ES:K=V,... -> env_shade.cu rebuilt with -DK=V"""
    name, _, kv = spec.partition(":")
    src = "occluder.cu"
    if kv.startswith("ES:"):
        src, kv = "env_shade.cu", kv[3:]
        defs = [f"-D{x}" for x in kv.split(",") if x] + ["-use_fast_math"]
    else:
        defs = [f"-DGSB_TRACE_{x}" if not x.startswith("GSB_") else f"-D{x}" for x in kv.split(",") if x]
    obj = os.path.join(OUT, f"{src[:-3]}_{name}.o")
    nvcc = b._nvcc()
    subprocess.run([nvcc, *b.ARCH, *[f for f in b.COMMON if f not in ("-Xptxas", "-v")], *defs, "-c", os.path.join(b.CSRC, src), "-o", obj], check=True)
    others = [os.path.join(b.OBJ_DIR, s[:-3] + ".o") for s in b.sources() if s != src]
    so = os.path.join(OUT, f"lib_{name}.so")
    subprocess.run([nvcc, *b.ARCH, "-shared", "-o", so, obj, *others], check=True)
    os.remove(obj)
    return so


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    b.build()
    with ThreadPoolExecutor(8) as ex:
        for so in ex.map(one, sys.argv[1:]):
            print(so)
