"""TEST INFRASTRUCTURE: compile "independent thread per element" translation units of gshell_b200/csrc for the CPU.

    lib = host_kernels.build(["mesh_ops.cu", "tangents.cu"])      -> ctypes.CDLL exporting the same `gsb_*` entry points

The kernel source is used as it is: the only rewrite is textual, `k<<<grid, block, smem, stream>>>(args);` ->
`gsb_host::launch(grid, block, [&] { k(args); });`, and <cuda_runtime.h> resolves to tests/native/cuda_host/cuda_runtime.h.
Compiled with -ffp-contract=off (the separately rounded operations of the -fmad=false units stay separately rounded).
`sanitize=True` adds AddressSanitizer (run the test process with LD_PRELOAD=$(gcc -print-file-name=libasan.so))."""
import ctypes
import hashlib
import os
import re
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "gshell_b200", "csrc")
_cache = {}


def _balanced(text, start, open_ch, close_ch):
    """index just past the bracket that closes the one at `start`"""
    depth = 0
    for i in range(start, len(text)):
        if text[i] == open_ch:
            depth += 1
        elif text[i] == close_ch:
            depth -= 1
            if depth == 0:
                return i + 1
    raise ValueError("unbalanced launch expression")


def rewrite_launches(src):
    out, pos = [], 0
    for m in re.finditer(r"([A-Za-z_]\w*(?:<[^<>;(){}]*>)?)\s*<<<", src):
        if m.start() < pos:
            continue
        cfg_end = src.index(">>>", m.end())
        cfg = [c.strip() for c in src[m.end():cfg_end].split(",")]
        arg_start = src.index("(", cfg_end)
        arg_end = _balanced(src, arg_start, "(", ")")
        out.append(src[pos:m.start()])
        out.append(f"gsb_host::launch({cfg[0]}, {cfg[1]}, [&] {{ {m.group(1)}{src[arg_start:arg_end]}; }})")
        pos = arg_end
    out.append(src[pos:])
    return "".join(out)


def build(units, sanitize=False):
    key = (tuple(units), sanitize)
    if key in _cache:
        return _cache[key]
    texts = [rewrite_launches(open(os.path.join(CSRC, u)).read()) for u in units]
    tag = hashlib.sha1(("".join(texts) + open(os.path.join(HERE, "cuda_host", "cuda_runtime.h")).read()
                        + str(sanitize)).encode()).hexdigest()[:16]
    work = os.path.join(tempfile.gettempdir(), f"gsb_host_kernels_{tag}")
    os.makedirs(work, exist_ok=True)
    so = os.path.join(work, "libgsb_host_kernels.so")
    if not os.path.exists(so):
        srcs = []
        for u, t in zip(units, texts):
            # the units include their neighbours relatively ("vec.cuh", "../../include/gshell_b200.h"): keep that layout
            d = os.path.join(work, "gshell_b200", "csrc")
            os.makedirs(d, exist_ok=True)
            p = os.path.join(d, u[:-3] + "_host.cpp")
            open(p, "w").write(t)
            srcs.append(p)
        for h in os.listdir(CSRC):
            if h.endswith((".cuh", ".h")):
                open(os.path.join(work, "gshell_b200", "csrc", h), "w").write(open(os.path.join(CSRC, h)).read())
        os.makedirs(os.path.join(work, "include"), exist_ok=True)
        open(os.path.join(work, "include", "gshell_b200.h"), "w").write(open(os.path.join(ROOT, "include", "gshell_b200.h")).read())
        seed = os.path.join(work, "seed.cpp")
        open(seed, "w").write('extern "C" { unsigned gsb_host_thread_order_seed = 0; }\n')
        cmd = ["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-I", os.path.join(HERE, "cuda_host"),
               *(["-fsanitize=address", "-fno-omit-frame-pointer"] if sanitize else []), *srcs, seed, "-o", so + ".tmp"]
        subprocess.run(cmd, check=True, capture_output=True, text=True)
        os.replace(so + ".tmp", so)
    lib = ctypes.CDLL(so)
    _cache[key] = lib
    return lib


def set_thread_order(lib, seed):
    ctypes.c_uint.in_dll(lib, "gsb_host_thread_order_seed").value = int(seed)
