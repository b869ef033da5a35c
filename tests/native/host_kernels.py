"""TEST INFRASTRUCTURE: compile "independent thread per element" translation units of gshell_b200/csrc for the CPU.

    lib = host_kernels.build(["mesh_ops.cu", "tangents.cu"])      -> ctypes.CDLL exporting the same `gsb_*` entry points

The kernel source is used as it is: the only rewrite is textual, `k<<<grid, block, smem, stream>>>(args);` ->
`gsb_host::launch(grid, block, [&] { k(args); });`, and <cuda_runtime.h> resolves to tests/native/cuda_host/cuda_runtime.h.
Compiled with -ffp-contract=off (the separately rounded operations of the -fmad=false units stay separately rounded).
`sanitize=True` adds AddressSanitizer (run the test process with LD_PRELOAD=$(gcc -print-file-name=libasan.so))."""
import ctypes
import fcntl
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


def _split_top_level(text):
    """comma-separated pieces of `text`, commas inside brackets not counted"""
    parts, depth, cur = [], 0, []
    for ch in text:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append("".join(cur).strip())
            cur = []
        else:
            cur.append(ch)
    parts.append("".join(cur).strip())
    return parts


# Device helpers that are inline PTX on the GPU and have a one-line meaning on the host.  The bulk copy (cp.async.bulk + mbarrier,
# csrc/mt_extract.cu) is issued by one thread BEFORE the block's barrier and waited for after it: on the host the copy happens at
# issue time and the waits are empty.  The cache-hinted loads of csrc/trace_core.cuh are plain loads.
HOST_BODIES = {
    "smem_u32": "{ (void)p; return 0u; }",
    "mbar_init": "{ (void)bar; (void)count; }",
    "mbar_fence_init": "{ }",
    "mbar_expect_tx": "{ (void)bar; (void)bytes; }",
    "bulk_load": "{ (void)bar; std::memcpy(dst_smem, src_gmem, bytes); }",
    "mbar_wait": "{ (void)bar; (void)phase; }",
}


def rewrite_ptx_helpers(src):
    for name, body in HOST_BODIES.items():
        m = re.search(r"__device__\s+__forceinline__\s+[\w:<> ]+?[\s*&]" + name + r"\s*\(", src)
        if not m:
            continue
        args_end = _balanced(src, src.index("(", m.start()), "(", ")")
        body_start = src.index("{", args_end)
        body_end = _balanced(src, body_start, "{", "}")
        src = src[:body_start] + body + src[body_end:]
    return src


def rewrite_launches(src):
    out, pos = [], 0
    for m in re.finditer(r"([A-Za-z_]\w*(?:<[^<>;(){}]*>)?)\s*<<<", src):
        if m.start() < pos:
            continue
        cfg_end = src.index(">>>", m.end())
        cfg = _split_top_level(src[m.end():cfg_end])
        arg_start = src.index("(", cfg_end)
        arg_end = _balanced(src, arg_start, "(", ")")
        out.append(src[pos:m.start()])
        smem = cfg[2] if len(cfg) > 2 and cfg[2] else "0"
        out.append(f"gsb_host::launch({cfg[0]}, {cfg[1]}, {smem}, [&] {{ {m.group(1)}{src[arg_start:arg_end]}; }})")
        pos = arg_end
    out.append(src[pos:])
    return rewrite_ptx_helpers(rewrite_dynamic_shared(rewrite_vector_reductions("".join(out))))


def rewrite_dynamic_shared(src):
    """`extern __shared__ T name[];` -> `T* name = (T*)gsb_host::dynamic_shared();` (the per-launch buffer of block_emulator.h)"""
    return re.sub(r"extern\s+__shared__\s+([\w:]+)\s+(\w+)\s*\[\s*\]\s*;", r"\1* \2 = (\1*)gsb_host::dynamic_shared();", src)


def rewrite_vector_reductions(src):
    """`asm volatile("red.global.add.vN.f32 [%0], {...};" ::"l"(p), "f"(a), ... : "memory");` -> gsb_host::red_add(p, a, ...);
    (the one kind of inline PTX in the units compiled here: a vector reduction to global memory)"""
    def repl(m):
        ops = re.findall(r'"[lf]"\(((?:[^()]|\([^()]*\))*)\)', m.group(1))
        return "gsb_host::red_add(" + ", ".join(ops) + ")"
    return re.sub(r'asm\s+volatile\(\s*"red\.global\.add\.v[24]\.f32[^"]*"\s*::((?:[^;]|\n)*?):\s*"memory"\s*\)', repl, src)


def build(units, sanitize=False, blocks=False, src_dir=None, defines=()):
    """`src_dir`: where the units live (default: the product's gshell_b200/csrc)"""
    src_dir = src_dir or CSRC
    key = (tuple(units), sanitize, blocks, src_dir, tuple(defines))
    if key in _cache:
        return _cache[key]
    texts = [rewrite_launches(open(os.path.join(src_dir, u)).read()) for u in units]
    headers = [os.path.join(CSRC, h) for h in sorted(os.listdir(CSRC)) if h.endswith((".cuh", ".h"))]
    headers += [os.path.join(ROOT, "include", "gshell_b200.h")]
    headers += [os.path.join(dp, f) for dp, _, fs in sorted(os.walk(os.path.join(HERE, "cuda_host"))) for f in sorted(fs)]
    tag = hashlib.sha1(("".join(texts) + "".join(open(h).read() for h in headers) + str(sanitize) + str(blocks)
                        + " ".join(defines)).encode()).hexdigest()[:16]
    work = os.path.join(tempfile.gettempdir(), f"gsb_host_kernels_{tag}")
    os.makedirs(work, exist_ok=True)
    so = os.path.join(work, "libgsb_host_kernels.so")
    # several processes (xdist workers, the ranks of a torchrun) may ask for the same build at once: one builds, the others wait
    lock = open(os.path.join(work, ".lock"), "w")
    fcntl.flock(lock, fcntl.LOCK_EX)
    try:
        _compile(so, units, texts, work, sanitize, blocks, defines)
    finally:
        fcntl.flock(lock, fcntl.LOCK_UN)
        lock.close()
    lib = ctypes.CDLL(so)
    _cache[key] = lib
    return lib


def _compile(so, units, texts, work, sanitize, blocks, defines):
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
               *(["-fsanitize=address", "-fno-omit-frame-pointer"] if sanitize else []), *(["-DGSB_HOST_BLOCKS"] if blocks else []), *[f"-D{d}" for d in defines],
               *srcs, seed, "-o", so + f".tmp{os.getpid()}"]
        subprocess.run(cmd, check=True, capture_output=True, text=True)
        os.replace(so + f".tmp{os.getpid()}", so)


def set_thread_order(lib, seed):
    ctypes.c_uint.in_dll(lib, "gsb_host_thread_order_seed").value = int(seed)
