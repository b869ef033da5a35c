"""`__graft_entry__.smoke()` -- the first thing the driver runs on the B200 -- on the host build of the kernels: the function's own
source with its device literal replaced and its build step dropped (tests/test_emulated_gpu_suite_cpu.py runs this in a process
of its own with GSB_HOST_EMULATION=1)."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)


def main():
    assert os.environ.get("GSB_HOST_EMULATION") == "1"
    import conftest
    conftest.bind_host_library()
    path = os.path.join(ROOT, "__graft_entry__.py")
    src = open(path).read()
    cut = src.index("def smoke()")
    body = src[cut:]
    for old, new in (("    build()\n", ""), ('assert torch.cuda.is_available(), "smoke() needs cuda:0"', ""), ('"cuda:0"', '"cpu"')):
        assert old in body, old
        body = body.replace(old, new)
    ns = {"__file__": path, "__name__": "graft_entry_on_host"}
    exec(compile(src[:cut], path, "exec"), ns)
    exec(compile(body, path, "exec"), ns)
    ns["smoke"]()
    print("SMOKE_ON_HOST_OK")


if __name__ == "__main__":
    main()
