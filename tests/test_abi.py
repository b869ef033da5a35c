"""The C-ABI library builds, loads and exports every symbol include/gshell_b200.h declares
(no compute calls: runs without a GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "gshell_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(gsb_\w+)\s*\(", txt)))


def test_library_builds_and_exports_all_symbols():
    from gshell_b200 import build
    out, _ = build.build()
    assert os.path.exists(out)
    lib = ctypes.CDLL(out)
    names = _declared()
    assert len(names) >= 6
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/gshell_b200.h but not exported"


def test_python_binding_covers_header():
    from gshell_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()
    assert _lib.lib.gsb_abi_version() >= 1
    assert _lib.lib.gsb_compiled_arch() == 100


def test_workspace_query_is_pure_host():
    from gshell_b200 import _lib
    small = _lib.lib.gsb_mt_workspace_bytes(1000, 1500)
    big = _lib.lib.gsb_mt_workspace_bytes(12985416, 15000000)
    assert 0 < small < big < 2 ** 31
