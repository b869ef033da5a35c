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


def test_shadow_chunking_arithmetic_is_pure_host():
    """Sizing of the shadow-ray wavefront (csrc/env_shade.cu): the ray list is sized for the covered pixels, visibility bytes
    for all pixels, chunk borders fall on multiples of 16 sample pairs, and int32 ray ids bound a chunk."""
    from gshell_b200 import _lib
    L = _lib.lib
    B, H, W, n = 8, 1024, 1024, 16
    npix, n2 = B * H * W, n * n
    budget = 24 << 30

    def pair(ncov):
        return ncov * 2 * 32 + npix * 2

    # dense views: 24 GB hold 46 pairs -> chunks of 32 -> 8 launches of each kernel
    nb = L.gsb_env_shade_scratch_bytes(B, H, W, npix, n, budget)
    assert nb == pair(npix) * 32 + 256 and nb <= budget
    assert L.gsb_env_shade_chunks(B, H, W, npix, n, nb) == 8
    assert L.gsb_env_shade_chunks(B, H, W, 0, n, nb) == 8                     # 0 = "all pixels"
    # the benchmark coverage (55 %): 80 pairs per chunk -> 4 chunks; a sphere covering 15 %: capped by the int32 ids
    ncov = int(0.55 * npix)
    nb2 = L.gsb_env_shade_scratch_bytes(B, H, W, ncov, n, budget)
    assert nb2 == pair(ncov) * 80 + 256
    assert L.gsb_env_shade_chunks(B, H, W, ncov, n, nb2) == 4
    ncov = int(0.15 * npix)
    id_limit = (1 << 31) // (2 * npix)                                          # 128 pairs at 8.4 M pixels
    nb3 = L.gsb_env_shade_scratch_bytes(B, H, W, ncov, n, budget)
    assert nb3 == pair(ncov) * id_limit + 256
    assert L.gsb_env_shade_chunks(B, H, W, ncov, n, nb3) == 2
    # a larger buffer than asked for only ever reduces the number of chunks; a tiny one still gets the 16-pair minimum
    assert L.gsb_env_shade_chunks(B, H, W, npix, n, budget) <= 8
    tiny = L.gsb_env_shade_scratch_bytes(B, H, W, npix, n, 1 << 20)
    assert tiny == pair(npix) * 16 + 256 and L.gsb_env_shade_chunks(B, H, W, npix, n, tiny) == 16
    # everything fits: one chunk
    assert L.gsb_env_shade_chunks(1, 64, 64, 64 * 64, 8, L.gsb_env_shade_scratch_bytes(1, 64, 64, 64 * 64, 8, budget)) == 1


def test_occluder_sizing_is_pure_host():
    from gshell_b200 import _lib
    L = _lib.lib
    assert L.gsb_occluder_brick_words(217) == 55 ** 3 and L.gsb_occluder_brick_words(4) == 1
    assert L.gsb_occluder_scan_ws_ints(217 ** 3) >= 217 ** 3 // 2048 + 1
    assert L.gsb_occluder_struct_bytes() % 8 == 0
