"""The thread-block emulator itself (tests/native/cuda_host/block_emulator.h) on kernels with known answers
(tests/native/emulator_selftest.cu, plain CUDA): block scan through shuffles + shared memory + barriers, ballot compaction, dynamic
shared memory, butterfly / any / all / reduce / segmented shuffle, 3-D indices, a coalesced group formed after some lanes left,
and a many-round shared-memory ping-pong.  The product's kernels passing their GPU-validated tests on the emulator
(tests/test_emulated_gpu_suite_cpu.py) is the stronger evidence; this file pins the primitives one at a time, in both thread orders
and with the block order shuffled."""
import ctypes
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def lib():
    native = os.path.join(HERE, "native")
    sys.path.insert(0, native)
    try:
        import host_kernels
    finally:
        sys.path.remove(native)
    so = host_kernels.build(["emulator_selftest.cu"], blocks=True, src_dir=native)
    return so, host_kernels


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.mark.parametrize("order_seed", [0, 1, 6])
def test_primitives(lib, order_seed):
    so, hk = lib
    hk.set_thread_order(so, order_seed)
    rng = np.random.RandomState(order_seed)
    nb = 5
    # block scan
    x = rng.randint(-50, 50, size=nb * 256).astype(np.int32)
    out = np.zeros_like(x)
    assert so.st_block_scan(P(x), P(out), nb) == 0
    assert np.array_equal(out, np.cumsum(x.reshape(nb, 256), axis=1).reshape(-1))
    # compaction
    out = np.full_like(x, -777)
    cnt = np.zeros(nb, dtype=np.int32)
    assert so.st_compact(P(x), P(out), P(cnt), nb) == 0
    for b in range(nb):
        tile = x[b * 256:(b + 1) * 256]
        keep = tile[tile > 0]
        assert cnt[b] == keep.size and np.array_equal(out[b * 256:b * 256 + keep.size], keep)
        assert np.all(out[b * 256 + keep.size:(b + 1) * 256] == -777)
    # dynamic shared memory
    for block in (32, 96, 256):
        f = rng.rand(4 * block).astype(np.float32)
        g = np.zeros_like(f)
        assert so.st_reverse(P(f), P(g), f.size, block) == 0
        assert np.array_equal(g, f.reshape(4, block)[:, ::-1].reshape(-1))
    # warp collectives
    v = rng.randint(-3, 20, size=nb * 64).astype(np.int32)
    v[64:96] = np.abs(v[64:96]) + 1                       # one all-positive warp
    outs = [np.zeros_like(v) for _ in range(5)]
    assert so.st_warp_ops(P(v), *[P(o) for o in outs], nb) == 0
    w = v.reshape(-1, 32)
    assert np.array_equal(outs[0], np.repeat(w.sum(1), 32)) and np.array_equal(outs[3], np.repeat(w.sum(1), 32))
    assert np.array_equal(outs[1], np.repeat((w < 0).any(1).astype(np.int32), 32))
    assert np.array_equal(outs[2], np.repeat((w > 0).all(1).astype(np.int32), 32))
    assert np.array_equal(outs[4], np.repeat(v.reshape(-1, 16)[:, 3], 16))
    assert outs[2].reshape(-1, 32)[2].all() and not outs[2].all()
    # 3-D indices: every (block, thread) pair exactly once
    idx = np.full(12 * 24 * 2, -1, dtype=np.int32)
    assert so.st_indices(P(idx)) == 0
    pairs = idx.reshape(-1, 2)
    assert np.array_equal(pairs[:, 0], np.repeat(np.arange(12), 24)) and np.array_equal(pairs[:, 1], np.tile(np.arange(24), 12))
    # coalesced group after early exits
    c = rng.randint(-4, 9, size=nb * 64).astype(np.int32)
    rank, excl, size = (np.zeros_like(c) for _ in range(3))
    assert so.st_coalesced(P(c), P(rank), P(excl), P(size), nb) == 0
    for wi, lanes in enumerate(c.reshape(-1, 32)):
        live = lanes >= 0
        sl = slice(wi * 32, wi * 32 + 32)
        assert np.all(rank[sl][~live] == -1) and np.all(size[sl][live] == live.sum())
        assert np.array_equal(rank[sl][live], np.arange(live.sum()))
        assert np.array_equal(excl[sl][live], np.cumsum(lanes[live]) - lanes[live])
    # many barriers: after r rounds a[t] = ((t + r) mod 128) + r
    o = np.zeros(nb * 128, dtype=np.int32)
    assert so.st_rounds(P(o), 37, nb) == 0
    assert np.array_equal(o, np.tile((np.arange(128) + 37) % 128 + 37, nb))
    hk.set_thread_order(so, 0)
