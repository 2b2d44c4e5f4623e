"""CPU replay of the tiled VAUT kernels' index arithmetic (aloha_b200/csrc/aut_plan.hpp, the header the
CUDA kernels include): the (point, offset) tiles cover Z_n exactly once for every odd Galois element, the
result equals the reference's dst[(i*k) mod N] = +-src[i] (vxu_lane.sv:594-599), and both sides of the
permutation move in nearly whole 32-byte sectors per warp instruction."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "native", "aut_plan_model.cpp")
HDR = os.path.join(HERE, "..", "aloha_b200", "csrc", "aut_plan.hpp")
LIB = os.path.join(HERE, "native", "libaut_plan_model.so")


@pytest.fixture(scope="module")
def model():
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", LIB, SRC])
    return C.CDLL(LIB)


def apply(model, n, k, q=(1 << 60) - 93, seed=1, fn="aut_model_apply"):
    src = np.random.default_rng(seed).integers(0, q, n, dtype=np.uint64)
    src[:4] = 0                                    # 0 -> q on the negated half
    dst = np.full(n, 0xDEAD, dtype=np.uint64)
    st = (C.c_uint64 * 8)()
    rc = getattr(model, fn)(C.c_uint32(n), C.c_uint64(k), C.c_uint64(q), src.ctypes.data_as(C.c_void_p),
                            dst.ctypes.data_as(C.c_void_p), st)
    assert rc == 0, (n, k, rc)
    i = np.arange(n, dtype=np.uint64)
    d = (i * np.uint64(k)) % np.uint64(n)
    neg = ((i * np.uint64(k)) % np.uint64(2 * n)) >= n
    want = np.zeros(n, dtype=np.uint64)
    want[d] = np.where(neg, np.uint64(q) - src, src)
    return bool((dst == want).all()), list(st)


@pytest.mark.parametrize("n", [256, 1024])
def test_every_odd_k_is_an_exact_cover(model, n):
    for k in range(1, 2 * n, 2):
        ok, st = apply(model, n, k)
        assert ok and st[0] == n and st[1] == 0, (n, k, st)
        assert st[4] <= 2048 + 1024


def test_full_size_galois_elements(model):
    n = 65536
    rng = np.random.default_rng(5)
    ks = [pow(3, s, 2 * n) for s in (1, 2, 4, 8, 16, 1000, n // 8, n // 4)] + [2 * n - 1, n + 1, n // 2 + 1, 43691, 21845, 5, 7]
    ks += [int(x) | 1 for x in rng.integers(1, 2 * n, 40)]
    worst = 0.0
    for k in ks:
        ok, st = apply(model, n, k)
        assert ok and st[0] == n and st[1] == 0, (k, st)
        src_ratio, dst_ratio = st[2] / (n / 4), st[3] / (n / 4)
        worst = max(worst, src_ratio, dst_ratio)
        # the planner trades a little sector efficiency for full warps; the design's claim: at most 1.6x the
        # whole-sector traffic on either side and at most 1.5x the ideal number of warp steps
        assert src_ratio <= 1.6 and dst_ratio <= 1.6, (k, src_ratio, dst_ratio)
        assert st[7] <= 1.5 * (2 * n / 32), (k, "warp steps", st[7])
        assert st[5] <= 2 and st[6] <= 2, (k, "shared-memory bank conflicts", st[5], st[6])
    assert worst >= 1.0


def test_python_oracle_agrees(model):
    """the same permutation through oracle.automorph (the C++ golden model's VAUT)"""
    from oracle import oracle as O
    n, q = 8192, O.Q0
    x = np.random.default_rng(3).integers(0, q, n, dtype=np.uint64)
    for k in (3, 9, 6561, 2 * n - 1, 8191):
        dst = np.zeros(n, dtype=np.uint64)
        st = (C.c_uint64 * 8)()
        assert model.aut_model_apply(C.c_uint32(n), C.c_uint64(k), C.c_uint64(q), x.ctypes.data_as(C.c_void_p),
                                     dst.ctypes.data_as(C.c_void_p), st) == 0
        assert (dst == O.automorph(x, k, q)).all(), k
