"""The integer steps of the pseudo-Mersenne product and fold (aloha_b200/csrc/modarith.cuh: mul_pm,
mul_pm_parts, fold_pm) restated with Python integers: congruence and the range claims the kernels' bound
tracking relies on, at the extremes of every operand.  (The identities do not need q to be prime, so the
largest admissible d = 2^27 is tested directly.)"""
import random

import pytest

from oracle import oracle as O

M64 = (1 << 64) - 1


def mul_pm(y, w, q):
    d = (1 << 60) - q
    w2 = (w << 32) % q
    yl, yh = y & 0xFFFFFFFF, y >> 32
    t = w * yl + w2 * yh                     # two 60 x 32-bit halves
    assert t < 1 << 93
    rh, rl = t >> 61, t & ((1 << 61) - 1)
    assert rh <= (1 << 32) - 2               # fits the 32-bit multiplier operand
    p, l = rh * (2 * d), rl                  # the two pieces the forward butterfly adds separately
    assert p < 1 << 61 and l < 1 << 61
    return p + l


def fold_pm(x, q):
    d = (1 << 60) - q
    return (x & ((1 << 60) - 1)) + (x >> 60) * d


def moduli():
    qs = O.synthetic_primes(3, 1 << 17) + [(1 << 60) - (1 << 27), (1 << 60) - 1, (1 << 60) - (1 << 27) + 1]
    return qs


@pytest.mark.parametrize("q", moduli())
def test_product_is_congruent_and_below_3q(q):
    rng = random.Random(q)
    ys = [0, 1, M64, M64 - 1, 1 << 63, (1 << 32) - 1, 1 << 32, 16 * q - 1, q, q - 1]
    ws = [0, 1, q - 1, q - 2, (1 << 59) + 12345, (1 << 32) - 1, 1 << 32]
    ys += [rng.getrandbits(64) for _ in range(400)]
    ws += [rng.randrange(q) for _ in range(40)]
    for w in ws:
        for y in ys:
            r = mul_pm(y, w, q)
            assert r % q == w * y % q
            assert r < 3 * q, (hex(w), hex(y), hex(r))
            assert r <= M64


@pytest.mark.parametrize("q", moduli())
def test_fold_is_congruent_and_below_2q(q):
    rng = random.Random(q + 1)
    for x in [0, 1, q - 1, q, 2 * q, 16 * q - 1, M64, (1 << 60) - 1, 1 << 60] + [rng.getrandbits(64) for _ in range(2000)]:
        r = fold_pm(x, q)
        assert r % q == x % q and r < 2 * q
        # canonical = one conditional subtract
        assert (r - q if r >= q else r) == x % q


def test_forward_bounds_never_overflow_64_bits():
    """Forward stage: x' = x + t, y' = x + 3q - t with t < 3q; the kernels fold an upper input back below 2q
    whenever the next stage would pass 16q (Arith<FORM_PM>: GROW = 3, REDB = 2) -- the running bound stays
    within 16q < 2^64 for any number of stages, and the lower input may be ANY 64-bit word."""
    q = (1 << 60) - 1
    b = 2
    for _ in range(64):
        if b + 3 > 16:
            b = 2
        b += 3
        assert b <= 16 and b * q <= M64
