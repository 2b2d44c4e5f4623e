"""Parameter helpers for synthetic (N up to 2^16) configurations: the prime rule of SURVEY 8(d)3 and
the twiddle-ROM root the reference uses (minimal primitive 2N-th root of unity,
sim/vp/tf_rom_generator/tf_rom_generator.sv:75 -- verified minimal for the three shipped moduli)."""
from __future__ import annotations


def is_prime(n: int) -> bool:
    if n < 2:
        return False
    small = (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37)
    for p in small:
        if n % p == 0:
            return n == p
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    for a in small:          # deterministic for n < 3.3e24
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def synthetic_primes(count: int, two_n: int, below: int = 1 << 60) -> list[int]:
    """The first `count` primes found scanning downward from `below` with q = 1 (mod 2N)."""
    out, q = [], (below - 1) // two_n * two_n + 1
    while len(out) < count:
        if q < below and is_prime(q):
            out.append(q)
        q -= two_n
    return out


def min_primitive_root(q: int, two_n: int) -> int:
    """Smallest psi with psi^(2N) = 1 and psi^N = -1 (mod q); 0 if 2N does not divide q-1."""
    if two_n < 2 or (q - 1) % two_n:
        return 0
    n = two_n // 2
    root = 0
    for g in range(2, 1000):
        r = pow(g, (q - 1) // two_n, q)
        if pow(r, n, q) == q - 1:
            root = r
            break
    if not root:
        return 0
    sq, cur, best = root * root % q, root, root
    for _ in range(1, n):
        cur = cur * sq % q
        if cur < best:
            best = cur
    return best
