"""Executable specification of the staged forward row pass (`ntt_fwd_rows_tma`, aloha_b200/csrc/ntt_kernels.cu):
a warp per 256-coefficient row, 8 coefficients per thread, levels 0-2 / 3-5 / 6-7 with two exchanges through
the padded slot, twiddles read from the row's 256-entry block in `row_slot8` order (kernels.cuh).  The model
moves data and indexes twiddles exactly as the kernel does, thread by thread, and must reproduce the oracle's
transform; it also checks the shared-memory access patterns the kernel relies on being conflict-free."""
import numpy as np
import pytest

from oracle import oracle as O


def row_slot8(u, j):          # kernels.cuh
    if u < 4:
        return (1 << u) + j
    if u == 4:
        return 16 + (j & 1) * 8 + (j >> 1)
    if u == 5:
        return 32 + (j & 3) * 8 + (j >> 2)
    if u == 6:
        return 64 + (j & 1) * 32 + (j >> 1)
    return 128 + (j & 3) * 32 + (j >> 2)


def pos(j):                   # padded slot: word jj at jj + 2 (jj >> 4)
    return j + 2 * (j >> 4)


def bitrev(v, bits):
    r = 0
    for _ in range(bits):
        r = (r << 1) | (v & 1)
        v >>= 1
    return r


def test_row_slot8_is_a_bijection_onto_1_255():
    seen = {row_slot8(u, j) for u in range(8) for j in range(1 << u)}
    assert seen == set(range(1, 256))


def wavefronts(addresses_bytes, width):
    """Shared-memory wavefronts of one warp access: 32 banks x 4 B; distinct 128-byte-wide bank rows that the
    access touches per bank, maximised over banks (same address = broadcast)."""
    per_bank = {}
    for a in addresses_bytes:
        for w in range(a // 4, (a + width) // 4):
            per_bank.setdefault(w % 32, set()).add(w)
    return max(len(v) for v in per_bank.values())


def test_exchange_patterns_are_conflict_free():
    lanes = range(32)
    for k in range(8):       # exchange 1 writes: element lane + 32 k            (8-byte accesses: 2 wavefronts ideal)
        assert wavefronts([8 * pos(l + 32 * k) for l in lanes], 8) == 2
    for m in range(8):       # exchange 1 reads / exchange 2 writes: hi*32 + m*4 + lo
        assert wavefronts([8 * pos((l >> 2) * 32 + m * 4 + (l & 3)) for l in lanes], 8) == 2
    for e in range(0, 8, 2):  # exchange 2 reads: 16 bytes at 8 lane + e, a quarter-warp per phase
        for quarter in range(4):
            ls = range(8 * quarter, 8 * quarter + 8)
            assert wavefronts([8 * pos(8 * l + e) for l in ls], 16) == 1


@pytest.mark.parametrize("logn", [8, 10])
def test_model_matches_oracle(logn):
    n = 1 << logn
    q = O.synthetic_primes(1, 2 * n)[0]
    psi = O.min_primitive_root(q, 2 * n)
    rng = np.random.default_rng(logn)
    x = rng.integers(0, q, n, dtype=np.uint64)
    want = O.NttTables(n, [q], [psi]).batch(x[None].copy(), np.array([0]))[0]
    tw = [pow(psi, bitrev(j, logn), q) for j in range(n)]
    R = n // 256

    def bf(a, b, w):
        t = w * b % q
        return (a + t) % q, (a - t) % q

    # column pass (plain CT stages on rows; the kernels' column pass computes exactly this)
    a = [int(v) for v in x]
    for s in range(logn - 8):
        half = n >> (s + 1)
        for blk in range(1 << s):
            w = tw[(1 << s) + blk]
            for i in range(half):
                lo = blk * 2 * half + i
                a[lo], a[lo + half] = bf(a[lo], a[lo + half], w)
    out = [0] * n
    for r in range(R):
        rr = R + r
        rtw = [0] * 256
        for u in range(8):
            for j in range(1 << u):
                rtw[row_slot8(u, j)] = tw[(rr << u) + j]
        row = a[r * 256:(r + 1) * 256]
        slot = [0] * 288
        regs = [[row[t + 32 * k] for k in range(8)] for t in range(32)]
        for t in range(32):                       # phase A: levels 0..2 pair k-bit (2-u)
            X = regs[t]
            for u in range(3):
                half = 4 >> u
                for k in range(8):
                    if not k & half:
                        X[k], X[k + half] = bf(X[k], X[k + half], rtw[row_slot8(u, (k & ~(2 * half - 1)) >> (3 - u))])
        for t in range(32):
            for k in range(8):
                slot[pos(t + 32 * k)] = regs[t][k]
        for t in range(32):                       # lane = hi*4 + lo reads hi*32 + m*4 + lo
            hi, lo = t >> 2, t & 3
            regs[t] = [slot[pos(hi * 32 + m * 4 + lo)] for m in range(8)]
        for t in range(32):                       # phase B: levels 3..5 pair m-bit (2-v)
            hi, X = t >> 2, regs[t]
            for v in range(3):
                half = 4 >> v
                for m in range(8):
                    if not m & half:
                        g0 = m & ~(2 * half - 1)
                        X[m], X[m + half] = bf(X[m], X[m + half], rtw[row_slot8(3 + v, (hi << v) + (g0 >> (3 - v)))])
        for t in range(32):
            hi, lo = t >> 2, t & 3
            for m in range(8):
                slot[pos(hi * 32 + m * 4 + lo)] = regs[t][m]
        for t in range(32):
            regs[t] = [slot[pos(8 * t + e)] for e in range(8)]
        for t in range(32):                       # phase C: levels 6..7 pair e-bit (1-v)
            X = regs[t]
            for v in range(2):
                half = 2 >> v
                for e in range(8):
                    if not e & half:
                        g0 = e & ~(2 * half - 1)
                        X[e], X[e + half] = bf(X[e], X[e + half], rtw[row_slot8(6 + v, (t << (v + 1)) + (g0 >> (2 - v)))])
        for t in range(32):
            for e in range(8):
                out[r * 256 + 8 * t + e] = regs[t][e]
    assert (np.array(out, dtype=np.uint64) == want).all()
