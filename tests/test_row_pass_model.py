"""Executable specifications of the staged row passes (aloha_b200/csrc/ntt_kernels.cu).  Forward
(`ntt_fwd_rows_tma`):
a warp per 256-coefficient row, 8 coefficients per thread, levels 0-2 / 3-5 / 6-7 with two exchanges through
the padded slot, twiddles read from the row's 256-entry block in `row_slot8` order (kernels.cuh).  The model
moves data and indexes twiddles exactly as the kernel does, thread by thread, and must reproduce the oracle's
transform; it also checks the shared-memory access patterns the kernel relies on being conflict-free.
Inverse (`ntt_inv_rows_tma`): a half-warp per row, 16 coefficients per thread, the row as the SWIZZLE_128B
tensor-map copy leaves it, exchange in place in the swizzled layout, twiddles in `row_slot` order, N^-1
folded into the transform's last stage."""
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


# ------------------------------------------------------------------ inverse staged row pass (ntt_inv_rows_tma)
def row_slot(u, j):           # kernels.cuh: the inverse pass's layout, GS level lt = 7 - u
    if u < 4:
        return (1 << u) + j
    return (1 << u) + (j & ((1 << (u - 4)) - 1)) * 16 + (j >> (u - 4))


def swz(line, chunk):         # SWIZZLE_128B: 16-byte chunk c of 128-byte line l lands at chunk c ^ (l mod 8)
    return 16 * line + 2 * (chunk ^ (line & 7))


def test_inverse_swizzled_accesses_are_conflict_free():
    for c in range(8):        # a thread's 16 contiguous words = line h; LDS.128 / STS.128 of chunk c, 8 lanes per phase
        for half in range(2):
            assert wavefronts([8 * swz(h, c) for h in range(8 * half, 8 * half + 8)], 16) == 1
    for k in range(16):       # after the exchange: word (line k, element h), 16 lanes, 8-byte loads
        assert wavefronts([8 * (swz(k, h >> 1) + (h & 1)) for h in range(16)], 8) == 1


@pytest.mark.parametrize("logn", [8, 10])
def test_inverse_model_matches_oracle(logn):
    n = 1 << logn
    q = O.synthetic_primes(1, 2 * n)[0]
    psi = O.min_primitive_root(q, 2 * n)
    rng = np.random.default_rng(100 + logn)
    x = rng.integers(0, q, n, dtype=np.uint64)
    tabs = O.NttTables(n, [q], [psi])
    want = tabs.batch(x[None].copy(), np.array([0]), inverse=True)[0]
    ipsi = pow(psi, q - 2, q)
    itw = [pow(ipsi, bitrev(j, logn), q) for j in range(n)]
    ninv = pow(n, q - 2, q)
    R = n // 256
    a = [int(v) for v in x]
    for r in range(R):
        rtw = [0] * 256
        for u in range(8):
            for j in range(1 << u):
                rtw[row_slot(u, j)] = itw[(1 << (logn - 8 + u)) + (r << u) + j]
        smem = [0] * 256
        for line in range(16):                     # what the swizzled tensor-map copy leaves in the slot
            for c in range(8):
                for w in range(2):
                    smem[swz(line, c) + w] = a[r * 256 + 16 * line + 2 * c + w]
        regs = [[0] * 16 for _ in range(16)]
        for h in range(16):
            for e in range(16):
                regs[h][e] = smem[swz(h, e >> 1) + (e & 1)]
        for h in range(16):                        # lt = 0..3 pair e-bit lt; twiddle j = (16 h + e) >> (lt + 1)
            X = regs[h]
            for lt in range(4):
                half = 1 << lt
                for e in range(16):
                    if not e & half:
                        g0 = e & ~(2 * half - 1)
                        w = rtw[row_slot(7 - lt, (16 * h + g0) >> (lt + 1))]
                        X[e], X[e + half] = (X[e] + X[e + half]) % q, (X[e] - X[e + half]) * w % q
        for h in range(16):                        # exchange in place, same swizzled words
            for e in range(16):
                smem[swz(h, e >> 1) + (e & 1)] = regs[h][e]
        for h in range(16):
            regs[h] = [smem[swz(k, h >> 1) + (h & 1)] for k in range(16)]
        for h in range(16):                        # lt = 4..7 pair k-bit (lt - 4); twiddle j = k >> (lt - 3)
            X = regs[h]
            for lt in range(4, 8):
                half = 1 << (lt - 4)
                for k in range(16):
                    if not k & half:
                        g0 = k & ~(2 * half - 1)
                        w = rtw[row_slot(7 - lt, g0 >> (lt - 3))]
                        s, d = (X[k] + X[k + half]) % q, (X[k] - X[k + half]) % q
                        if R == 1 and lt == 7:     # last stage of the whole transform: N^-1 folded into both outputs
                            X[k], X[k + half] = s * ninv % q, d * (w * ninv % q) % q
                        else:
                            X[k], X[k + half] = s, d * w % q
        for h in range(16):
            for k in range(16):
                a[r * 256 + h + 16 * k] = regs[h][k]
    # column pass: GS stages on row distance, N^-1 folded into the last one
    for b in range(logn - 8):
        m = 1 << (logn - 8 - 1 - b)
        dist = 256 << b
        for i in range(n):
            if i & dist:
                continue
            w = itw[m + (i >> (8 + b + 1))]
            s, d = (a[i] + a[i + dist]) % q, (a[i] - a[i + dist]) % q
            if b == logn - 9:
                a[i], a[i + dist] = s * ninv % q, d * (w * ninv % q) % q
            else:
                a[i], a[i + dist] = s, d * w % q
    assert (np.array(a, dtype=np.uint64) == want).all()
