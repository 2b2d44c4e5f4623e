"""The generated hybrid key-switch streams are a VALID key switch, not just a self-consistent instruction
sequence: an RLWE ciphertext is rotated by the streams (on the oracle machine, CPU) under a freshly made
key-switching key for (dnum digits, K special primes), decrypted with the original secret, and the plaintext
comes out as the automorphism of the message up to noise far below the modulus.

Conventions (the reference's keyswitch.mem): ct = (c0, c1) with c0 + c1*s = m + e; the stream
outputs (aut(c0) + r0, r1) where (r0, r1) key-switches aut(c1) from aut(s) to s.  Key of digit b under modulus t:
(k0, k1) = (-a*s + e + g_b*aut(s), a) with gadget g_b = P * Qhat_b * [Qhat_b^-1 mod Q_b], i.e. P mod q_t on the
limbs of the digit's own group and 0 elsewhere."""
import numpy as np
import pytest

from aloha_b200 import hks, params
from oracle import oracle as O
import test_hks as T


def ntt(x, q, psi, inverse=False):
    return np.array([int(v) for v in O.ntt(np.array([int(v) % q for v in x], dtype=np.uint64), q, psi, inverse=inverse)], dtype=object)


def polymul(a, b, q, psi):
    return ntt(ntt(a, q, psi) * ntt(b, q, psi) % q, q, psi, inverse=True)


def aut(x, k, q):
    return np.array([int(v) % q for v in O.automorph(np.array([int(v) % q for v in x], dtype=np.uint64), k, q)], dtype=object)


@pytest.mark.parametrize("L,K,dnum", [(4, 2, 2), (6, 3, 2), (4, 1, 4)])
def test_rotation_decrypts_to_the_rotated_message(L, K, dnum):
    n = 256
    q, p, psi, _ = T.synth(n, L, K)
    prm = hks.Params(n, q, p, dnum)
    rng = np.random.default_rng(2024)
    k = pow(3, 7, 2 * n)
    small = lambda: np.array([int(v) for v in rng.integers(-4, 5, n)], dtype=object)
    s = np.array([int(v) for v in rng.integers(-1, 2, n)], dtype=object)
    m = np.array([int(v) << 30 for v in rng.integers(0, 1 << 12, n)], dtype=object)
    Q = 1
    for qi in q:
        Q *= qi
    c1 = np.array([int.from_bytes(rng.bytes(64), "little") % Q for _ in range(n)], dtype=object)
    e = small()
    # ciphertext limbs in evaluation (NTT) form, as the streams expect them
    ct0, ct1 = [], []
    for qi in q:
        c0_i = (m + e - polymul(c1 % qi, s, qi, psi[qi])) % qi
        ct0.append(np.array(ntt(c0_i, qi, psi[qi]), dtype=np.uint64))
        ct1.append(np.array(ntt(c1 % qi, qi, psi[qi]), dtype=np.uint64))
    # key-switching key from aut_k(s) to s; a digit's error is ONE small polynomial, seen under every modulus
    key_err = [small() for _ in prm.groups]
    ksk = []
    for t, mt in enumerate(prm.moduli):
        s_rot = aut(s, k, mt)
        per_digit = []
        for b, g in enumerate(prm.groups):
            a = np.array([int(v) for v in rng.integers(0, mt, n, dtype=np.uint64)], dtype=object)
            gadget = prm.P % mt if t in g else 0
            k0 = (key_err[b] - polymul(a, s, mt, psi[mt]) + gadget * s_rot) % mt
            per_digit.append([np.array(ntt(k0, mt, psi[mt]), dtype=np.uint64), np.array(ntt(a, mt, psi[mt]), dtype=np.uint64)])
        ksk.append(per_digit)
    out = T.run_machine(prm, psi, [ct0, ct1], ksk, k, "rotate")
    for i, qi in enumerate(q):
        o0, o1 = out[0, i]
        dec = (ntt(o0, qi, psi[qi], inverse=True) + polymul(ntt(o1, qi, psi[qi], inverse=True), s, qi, psi[qi])) % qi
        diff = (dec - aut(m, k, qi)) % qi
        noise = max(min(int(v), qi - int(v)) for v in diff)
        assert noise < 1 << 24, (L, K, dnum, i, noise.bit_length())
