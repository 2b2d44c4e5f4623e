"""Ciphertext x ciphertext as instruction streams (aloha_b200.hks.Multiply: tensor product, relinearise,
rescale over one scratchpad image) on the oracle machine:
  * bit-exact against the same computation written with Python integers (test_hks.textbook + the rescale formula);
  * world_size 2 over gloo equals the one-machine run;
  * it IS a multiplication: two RLWE ciphertexts are multiplied under a freshly made relinearisation key and the
    result decrypts, with the original secret, to the product of the messages divided by the dropped prime."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from aloha_b200 import hks
from oracle import oracle as O
import test_hks as T
from test_hks_rlwe import ntt, polymul

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
obj = lambda a: np.array([int(v) for v in a], dtype=object)


def problem(n, L, K, dnum, seed=11):
    q, p, psi, rng = T.synth(n, L, K, seed)
    prm = hks.Params(n, q, p, dnum)
    a = [[rng.integers(0, qi, n, dtype=np.uint64) for qi in q] for _ in range(2)]
    b = [[rng.integers(0, qi, n, dtype=np.uint64) for qi in q] for _ in range(2)]
    ksk = [[[rng.integers(0, m, n, dtype=np.uint64) for _ in (0, 1)] for _ in range(prm.dnum)] for m in prm.moduli]
    return prm, psi, a, b, ksk


def run_multiply(prm, psi, a, b, ksk, world=1, rank=0, comm=None, overlap=False, rescale=True):
    rows = hks.Multiply.spm_rows(prm, world, rescale)
    lay = hks.Layout(prm, world, rank, 1, "relin")
    m = T.oracle_machine(rows, prm.n, [(mm, psi[mm]) for mm in prm.moduli], ksk_rows=lay.ksk_rows)
    mul = hks.Multiply(m, prm, world, rank, comm, overlap=overlap, rescale=rescale)
    assert mul.rs is None or mul.rs.spm_rows == rows
    for i in mul.lay.owned():
        if i < prm.L:
            mul.load_input(i, (a[0][i], a[1][i]), (b[0][i], b[1][i]))
    for t in mul.lay.owned():
        mul.load_ksk(t, np.stack([ksk[t][d][c] for d in range(prm.dnum) for c in (0, 1)]))
    mul.run()
    last = prm.L - 1 if rescale else prm.L
    return {i: mul.read_output(i) for i in mul.lay.owned() if i < last}


def textbook_multiply(prm, psi, a, b, ksk, rescale=True):
    L, q = prm.L, prm.q
    d0 = [obj(a[0][i]) * obj(b[0][i]) % q[i] for i in range(L)]
    d1 = [(obj(a[0][i]) * obj(b[1][i]) + obj(a[1][i]) * obj(b[0][i])) % q[i] for i in range(L)]
    d2 = [np.array(obj(a[1][i]) * obj(b[1][i]) % q[i], dtype=np.uint64) for i in range(L)]
    ct = T.textbook(prm, psi, d2, ksk, None, [d0, d1])
    if not rescale:
        return ct
    ql = q[-1]
    out = [[None] * (L - 1), [None] * (L - 1)]
    for c in (0, 1):
        t = (obj(O.ntt(np.array(ct[c][L - 1], dtype=np.uint64), ql, psi[ql], inverse=True)) + ql // 2) % ql
        for i in range(L - 1):
            u = obj(O.ntt(np.array((t - ql // 2) % q[i], dtype=np.uint64), q[i], psi[q[i]]))
            out[c][i] = (ct[c][i] - u) % q[i] * pow(ql, -1, q[i]) % q[i]
    return out


@pytest.mark.parametrize("L,K,dnum,rescale", [(4, 2, 2, True), (6, 2, 3, True), (5, 1, 5, True), (4, 2, 2, False)])
def test_multiply_streams_compute_the_textbook_product(L, K, dnum, rescale):
    n = 256
    prm, psi, a, b, ksk = problem(n, L, K, dnum)
    got = run_multiply(prm, psi, a, b, ksk, rescale=rescale)
    want = textbook_multiply(prm, psi, a, b, ksk, rescale)
    assert sorted(got) == list(range(L - 1 if rescale else L))
    for i, (x, y) in got.items():
        assert [int(v) for v in x] == list(want[0][i]) and [int(v) for v in y] == list(want[1][i]), (L, K, dnum, i)


def test_tensor_stream_obeys_the_register_banks():
    """every vv instruction of the tensor stream names one even and one odd source register (expander.v:183-200):
    asm.Program asserts it while the stream is built"""
    prm, *_ = problem(256, 4, 1, 4)
    words = hks.tensor_stream(prm, 2).words()
    assert len(words) == 1 + 2 + 4 + 2 + 4 + 2 + 1      # vsetvl, vsetq + vsetiq, 4 loads, d0, d1, d2, break


# ---- world_size 2 over gloo
def _worker(rank, world, shape, overlap, port, qout):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        prm, psi, a, b, ksk = problem(*shape)
        out = run_multiply(prm, psi, a, b, ksk, world, rank, hks.TorchComm(), overlap=overlap)
        qout.put((rank, {i: (x.tolist(), y.tolist()) for i, (x, y) in out.items()}))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("overlap,port", [(False, 29561), ("own", 29562)])
def test_world2_gloo_equals_single_machine(overlap, port):
    shape = (256, 6, 2, 3)
    prm, psi, a, b, ksk = problem(*shape)
    single = run_multiply(prm, psi, a, b, ksk)
    ctx = mp.get_context("spawn")
    qout = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, shape, overlap, port, qout)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(qout.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    seen = set()
    for rank, out in results.items():
        for i, (x, y) in out.items():
            assert (np.array(x, dtype=np.uint64) == single[i][0]).all() and (np.array(y, dtype=np.uint64) == single[i][1]).all(), (rank, i)
            seen.add(i)
    assert seen == set(range(prm.L - 1))


# ---- it is a multiplication
@pytest.mark.parametrize("L,K,dnum", [(4, 2, 2), (5, 1, 5)])
def test_product_decrypts_to_the_product_of_the_messages(L, K, dnum):
    n = 256
    q, p, psi, _ = T.synth(n, L, K)
    prm = hks.Params(n, q, p, dnum)
    rng = np.random.default_rng(77)
    small = lambda: np.array([int(v) for v in rng.integers(-4, 5, n)], dtype=object)
    s = np.array([int(v) for v in rng.integers(-1, 2, n)], dtype=object)
    Q = 1
    for qi in q:
        Q *= qi

    def negacyclic(x, y):                       # over the integers
        r = [0] * n
        for i in range(n):
            xi = int(x[i])
            if xi:
                for j in range(n):
                    k = i + j
                    if k < n:
                        r[k] += xi * int(y[j])
                    else:
                        r[k - n] -= xi * int(y[j])
        return np.array(r, dtype=object)

    def encrypt(msg):
        c1 = np.array([int.from_bytes(rng.bytes(48), "little") % Q for _ in range(n)], dtype=object)
        e = small()
        c0s, c1s = [], []
        for qi in q:
            c0_i = (msg + e - polymul(c1 % qi, s, qi, psi[qi])) % qi
            c0s.append(np.array(ntt(c0_i, qi, psi[qi]), dtype=np.uint64))
            c1s.append(np.array(ntt(c1 % qi, qi, psi[qi]), dtype=np.uint64))
        return [c0s, c1s]
    m1 = np.array([int(v) << 40 for v in rng.integers(-(1 << 10), 1 << 10, n)], dtype=object)
    m2 = np.array([int(v) << 40 for v in rng.integers(-(1 << 10), 1 << 10, n)], dtype=object)
    a, b = encrypt(m1), encrypt(m2)
    # relinearisation key: s^2 -> s, digit b's gadget = P on the limbs of its own group
    s2 = negacyclic(s, s)
    key_err = [small() for _ in prm.groups]
    ksk = []
    for t, mt in enumerate(prm.moduli):
        per_digit = []
        for d, g in enumerate(prm.groups):
            r = np.array([int(v) for v in rng.integers(0, mt, n, dtype=np.uint64)], dtype=object)
            gadget = prm.P % mt if t in g else 0
            k0 = (key_err[d] - polymul(r, s, mt, psi[mt]) + gadget * (s2 % mt)) % mt
            per_digit.append([np.array(ntt(k0, mt, psi[mt]), dtype=np.uint64), np.array(ntt(r, mt, psi[mt]), dtype=np.uint64)])
        ksk.append(per_digit)
    out = run_multiply(prm, psi, a, b, ksk)
    want = negacyclic(m1, m2)                    # exact integer product of the messages
    ql = q[-1]
    noises = []
    for i in range(L - 1):
        qi = q[i]
        o0, o1 = out[i]
        dec = (ntt(o0, qi, psi[qi], inverse=True) + polymul(ntt(o1, qi, psi[qi], inverse=True), s, qi, psi[qi])) % qi
        # dec = (m1 m2 + noise) / ql rounded, an integer far below q_i: compare with the exact quotient
        centred = [int(v) if int(v) < qi // 2 else int(v) - qi for v in dec]
        noise = [c - (2 * int(w) + ql) // (2 * ql) for c, w in zip(centred, want)]
        noises.append(noise)
        assert max(abs(v) for v in noise) < 1 << 12, (L, K, dnum, i, max(abs(v) for v in noise).bit_length())
    # the noise is ONE small integer polynomial, seen under every modulus
    assert all(nz == noises[0] for nz in noises)
