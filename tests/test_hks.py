"""Hybrid key-switch / relinearise / rescale stream generator (aloha_b200.hks), host logic on CPU with the
oracle (CPU golden model of the ISA) as the machine:
  * L = 2, K = 1, dnum = 2: bit-exact against the reference's kernel-level rotate vectors (tests/golden) --
    the generator degenerates to sim/vp/isram_file_generator/keyswitch.mem's sequence;
  * general (L, K, dnum): the streams compute the textbook hybrid key switch (restated here with Python
    integers: fast basis extension per digit, inner product with the key, mod-down with rounding);
  * world_size 2 over gloo, with and without the chunked (overlapped) phase 2, equals the one-machine run;
  * batch of two key-switches; relinearise; rescale."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import golden_util as G
from aloha_b200 import hks
from aloha_b200 import params
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def oracle_machine(lay_or_rows, n, moduli_psi, ksk_rows=0):
    rows = lay_or_rows if isinstance(lay_or_rows, int) else lay_or_rows.spm_rows
    ksk = ksk_rows if isinstance(lay_or_rows, int) else lay_or_rows.ksk_rows
    return O.GoldenModel(vlmax_bits=n * 64, spm_rows=rows, ksk_rows=max(ksk, 1), moduli=moduli_psi)


def test_degenerates_to_the_reference_kernel():
    n = G.manifest()["n"]
    items = [i for i in G.manifest()["kernels"] if i["op"] == "rotate"]
    assert items
    prm = hks.Params(n, [O.Q0, O.Q1], [O.Q2])
    assert prm.alpha == 1 and prm.dnum == 2
    for item in items:
        lay = hks.Layout(prm)
        m = oracle_machine(lay, n, [(O.Q0, O.PSI0), (O.Q1, O.PSI1), (O.Q2, O.PSI2)])
        ks = hks.KeySwitch(m, lay)
        ct = G.pool(item["src0"]).reshape(4, n)
        ksk = G.pool(item["ksk"]).reshape(3, 4 * n)          # index = mod*4 + digit*2 + comp (App. B.4)
        for i in range(2):
            ks.load_input(i, (ct[i], ct[2 + i]))
        for t in range(3):
            ks.load_ksk(t, ksk[t])
        ks.run(pow(3, item["step"], 2 * n))
        out = np.concatenate([ks.read_output(0)[0], ks.read_output(1)[0], ks.read_output(0)[1], ks.read_output(1)[1]])
        assert G.poly_hashes(out, np.ones(len(out), bool), n) == item["want"], (item["case"], item["kernel"])


# ---- the textbook computation, Python integers + the oracle's transforms
def synth(n, L, K, seed=3):
    primes = params.synthetic_primes(L + K, 2 * n)
    p, q = primes[:K], primes[K:]                           # the largest primes are the special ones
    psi = {m: params.min_primitive_root(m, 2 * n) for m in primes}
    rng = np.random.default_rng(seed)
    return q, p, psi, rng


def automorph_int(x, k, q):
    return O.automorph(np.asarray(x, dtype=np.uint64), k, q) % np.uint64(q)     # raw q - 0 = q -> 0


def textbook(prm, psi, sw, ksk, k, addends):
    """sw[j]: limb j (NTT form) of the polynomial to switch; ksk[t][b][c]; addends[c][i] or None.
    Returns out[c][i] as object arrays."""
    n, L, K = prm.n, prm.L, prm.K
    obj = lambda a: np.array([int(v) for v in a], dtype=object)
    c = []
    for j in range(L):
        x = O.ntt(np.asarray(sw[j], dtype=np.uint64), prm.q[j], psi[prm.q[j]], inverse=True)
        c.append(obj(automorph_int(x, k, prm.q[j]) if k is not None else x))
    acc = [[None, None] for _ in prm.moduli]
    for t, m in enumerate(prm.moduli):
        for b, g in enumerate(prm.groups):
            if t in g:
                e = c[t]
            else:
                e = sum((c[j] * prm.qhat_inv[j] % prm.q[j]) % m * prm.qhat_mod[j, t] % m for j in g) % m
            E = obj(O.ntt(np.array(e, dtype=np.uint64), m, psi[m]))
            for cc in (0, 1):
                term = E * obj(ksk[t][b][cc]) % m
                acc[t][cc] = term if acc[t][cc] is None else (acc[t][cc] + term) % m
    out = [[None] * L, [None] * L]
    for cc in (0, 1):
        T = []
        for kk, pk in enumerate(prm.p):
            t_ = obj(O.ntt(np.array(acc[L + kk][cc], dtype=np.uint64), pk, psi[pk], inverse=True))
            T.append((t_ + prm.half) % pk * prm.phat_inv[kk] % pk)
        for i, qi in enumerate(prm.q):
            conv = sum(T[kk] % qi * prm.phat_mod[kk, i] % qi for kk in range(K)) % qi
            u = obj(O.ntt(np.array((conv - prm.half) % qi, dtype=np.uint64), qi, psi[qi]))
            r = (acc[i][cc] - u) % qi * prm.pinv[i] % qi
            out[cc][i] = r if addends[cc] is None else (r + obj(addends[cc][i])) % qi
    return out


def make_problem(n, L, K, dnum, kind, seed=3):
    q, p, psi, rng = synth(n, L, K, seed)
    prm = hks.Params(n, q, p, dnum)
    polys = 3 if kind == "relin" else 2
    ct = [[rng.integers(0, qi, n, dtype=np.uint64) for qi in q] for _ in range(polys)]
    ksk = [[[rng.integers(0, m, n, dtype=np.uint64) for _ in (0, 1)] for _ in range(prm.dnum)] for m in prm.moduli]
    return prm, psi, ct, ksk


def run_machine(prm, psi, ct, ksk, k, kind, world=1, rank=0, comm=None, overlap=False, batch=1, only=None):
    lay = hks.Layout(prm, world, rank, batch, kind)
    m = oracle_machine(lay, prm.n, [(mm, psi[mm]) for mm in prm.moduli])
    ks = hks.KeySwitch(m, lay, comm, overlap=overlap)
    for b in range(batch):
        for i in lay.owned():
            if i < prm.L:
                ks.load_input(i, [np.roll(ct[c][i], b) for c in range(len(ct))], b)
    for t in lay.owned():
        ks.load_ksk(t, np.stack([ksk[t][b][c] for b in range(prm.dnum) for c in (0, 1)]))
    ks.run(k if kind == "rotate" else 1, only=only)
    return {(b, i): ks.read_output(i, b) for b in range(batch) for i in lay.owned() if i < prm.L and (only is None or i in only)}


@pytest.mark.parametrize("L,K,dnum", [(4, 1, 4), (4, 2, 2), (6, 2, 3), (5, 3, 2), (6, 1, 1)])
def test_rotate_streams_compute_the_textbook_key_switch(L, K, dnum):
    n = 256
    prm, psi, ct, ksk = make_problem(n, L, K, dnum, "rotate")
    k = pow(3, 5, 2 * n)
    got = run_machine(prm, psi, ct, ksk, k, "rotate")
    a_rot = [np.array(O.ntt(automorph_int(O.ntt(ct[0][i], prm.q[i], psi[prm.q[i]], inverse=True), k, prm.q[i]),
                            prm.q[i], psi[prm.q[i]])) for i in range(L)]
    want = textbook(prm, psi, ct[1], ksk, k, [a_rot, None])
    for i in range(L):
        for c in (0, 1):
            assert [int(v) for v in got[0, i][c]] == list(want[c][i]), (L, K, dnum, i, c)


def test_relinearise_streams():
    n, L, K, dnum = 256, 6, 2, 3
    prm, psi, ct, ksk = make_problem(n, L, K, dnum, "relin")
    got = run_machine(prm, psi, ct, ksk, 1, "relin")
    want = textbook(prm, psi, ct[2], ksk, None, [ct[0], ct[1]])
    for i in range(L):
        for c in (0, 1):
            assert [int(v) for v in got[0, i][c]] == list(want[c][i]), (i, c)


def test_batch_and_subset():
    n, L, K, dnum = 256, 6, 2, 3
    prm, psi, ct, ksk = make_problem(n, L, K, dnum, "rotate")
    k = pow(3, 2, 2 * n)
    one = run_machine(prm, psi, ct, ksk, k, "rotate")
    two = run_machine(prm, psi, ct, ksk, k, "rotate", batch=2)
    rolled = run_machine(prm, psi, [[np.roll(x, 1) for x in poly] for poly in ct], ksk, k, "rotate")
    sub = run_machine(prm, psi, ct, ksk, k, "rotate", only=[1, 4])
    for i in range(L):
        for c in (0, 1):
            assert (two[0, i][c] == one[0, i][c]).all() and (two[1, i][c] == rolled[0, i][c]).all()
    assert sorted(sub) == [(0, 1), (0, 4)]
    for key, v in sub.items():
        assert (v[0] == one[key][0]).all() and (v[1] == one[key][1]).all()


def test_rescale_streams():
    n, L = 256, 4
    q, _, psi, rng = synth(n, L, 0)
    ct = [[rng.integers(0, qi, n, dtype=np.uint64) for qi in q] for _ in range(2)]
    m = oracle_machine(8 * L * n // 128, n, [(mm, psi[mm]) for mm in q])
    rs = hks.Rescale(m, n, q)
    for i in range(L):
        rs.load_input(i, ct[0][i], ct[1][i])
    rs.run()
    ql = q[-1]
    obj = lambda a: np.array([int(v) for v in a], dtype=object)
    for c in (0, 1):
        t = (obj(O.ntt(ct[c][L - 1], ql, psi[ql], inverse=True)) + ql // 2) % ql
        for i in range(L - 1):
            qi = q[i]
            u = obj(O.ntt(np.array((t - ql // 2) % qi, dtype=np.uint64), qi, psi[qi]))
            want = (obj(ct[c][i]) - u) % qi * pow(ql, -1, qi) % qi
            assert [int(v) for v in rs.read_output(i)[c]] == list(want), (c, i)


# ---- world_size 2 over gloo
def _worker(rank, world, shape, overlap, port, qout):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, L, K, dnum = shape
        prm, psi, ct, ksk = make_problem(n, L, K, dnum, "rotate")
        out = run_machine(prm, psi, ct, ksk, pow(3, 2, 2 * n), "rotate", world, rank, hks.TorchComm(), overlap=overlap)
        qout.put((rank, {i: (x.tolist(), y.tolist()) for (b, i), (x, y) in out.items()}))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("overlap,port", [(False, 29541), ("chunks", 29542), ("own", 29543)])
def test_world2_gloo_equals_single_machine(overlap, port):
    shape = (256, 6, 2, 3)
    world = 2
    n, L, K, dnum = shape
    prm, psi, ct, ksk = make_problem(n, L, K, dnum, "rotate")
    single = run_machine(prm, psi, ct, ksk, pow(3, 2, 2 * n), "rotate")
    ctx = mp.get_context("spawn")
    qout = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, shape, overlap, port, qout)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(qout.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    seen = set()
    for rank, out in results.items():
        for i, (x, y) in out.items():
            assert (np.array(x, dtype=np.uint64) == single[0, i][0]).all(), (rank, i)
            assert (np.array(y, dtype=np.uint64) == single[0, i][1]).all(), (rank, i)
            seen.add(i)
    assert seen == set(range(L))


@pytest.mark.parametrize("L,K,dnum", [(47, 1, 47), (40, 8, 5)])
@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("overlap", [False, "own", "chunks"])
def test_every_rank_issues_the_same_collectives(L, K, dnum, world, overlap):
    """The config-5 shapes at 2 / 4 / 8 machines: whatever a rank owns (at 8 machines one of them holds nothing but
    special primes), its op list carries the same transfers in the same order as every other rank's, at least
    three run ops (phase 1, phase 2 chunk(s), phase 3), and a wait before anything that reads transferred rows."""
    n = 256
    primes = params.synthetic_primes(L + K, 2 * n)
    prm = hks.Params(n, primes[K:], primes[:K], dnum)
    seqs = []
    for r in range(world):
        lay = hks.Layout(prm, world, r, batch=2)
        ks = hks.KeySwitch(hks.Recorder(), lay, type("C", (), {"world": world, "rank": r})(), overlap=overlap)
        prog = ks.program(pow(3, 5, 2 * n))
        seqs.append([op for op in prog if op[0] in ("all_gather", "broadcast")])
        runs = [op for op in prog if op[0] == "run"]
        assert len(runs) >= 3, (r, len(runs))
        kinds = [op[0] for op in prog]
        assert kinds[0] == "run" and kinds[1] == "all_gather" and kinds[-1] == "run" and kinds[-2] == "wait"
        # every limb of every batch element is covered exactly once by phase 1 and by phase 3 over the ranks
        # (+ 1: the retire stream that closes every non-empty batch; a rank without ciphertext limbs has empty batches)
        limbs = 2 * len([t for t in lay.owned() if t < L])
        assert len(runs[0][1]) == len(runs[-1][1]) == (limbs + 1 if limbs else 0)
    assert all(s == seqs[0] for s in seqs), "ranks disagree on the transfers"
    assert len(seqs[0]) >= 2


def test_layout_limits_and_counts():
    n = 65536
    q = list(range(40))
    prm = hks.Params.__new__(hks.Params)       # shapes only: no number theory on fake moduli
    prm.n, prm.rp, prm.L, prm.K, prm.alpha, prm.dnum = n, n // 128, 40, 8, 8, 5
    lay = hks.Layout(prm, world=8, rank=3)
    assert lay.per_rank == 6 and lay.slots == 48
    assert sorted(sum((lay.owned(r) for r in range(8)), [])) == list(range(48))
    assert lay.OUT_size <= 65536 and lay.ACC_size <= 65536
    assert hks.Params.transform_count(prm) == 120 + 5 * 48 + 16 + 80
    prm.L, prm.K, prm.alpha, prm.dnum = 70, 1, 1, 70
    with pytest.raises(ValueError):
        hks.Layout(prm)
