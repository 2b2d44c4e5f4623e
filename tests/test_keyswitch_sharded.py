"""The limb-sharded key-switch stream (aloha_b200.keyswitch), host logic on CPU:
  * L = 2, one machine: bit-exact against the reference's kernel-level rotate vectors (tests/golden);
  * L = 5 synthetic, world_size 2 over gloo: every rank's output limbs equal the one-machine run.
The machine here is the oracle (CPU golden model) -- the GPU engine runs the same code in
tests/test_gpu_parity.py."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import golden_util as G
from aloha_b200 import keyswitch as KS
from aloha_b200 import params
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make_machine(lay, moduli_psi, factory=None):
    if factory is not None:
        return factory(lay, moduli_psi)
    return O.GoldenModel(vlmax_bits=lay.n * 64, spm_rows=lay.spm_rows, ksk_rows=lay.ksk_rows,
                         moduli=moduli_psi)


def run_reference_rotate_vector(item, factory=None):
    n = G.manifest()["n"]
    lay = KS.KeySwitchLayout(n, [O.Q0, O.Q1], O.Q2)
    m = make_machine(lay, [(O.Q0, O.PSI0), (O.Q1, O.PSI1), (O.Q2, O.PSI2)], factory)
    ks = KS.ShardedKeySwitch(m, lay)
    ct = G.pool(item["src0"]).reshape(4, n)
    ksk = G.pool(item["ksk"]).reshape(3, 4 * n)          # index = mod*4 + digit*2 + comp (App. B.4)
    for i in range(2):
        ks.load_input(i, ct[i], ct[2 + i])
    for i in range(3):
        ks.load_ksk(i, ksk[i])
    ks.run(pow(3, item["step"], 2 * n))
    out = np.concatenate([ks.read_output(0)[0], ks.read_output(1)[0], ks.read_output(0)[1], ks.read_output(1)[1]])
    return G.poly_hashes(out, np.ones(len(out), bool), n), item["want"]


def test_generalised_stream_reproduces_reference_rotate_vectors():
    items = [i for i in G.manifest()["kernels"] if i["op"] == "rotate"]
    assert items
    for item in items:
        got, want = run_reference_rotate_vector(item)
        assert got == want, (item["case"], item["kernel"])


def synth_problem(n, L, seed=3):
    primes = params.synthetic_primes(L + 1, 2 * n)
    P, q = primes[0], primes[1:]                           # the largest prime is the special one
    psi = {p: params.min_primitive_root(p, 2 * n) for p in primes}
    rng = np.random.default_rng(seed)
    a = [rng.integers(0, qi, n, dtype=np.uint64) for qi in q]
    b = [rng.integers(0, qi, n, dtype=np.uint64) for qi in q]
    ksk = [np.stack([rng.integers(0, (q + [P])[i], n, dtype=np.uint64) for _ in range(2 * L)]) for i in range(L + 1)]
    return q, P, psi, a, b, ksk


def run_sharded(n, L, world, rank, comm, factory=None):
    q, P, psi, a, b, ksk = synth_problem(n, L)
    lay = KS.KeySwitchLayout(n, q, P, world, rank)
    m = make_machine(lay, [(p, psi[p]) for p in q + [P]], factory)
    ks = KS.ShardedKeySwitch(m, lay, comm)
    for i in lay.owned():
        if i < L:
            ks.load_input(i, a[i], b[i])
        ks.load_ksk(i, ksk[i])
    ks.run(pow(3, 2, 2 * n))
    return {i: ks.read_output(i) for i in lay.owned() if i < L}


def _worker(rank, world, n, L, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        out = run_sharded(n, L, world, rank, KS.TorchComm())
        q.put((rank, {i: (x.tolist(), y.tolist()) for i, (x, y) in out.items()}))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_equals_single_machine():
    n, L, world = 256, 5, 2
    single = run_sharded(n, L, 1, 0, KS.LocalComm())
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, n, L, 29533, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    seen = set()
    for rank, out in results.items():
        for i, (x, y) in out.items():
            assert (np.array(x, dtype=np.uint64) == single[i][0]).all(), (rank, i)
            assert (np.array(y, dtype=np.uint64) == single[i][1]).all(), (rank, i)
            seen.add(i)
    assert seen == set(range(L))


def test_layout_partition_and_counts():
    lay = KS.KeySwitchLayout(65536, list(range(47)), 99, world=8, rank=3)
    assert lay.per_rank == 6 and lay.slots == 48
    assert sorted(sum((lay.owned(r) for r in range(8)), [])) == list(range(48))
    assert lay.owner(47) == 7 and KS.transform_count(47) == 3 * 47 + 47 * 48 + 2 + 94
    with pytest.raises(ValueError):
        KS.KeySwitchLayout(65536, list(range(70)), 99)
