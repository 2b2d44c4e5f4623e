"""Random key-switch shapes through the batcher on the simulated device (tests/sim_engine.py): L, K, dnum, batch,
rotate / relinearise, output-limb subsets, engine flags and pool sizes drawn at random, every output word against
the oracle machine, each shape run twice (the second time on cached plans); likewise the multiply chain.
18 000 shapes of this sweep have run clean; the suite keeps a sample."""
import os
import random

import numpy as np
import pytest

import sim_engine
from aloha_b200 import hks
import test_hks as T
import test_hks_multiply as M


def keyswitch_case(A, seed):
    rng = random.Random(seed)
    flagsets = [0, A.F_DEFER, A.F_NO_FUSE, A.F_GRAPHS, A.F_DEFER | A.F_GRAPHS, A.F_AUT_GATHER, A.F_AUT_TILED, A.F_GENERIC_MODMUL, A.F_STRICT]
    n = rng.choice([256, 256, 512])
    L, K = rng.randrange(1, 10), rng.randrange(1, 5)
    dnum = rng.randrange(1, L + 1)
    kind, batch = rng.choice(["rotate", "relin"]), rng.randrange(1, 4)
    flags, pool = rng.choice(flagsets), rng.choice([64, 128, 512])
    only = None if rng.random() < 0.7 else sorted(rng.sample(range(L), rng.randrange(1, L + 1)))
    prm, psi, ct, ksk = T.make_problem(n, L, K, dnum, kind, seed=seed)
    k = pow(3, rng.randrange(1, 50), 2 * n)
    want = T.run_machine(prm, psi, ct, ksk, k, kind, batch=batch, only=only)
    lay = hks.Layout(prm, 1, 0, batch, kind)
    eng = A.Engine(vlmax_bits=n * 64, spm_rows=lay.spm_rows, ksk_rows=max(lay.ksk_rows, 1), moduli=[(m, psi[m]) for m in prm.moduli],
                   pool_buffers=pool, isram_depth=65536, flags=flags)
    ks = hks.KeySwitch(eng, lay)
    for b in range(batch):
        for i in range(prm.L):
            ks.load_input(i, [np.roll(ct[c][i], b) for c in range(len(ct))], b)
    for t in range(prm.L + prm.K):
        ks.load_ksk(t, np.stack([ksk[t][d][c] for d in range(prm.dnum) for c in (0, 1)]))
    for _ in range(2):
        ks.run(k if kind == "rotate" else 1, only=only)
    for (b, i), (x, y) in want.items():
        gx, gy = ks.read_output(i, b)
        assert (gx == x).all() and (gy == y).all(), (seed, (n, L, K, dnum, kind, batch, hex(flags), pool, only), b, i)
    eng.close()


def multiply_case(A, seed):
    rng = random.Random(seed)
    n, L, K = 256, rng.randrange(2, 8), rng.randrange(1, 4)
    dnum = rng.randrange(1, L + 1)
    flags = rng.choice([0, A.F_DEFER, A.F_NO_FUSE, A.F_GRAPHS, A.F_STRICT])
    prm, psi, a, b, ksk = M.problem(n, L, K, dnum, seed=seed)
    want = M.run_multiply(prm, psi, a, b, ksk)
    lay = hks.Layout(prm, 1, 0, 1, "relin")
    eng = A.Engine(vlmax_bits=n * 64, spm_rows=hks.Multiply.spm_rows(prm), ksk_rows=max(lay.ksk_rows, 1),
                   moduli=[(m, psi[m]) for m in prm.moduli], pool_buffers=rng.choice([64, 256]), isram_depth=65536, flags=flags)
    mul = hks.Multiply(eng, prm)
    for i in range(L):
        mul.load_input(i, (a[0][i], a[1][i]), (b[0][i], b[1][i]))
    for t in range(L + K):
        mul.load_ksk(t, np.stack([ksk[t][d][c] for d in range(prm.dnum) for c in (0, 1)]))
    for _ in range(2):
        mul.run()
    for i, (x, y) in want.items():
        gx, gy = mul.read_output(i)
        assert (gx == x).all() and (gy == y).all(), (seed, (L, K, dnum, hex(flags)), i)
    eng.close()


@pytest.mark.parametrize("block", range(4))
def test_key_switch_shapes(block):
    per = int(os.environ.get("ALOHA_SWEEP_SHAPES", "15"))
    with sim_engine.simulated() as A:
        for seed in range(block * per, (block + 1) * per):
            keyswitch_case(A, seed)


def test_multiply_shapes():
    with sim_engine.simulated() as A:
        for seed in range(int(os.environ.get("ALOHA_SWEEP_SHAPES", "15"))):
            multiply_case(A, seed)
