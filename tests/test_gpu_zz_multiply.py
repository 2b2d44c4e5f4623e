"""hks.Multiply (tensor product -> relinearise -> rescale over one scratchpad image) on the engine against the
oracle machine, every output word.

This file sorts after every other GPU test on purpose, and its test is xfail(strict=False): the composite was
written after the round's GPU minutes were spent, so it has never run on a B200.  XPASS in the driver's log
means it ran bit-exact; an xfail would be a finding about the engine, not about the streams (which
tests/test_hks_multiply.py pins on the oracle machine)."""
import numpy as np
import pytest

import aloha_b200 as A
from aloha_b200 import hks
import test_hks as T
import test_hks_multiply as M

pytestmark = pytest.mark.gpu


@pytest.mark.xfail(strict=False, reason="first run on a GPU happens at round end")
@pytest.mark.parametrize("n,L,K,dnum", [(1024, 6, 2, 3), (4096, 4, 1, 4)])
def test_multiply_engine_equals_oracle(n, L, K, dnum):
    prm, psi, a, b, ksk = M.problem(n, L, K, dnum)
    want = M.run_multiply(prm, psi, a, b, ksk)
    rows = hks.Multiply.spm_rows(prm)
    lay = hks.Layout(prm, 1, 0, 1, "relin")
    eng = A.Engine(vlmax_bits=n * 64, spm_rows=rows, ksk_rows=max(lay.ksk_rows, 1), moduli=[(m, psi[m]) for m in prm.moduli],
                   pool_buffers=512, isram_depth=65536)
    mul = hks.Multiply(eng, prm)
    for i in range(L):
        mul.load_input(i, (a[0][i], a[1][i]), (b[0][i], b[1][i]))
    for t in range(L + K):
        mul.load_ksk(t, np.stack([ksk[t][d][c] for d in range(prm.dnum) for c in (0, 1)]))
    for _ in range(2):                          # second pass: cached plans
        mul.run()
    for i, (x, y) in want.items():
        gx, gy = mul.read_output(i)
        assert (gx == x).all() and (gy == y).all(), i
    eng.close()
