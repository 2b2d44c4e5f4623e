"""Differential fuzzing of the batcher: random (legal) instruction streams, several run_vp calls per
machine so register state, aliases and CSR context carry over, executed on the GPU engine under
different flag sets and on the oracle; the whole SPM image and its written-mask must agree.

The generator keeps to streams with defined RTL behaviour (vv operands in different banks, vd != vs1 on
VNTT / VINTT / VAUT / VROLI, odd Galois elements) and, for the fast path, never reads a register that a
transform has clobbered (ALOHA_F_STRICT runs lift that restriction: there the clobbered source holds the
RTL's intermediate and is read on purpose)."""
import os
import random

import numpy as np
import pytest

import aloha_b200 as A
from aloha_b200 import asm
from oracle import oracle as O

pytestmark = pytest.mark.gpu

N = 256
RP = N // 128
SLOTS = 24                       # polynomial slots in SPM
PRIMES = [O.Q0, O.Q1]
PSIS = [pow(O.PSI0, 8192 // N, O.Q0), pow(O.PSI1, 8192 // N, O.Q1)]


def gen_program(rng: random.Random, defined: set, strict: bool, n_ops: int):
    """-> (Program, csr tuple).  `defined` is updated in place (registers holding a defined value)."""
    p = asm.Program()
    bases = [rng.randrange(0, SLOTS // 3) * RP for _ in range(3)]       # src0, src1, rslt row pointers
    step = rng.randrange(0, 64)
    p.vsetvl(N)
    q = rng.choice(PRIMES)
    p.vsetq(q)

    def slot_for(base_sel):
        return rng.randrange(0, SLOTS - bases[base_sel] // RP) * RP

    for _ in range(n_ops):
        kind = rng.choice(["vle", "vle", "vse", "vv", "vv", "vs", "mod", "aut", "roli", "ntt", "intt", "setq"])
        live = sorted(defined)
        if kind == "setq":
            q = rng.choice(PRIMES)
            p.vsetq(q)
        elif kind == "vle" or not live:
            vd, b = rng.randrange(32), rng.randrange(3)
            p.vle(vd, b, slot_for(b))
            defined.add(vd)
        elif kind == "vse":
            b = rng.randrange(3)
            p.vse(rng.choice(live), b, slot_for(b))
        elif kind == "vv":
            vs1 = rng.choice(live)
            partners = [r for r in live if (r ^ vs1) & 1]
            if not partners:
                continue
            vd = rng.randrange(32)
            getattr(p, rng.choice(["vfqmul", "vfqadd", "vfqsub"]))(vd, vs1, rng.choice(partners))
            defined.add(vd)
        elif kind == "vs":
            vd, vs1 = rng.randrange(32), rng.choice(live)
            # raw 64-bit scalars can push a value past 2q, outside the fast transforms' domain
            imm = rng.getrandbits(64) if (strict and rng.random() < 0.3) else rng.randrange(q)
            if rng.random() < 0.2:
                p.vfqsub_sv(vd, imm, vs1)
            else:
                getattr(p, rng.choice(["vfqmul", "vfqadd", "vfqsub"]))(vd, vs1, imm=imm)
            defined.add(vd)
        elif kind == "mod":
            vd, vs1 = rng.randrange(32), rng.choice(live)
            (p.vfqmod if rng.random() < 0.5 else p.vcpy)(vd, vs1)
            defined.add(vd)
        elif kind in ("aut", "roli", "ntt", "intt"):
            vs1 = rng.choice(live)
            vd = rng.choice([r for r in range(32) if r != vs1])
            if kind == "aut":
                if not strict and q != max(PRIMES):
                    continue      # q - x wraps for x > q (a residue of the larger prime): garbage beyond 2q
                imm = rng.randrange(0, 64)
                if (step + imm) % 2 == 0:
                    imm += 1
                p.vaut(vd, vs1, imm)
            elif kind == "roli":
                p.vroli(vd, vs1, rng.randrange(0, 4 * N))
            else:
                (p.vntt if kind == "ntt" else p.vintt)(vd, vs1)
                if not strict:
                    defined.discard(vs1)            # fast path: the source becomes undefined
            defined.add(vd)
    # make the register file observable: store a few registers at the end
    for r in rng.sample(sorted(defined), min(4, len(defined))):
        p.vse(r, 2, slot_for(2))
    return p.brk(), (bases[0], bases[1], bases[2], 0, step)


def run_case(seed: int, strict: bool, flags: int):
    rng = random.Random(seed)
    data_rng = np.random.default_rng(seed)
    spm0 = data_rng.integers(0, PRIMES[0], SLOTS * N, dtype=np.uint64)     # canonical under both primes
    machines = [O.GoldenModel(vlmax_bits=N * 64, spm_rows=SLOTS * RP, ksk_rows=0, moduli=list(zip(PRIMES, PSIS))),
                A.Engine(vlmax_bits=N * 64, spm_rows=SLOTS * RP, ksk_rows=0, moduli=list(zip(PRIMES, PSIS)), flags=flags)]
    defined: set = set()
    programs = []
    for call in range(3):
        prog, csr = gen_program(rng, defined, strict, rng.randrange(8, 40))
        programs.append((prog, csr))
    images = []
    for m in machines:
        m.dma_mem_h2d(0, spm0[: (SLOTS - 4) * N])          # the last four slots stay never-written ('x')
        pc = 0
        for prog, csr in programs:
            m.load_isram(prog.words(), pc)
            m.run_vp(pc, *csr)
            pc += len(prog)
        images.append((m.dma_mem_d2h(0, SLOTS * N), m.spm_written(0, SLOTS * N)))
    (gd, gw), (ed, ew) = images
    assert (gw == ew).all(), f"seed {seed}: written-mask differs"
    bad = np.nonzero(gd != ed)[0]
    assert bad.size == 0, f"seed {seed}: {bad.size} words differ, first in slot {bad[0] // N}"


@pytest.mark.parametrize("seed", range(int(os.environ.get("ALOHA_FUZZ_SEEDS", "24"))))
def test_fuzz_fast_path(seed):
    run_case(seed, strict=False, flags=[0, A.F_DEFER, A.F_NO_FUSE, A.F_GRAPHS][seed % 4])


@pytest.mark.parametrize("seed", range(100000, 100000 + int(os.environ.get("ALOHA_FUZZ_SEEDS", "24")) // 2))
def test_fuzz_strict(seed):
    run_case(seed, strict=True, flags=A.F_STRICT | (A.F_DEFER if seed % 2 else 0))


@pytest.mark.parametrize("seed", range(8))
def test_fuzz_batched_transforms(seed):
    """Random batch shapes through the transform launches: N, the number of polynomials (so the split between
    the TMA-staged row passes, which take runs of 16 per modulus, and the plain ones varies), a random mix of
    pseudo-Mersenne and generic moduli, inputs with q-1 runs and words in [q, 2q); forward then inverse, every
    word against the oracle."""
    rng = np.random.default_rng(4200 + seed)
    n = 1 << int(rng.integers(8, 13))
    rp = n // 128
    B = int(rng.integers(1, 41))
    n_pm, n_gen = int(rng.integers(0, 3)), int(rng.integers(0, 3))
    if n_pm + n_gen == 0:
        n_pm = 1
    primes = O.synthetic_primes(n_pm, 2 * n) + O.synthetic_primes(n_gen, 2 * n, below=(1 << 60) - (1 << 40))
    order = rng.permutation(len(primes))
    primes = [primes[i] for i in order]
    psis = [O.min_primitive_root(q, 2 * n) for q in primes]
    L = len(primes)
    rows = B * L * rp
    flags = [0, A.F_GRAPHS, A.F_DEFER, A.F_GENERIC_MODMUL][seed % 4]
    eng = A.Engine(vlmax_bits=n * 64, spm_rows=2 * rows, ksk_rows=0, moduli=list(zip(primes, psis)), flags=flags)
    eng.load_isram(asm.transform_stream(n, primes).words(), 0)
    eng.load_isram(asm.transform_stream(n, primes, inverse=True).words(), 1024)
    qv = np.array(primes, dtype=np.uint64)[None, :, None]
    x = rng.integers(0, 1 << 59, (B, L, n), dtype=np.uint64) % qv
    x[0, :, : n // 2] = qv[0, :, :] - np.uint64(1)
    x[B - 1, :, 1::2] += qv[0, :, :]                      # [q, 2q)
    eng.dma_mem_h2d(0, x.reshape(-1))
    eng.run_vp_batch(0, [(b * L * rp, 0, rows + b * L * rp, 0, 0) for b in range(B)])
    F = eng.dma_mem_d2h(rows, B * L * n).reshape(B, L, n)
    tabs = O.NttTables(n, primes, psis)
    for b in range(B):
        assert (F[b] == tabs.batch(x[b].copy(), np.arange(L))).all(), (seed, n, B, b)
    eng.run_vp_batch(1024, [(rows + b * L * rp, 0, b * L * rp, 0, 0) for b in range(B)])
    back = eng.dma_mem_d2h(0, B * L * n).reshape(B, L, n)
    assert (back == x % qv).all(), (seed, n, B)
