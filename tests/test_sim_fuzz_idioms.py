"""Differential fuzzing of the batcher on the simulated device (tests/sim_engine.py) with programs made of the
IDIOMS its fusion passes look for -- multiply-accumulate chains, rotate-multiply-accumulate, base-extension sums,
load-op + transform -- instead of uniformly random instructions (tests/test_gpu_fuzz.py, whose streams almost never
fuse), with registers and scratchpad slots reused aggressively, the smallest renaming pool, and several calls per
plan (run_vp_batch / run_vp_multi / the deferred queue) whose row ranges overlap.  The oracle executes the same
calls one after the other; the whole scratchpad image and its written-mask must agree.

The simulated kernels evaluate the RTL arithmetic on every input word, so unlike the GPU fuzz there is no input
domain to respect; the one rule kept is that a register clobbered by a fast-path transform is not read again."""
import os
import random

import numpy as np
import pytest

import sim_engine
from aloha_b200 import asm
from oracle import oracle as O

N = 256
RP = N // 128
SLOTS = 20
KSK_SLOTS = 4
PRIMES = [O.Q0, O.Q1]
PSIS = [pow(O.PSI0, 8192 // N, O.Q0), pow(O.PSI1, 8192 // N, O.Q1)]


class Gen:
    def __init__(self, rng: random.Random, defined: set, strict: bool):
        self.rng, self.defined, self.strict = rng, defined, strict
        self.p = asm.Program()
        self.bases = [rng.randrange(0, SLOTS // 3) * RP for _ in range(3)]
        self.step = rng.randrange(0, 64)
        self.q = rng.choice(PRIMES)
        self.p.vsetvl(N).vsetq(self.q)

    def slot(self, b):
        return self.rng.randrange(0, SLOTS - self.bases[b] // RP) * RP

    def reg(self, parity=None, avoid=()):
        c = [r for r in range(32) if (parity is None or r & 1 == parity) and r not in avoid]
        return self.rng.choice(c)

    def have(self, parity=None, avoid=()):
        """a defined register of the given bank, loading one if there is none (or sometimes anyway)"""
        c = [r for r in sorted(self.defined) if (parity is None or r & 1 == parity) and r not in avoid]
        if c and self.rng.random() < 0.7:
            return self.rng.choice(c)
        r = self.reg(parity, avoid)
        if self.rng.random() < 0.15:
            self.p.vle(r, asm.BASE_KSK, self.rng.randrange(KSK_SLOTS) * RP)      # key memory: read-only for the VP
        else:
            b = self.rng.randrange(3)
            self.p.vle(r, b, self.slot(b))
        self.defined.add(r)
        return r

    def define(self, r):
        self.defined.add(r)
        return r

    def maybe_store(self, r, prob=0.5):
        if self.rng.random() < prob:
            b = self.rng.randrange(3)
            self.p.vse(r, b, self.slot(b))

    # ---- idioms
    def mac_chain(self):
        terms = self.rng.randrange(2, 6)
        acc_par = self.rng.randrange(2)
        acc = prod = None
        for t in range(terms):
            a = self.have(0)
            b = self.have(1, avoid=(a,))
            if t == 0:
                acc = self.define(self.reg(acc_par))
                self.p.vfqmul(acc, a, b)
            else:
                prod = self.define(self.reg(1 - acc_par, avoid=(acc,)))
                self.p.vfqmul(prod, a, b)
                nxt = acc if self.rng.random() < 0.7 else self.define(self.reg(acc_par, avoid=(prod,)))
                self.p.vfqadd(nxt, acc, prod)
                acc = nxt
        self.maybe_store(acc, 0.8)

    def rotate_mac(self):
        x = self.have()
        t = self.define(self.reg(avoid=(x,)))
        imm = self.rng.randrange(0, 64)
        if (self.step + imm) % 2 == 0:
            imm += 1
        self.p.vaut(t, x, imm)
        pk = self.have(1 - (t & 1), avoid=(t,))
        m = self.define(self.reg())
        self.p.vfqmul(m, t, pk)
        c = self.have(1 - (m & 1), avoid=(m,))
        out = c if self.rng.random() < 0.5 else self.define(self.reg())
        self.p.vfqadd(out, c, m)
        self.maybe_store(out, 0.8)

    def base_extension(self):
        terms = self.rng.randrange(1, 6)
        total = None
        for t in range(terms):
            x = self.have()
            pre = self.rng.choice(["vcpy", "vfqmod", None])
            e = x
            if pre:
                e = self.define(self.reg(avoid=(x,)))
                getattr(self.p, pre)(e, x)
            s = self.rng.randrange(self.q) if self.rng.random() < 0.9 else self.rng.getrandbits(64)
            m = self.define(self.reg(0 if total is None else 1 - (total & 1), avoid=(e,) if total is None else (e, total)))
            self.p.vfqmul(m, e, imm=s)
            if total is None:
                total = m
            else:
                nxt = total if self.rng.random() < 0.7 else self.define(self.reg(total & 1, avoid=(m,)))
                self.p.vfqadd(nxt, total, m)
                total = nxt
        if self.rng.random() < 0.5:
            out = self.define(self.reg())
            self.p.vfqsub(out, total, imm=self.rng.randrange(self.q))
            total = out
        if self.rng.random() < 0.5:
            y = self.define(self.reg(avoid=(total,)))
            (self.p.vntt if self.rng.random() < 0.7 else self.p.vintt)(y, total)
            if not self.strict:
                self.defined.discard(total)
            total = y
        self.maybe_store(total, 0.8)

    def load_op_transform(self):
        x = self.have()
        e = self.define(self.reg(avoid=(x,)))
        (self.p.vfqmod if self.rng.random() < 0.5 else self.p.vcpy)(e, x)
        y = self.define(self.reg(avoid=(e,)))
        (self.p.vntt if self.rng.random() < 0.5 else self.p.vintt)(y, e)
        if not self.strict:
            self.defined.discard(e)
        self.maybe_store(y, 0.7)

    def in_place_memory(self):
        """load, compute, store back over the row that was loaded -- the aliased source must be moved first"""
        b = self.rng.randrange(3)
        slot = self.slot(b)
        r = self.define(self.reg())
        self.p.vle(r, b, slot)
        kind = self.rng.choice(["aut", "roli", "ntt", "mul", "add"])
        out = self.define(self.reg(avoid=(r,)))
        if kind == "aut":
            imm = self.rng.randrange(0, 64)
            if (self.step + imm) % 2 == 0:
                imm += 1
            self.p.vaut(out, r, imm)
        elif kind == "roli":
            self.p.vroli(out, r, self.rng.randrange(0, 4 * N))
        elif kind == "ntt":
            (self.p.vntt if self.rng.random() < 0.5 else self.p.vintt)(out, r)
            if not self.strict:
                self.defined.discard(r)
        elif kind == "mul":
            self.p.vfqmul(out, r, imm=self.rng.randrange(self.q))
        else:
            o = self.have(1 - (r & 1), avoid=(r, out))
            self.p.vfqadd(out, r, o)
        self.p.vse(out, b, slot)

    def single(self):
        kind = self.rng.choice(["setq", "vle", "vse", "roli", "sub", "subsv"])
        if kind == "setq":
            self.q = self.rng.choice(PRIMES)
            self.p.vsetq(self.q)
        elif kind == "vle":
            self.have()
        elif kind == "vse" and self.defined:
            self.maybe_store(self.rng.choice(sorted(self.defined)), 1.0)
        elif kind == "roli":
            x = self.have()
            self.p.vroli(self.define(self.reg(avoid=(x,))), x, self.rng.randrange(0, 4 * N))
        elif kind == "sub":
            a = self.have(0)
            b = self.have(1, avoid=(a,))
            self.p.vfqsub(self.define(self.reg()), a, b)
        elif kind == "subsv":
            x = self.have()
            self.p.vfqsub_sv(self.define(self.reg()), self.rng.randrange(self.q), x)

    def program(self, n_idioms):
        table = [self.mac_chain, self.rotate_mac, self.base_extension, self.load_op_transform, self.in_place_memory,
                 self.single, self.single]
        for _ in range(n_idioms):
            self.rng.choice(table)()
        for r in self.rng.sample(sorted(self.defined), min(3, len(self.defined))):
            self.p.vse(r, 2, self.slot(2))
        return self.p.brk(), (self.bases[0], self.bases[1], self.bases[2], 0, self.step)


def run_case(A, seed: int):
    rng = random.Random(seed)
    strict = seed % 5 == 4
    flag_sets = [0, A.F_DEFER, A.F_NO_FUSE, A.F_GRAPHS, A.F_DEFER | A.F_GRAPHS, A.F_AUT_GATHER, A.F_AUT_TILED, A.F_GENERIC_MODMUL]
    flags = (A.F_STRICT | (A.F_DEFER if seed % 2 else 0)) if strict else flag_sets[seed % len(flag_sets)]
    pool = rng.choice([0, 34, 34, 40])
    data = np.random.default_rng(seed)
    spm0 = data.integers(0, PRIMES[0], SLOTS * N, dtype=np.uint64)
    if seed % 3 == 0:                                     # raw words: nothing in the batcher may depend on the domain
        spm0[: 4 * N] = data.integers(0, 1 << 63, 4 * N, dtype=np.uint64) * np.uint64(2) + np.uint64(1)
    defined: set = set()
    # call pattern: one by one, or the last two programs as one multi call, or the first as a batch over two sets of rows
    pattern = rng.choice(["single", "multi", "batch"])
    programs = [Gen(rng, defined, strict).program(rng.randrange(2, 7)) for _ in range(2 if pattern == "batch" else 3)]
    # (bases no larger than the first call's, so that the stream's row offsets stay inside the scratchpad)
    other_rows = tuple(rng.randrange(0, programs[0][1][b] // RP + 1) * RP for b in range(3)) + (0, programs[0][1][4])
    machines = [O.GoldenModel(vlmax_bits=N * 64, spm_rows=SLOTS * RP, ksk_rows=KSK_SLOTS * RP, moduli=list(zip(PRIMES, PSIS))),
                A.Engine(vlmax_bits=N * 64, spm_rows=SLOTS * RP, ksk_rows=KSK_SLOTS * RP, moduli=list(zip(PRIMES, PSIS)), flags=flags, pool_buffers=pool)]
    ksk = data.integers(0, PRIMES[0], KSK_SLOTS * N, dtype=np.uint64)
    # between the calls the host overwrites a few rows (registers aliasing them must keep their value) and reads some
    host_rng = random.Random(seed ^ 0x5EED)
    passes = 3 if seed % 2 else 1                         # later passes: cached plans replayed, on other data and entry states
    host_writes = [(host_rng.randrange(SLOTS) * RP, data.integers(0, PRIMES[0], N, dtype=np.uint64)) for _ in range(2 * passes)]
    images = []
    for m in machines:
        m.dma_mem_h2d(0, spm0[: (SLOTS - 3) * N])
        m.dma_ksk_h2d(0, ksk)
        pcs, pc = [], 0
        for prog, _ in programs:
            m.load_isram(prog.words(), pc)
            pcs.append(pc)
            pc += len(prog)
        seen, keep = [], []
        asynchronous = hasattr(m, "dma_mem_h2d_async") and seed % 4 < 2     # the DMA channels beside the VP (engine only)

        def host_write(row, words):
            if asynchronous:
                buf = np.ascontiguousarray(words)
                keep.append(buf)                          # stays valid until the sync
                m.dma_mem_h2d_async(row, buf.ctypes.data, buf.nbytes)
            else:
                m.dma_mem_h2d(row, words)

        def host_read(row):
            if asynchronous:
                buf = np.empty(N, dtype=np.uint64)
                keep.append(buf)
                m.dma_mem_d2h_async(buf.ctypes.data, row, buf.nbytes)
                return buf                                # filled by the time of the final sync
            return m.dma_mem_d2h(row, N).copy()
        for ps in range(passes):
            hw = iter(host_writes[2 * ps: 2 * ps + 2])
            if pattern == "single":
                for pc_, (_, csr) in zip(pcs, programs):
                    m.run_vp(pc_, *csr)
                    if pc_ == pcs[0]:
                        host_write(*next(hw))
            elif pattern == "multi":
                m.run_vp(pcs[0], *programs[0][1])
                host_write(*next(hw))
                m.run_vp_multi([(pcs[1], *programs[1][1]), (pcs[2], *programs[2][1])])
            else:
                m.run_vp_batch(pcs[0], [programs[0][1], other_rows])      # the same stream over two sets of rows
                host_write(*next(hw))
                m.run_vp(pcs[1], *programs[1][1])
            row, data_ = next(hw)
            seen.append(host_read(row))
            host_write(row, data_)
        if asynchronous:
            m.sync()
        images.append((np.concatenate([m.dma_mem_d2h(0, SLOTS * N)] + seen), m.spm_written(0, SLOTS * N)))
    (gd, gw), (ed, ew) = images
    assert (gw == ew).all(), f"seed {seed}: written-mask differs"
    bad = np.nonzero(gd != ed)[0]
    assert bad.size == 0, f"seed {seed} ({pattern}, flags {flags:#x}, pool {pool}, passes {passes}): {bad.size} words differ, first in slot {bad[0] // N}"
    return machines[1].stats()


@pytest.mark.parametrize("block", range(8))
def test_idiom_fuzz_on_the_simulated_device(block):
    per = int(os.environ.get("ALOHA_IDIOM_SEEDS", "40"))
    fused = 0
    with sim_engine.simulated() as A:
        for seed in range(block * per, (block + 1) * per):
            fused += run_case(A, seed)["ops_fused"]
    assert fused > 0, "no stream of this block fused anything: the generator has lost its point"
