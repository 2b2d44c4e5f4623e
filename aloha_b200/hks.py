"""Generator for limb-sharded key-switch / relinearise / rescale instruction streams with dnum digits of
alpha = ceil(L / dnum) limbs and K special primes (hybrid key switching), SURVEY 8(f)1.

The reference ships ONE such kernel, sim/vp/isram_file_generator/keyswitch.mem (122 instructions; SURVEY
App. B.4): L = 2 ciphertext primes, K = 1 special prime, one limb per digit (dnum = L).  Everything below is
that kernel's instruction pattern with the two loops it has unrolled by hand made general:

  * digits of alpha limbs.  A digit is the integer X_b < Q_b = prod_{j in G_b} q_j whose residues are the
    coefficient-form limbs c_j.  Its image under another modulus m is the fast basis extension
        ext_m(X_b) = sum_{j in G_b} [c_j * (Q_b/q_j)^-1 mod q_j] * (Q_b/q_j mod m)   (mod m),
    i.e. per source limb one VFQMUL.vs (done once, by the limb's owner: d_j), then per target modulus a
    VCPY / VFQMOD of d_j into the target modulus (the reference's own base-extension step, insts 8, 12, 28,
    32), a VFQMUL.vs and a VFQADD.  With alpha = 1 both scalars are 1 and the multiplies are not emitted --
    what remains is the reference's sequence word for word.
  * K special primes.  The mod-down by P = prod p_k is the same extension from {p_k} to each q_i, with the
    reference's rounding (add floor(P/2) before, subtract it after: insts 79-82, 85-88).  With K = 1 the
    scalars are 1 and the sequence is the reference's.

Phases (every phase is a set of independent per-limb streams; the machine that owns a limb runs its streams):
  1  limb j < L:      c_j = VAUT(INTT(sw_j))   [no VAUT for relinearise],  d_j = c_j * qhatinv_j -> D[j]
                      rotate only: a'_j = NTT(VAUT(INTT(a_j))) -> A'[j]
     -- all-gather of D --
  2  modulus t:       e_b = NTT_t(ext_t(digit b))  (own group: NTT_t(c_t)),  acc_t,c = sum_b e_b * KSK[t][b][c]
                      special t = L + k: T[k][c] = (INTT(acc_t,c) + floor(P/2)) * phatinv_k
     -- broadcast of T from the owners of the special primes --
  3  limb i < L:      r_c = (acc_i,c - NTT_i(ext_i(T[.][c]) - floor(P/2))) * P^-1 ;  out_c = addend_c + r_c
                      (rotate: addend_0 = a'_i, none for c = 1;  relinearise: addend_c = ct_c,i)
Rescale (drop the last prime) is phase 2's special-prime tail and phase 3 on their own.

Memory: only the three pointer CSRs (+ ksk_ptr) address memory and a VLE / VSE immediate reaches 65536 rows,
so the SPM is laid out in regions of at most 65536 rows and each phase picks three of them.
"""
from __future__ import annotations

from math import prod

import numpy as np

from . import asm


def _ceil_div(a: int, b: int) -> int:
    return -(-a // b)


class Params:
    """Scheme constants of one key-switch shape."""

    def __init__(self, n: int, q: list[int], p: list[int], dnum: int | None = None):
        self.n, self.rp = n, n // 128
        self.q, self.p = list(q), list(p)
        self.L, self.K = len(self.q), len(self.p)
        dnum = self.L if dnum is None else dnum
        assert 1 <= dnum <= self.L and self.K >= 1
        self.alpha = _ceil_div(self.L, dnum)
        self.groups = [list(range(b, min(b + self.alpha, self.L))) for b in range(0, self.L, self.alpha)]
        self.dnum = len(self.groups)
        self.moduli = self.q + self.p
        self.group_of = {j: b for b, g in enumerate(self.groups) for j in g}
        self.P = prod(self.p)
        self.half = self.P // 2
        self.qhat_inv, self.qhat_mod = {}, {}
        for g in self.groups:
            Qb = prod(self.q[j] for j in g)
            for j in g:
                qhat = Qb // self.q[j]
                self.qhat_inv[j] = pow(qhat, -1, self.q[j])
                for t, m in enumerate(self.moduli):
                    if t not in g:
                        self.qhat_mod[j, t] = qhat % m
        self.phat_inv = [pow(self.P // pk, -1, pk) for pk in self.p]
        self.phat_mod = {(k, i): (self.P // pk) % qi for k, pk in enumerate(self.p) for i, qi in enumerate(self.q)}
        self.pinv = [pow(self.P, -1, qi) for qi in self.q]

    def transform_count(self, kind: str = "rotate") -> int:
        """limb-(I)NTTs of one key-switch"""
        L, K = self.L, self.K
        phase1 = 3 * L if kind == "rotate" else L
        return phase1 + self.dnum * (L + K) + 2 * K + 2 * L


class Layout:
    """SPM / KSK row map and limb ownership for one machine of `world`, `batch` key-switches side by side."""

    def __init__(self, prm: Params, world: int = 1, rank: int = 0, batch: int = 1, kind: str = "rotate", base: int = 0):
        """base: first SPM row of the layout (other data may live below it)"""
        assert kind in ("rotate", "relin")
        self.prm, self.world, self.rank, self.batch, self.kind = prm, world, rank, batch, kind
        L, K, rp = prm.L, prm.K, prm.rp
        self.nm = L + K
        self.per_rank = _ceil_div(self.nm, world)
        self.slots = self.per_rank * world
        self.in_polys = (3 if kind == "relin" else 2) * L
        self.IN_size = self.in_polys * rp                       # a | b | (d2)
        self.S_size = (self.slots + L) * rp                     # D[slots] | A'[L]
        self.ACC_size = (2 * self.nm + 2 * K) * rp              # acc[t][c] | T[k][c]
        self.OUT_size = (2 * L + (L if prm.alpha > 1 else 0)) * rp   # out_0 | out_1 | C[L] (unscaled digits)
        for name in ("IN_size", "S_size", "ACC_size", "OUT_size"):
            if getattr(self, name) > 65536:
                raise ValueError(f"region {name[:-5]} exceeds the 16-bit row offset of VLE/VSE; reduce L or N")
        self.IN = base
        self.S = self.IN + batch * self.IN_size
        self.ACC = self.S + batch * self.S_size
        self.OUT = self.ACC + batch * self.ACC_size
        self.spm_rows = self.OUT + batch * self.OUT_size
        self.T_off = 2 * self.nm * rp                            # T inside the ACC region
        self.ksk_slice_rows = 2 * prm.dnum * rp                  # KSK[t] = [b][c] polynomials under modulus t
        self.ksk_rows = self.per_rank * self.ksk_slice_rows

    def owner(self, limb: int) -> int:
        return limb // self.per_rank

    def owned(self, rank: int | None = None) -> list[int]:
        r = self.rank if rank is None else rank
        return [t for t in range(self.nm) if self.owner(t) == r]

    def modulus(self, t: int) -> int:
        return self.prm.moduli[t]

    def ksk_ptr(self, t: int) -> int:
        return (t - self.owner(t) * self.per_rank) * self.ksk_slice_rows

    def region(self, name: str, b: int) -> int:
        return getattr(self, name) + b * getattr(self, name + "_size")


def _extend(p: asm.Program, src_mod: int, dst_mod: int, vd: int, vs: int):
    """keyswitch.mem: VFQMOD when the source modulus is larger than the target (inst 28), VCPY when the
    target is larger (insts 8, 12, 32)"""
    (p.vfqmod if src_mod > dst_mod else p.vcpy)(vd, vs)


def phase1_stream(lay: Layout, j: int) -> asm.Program:
    """src0 = IN, src1 = S, rslt = OUT.  Registers as keyswitch.mem insts 3-5, 17-20."""
    prm, rp, L = lay.prm, lay.prm.rp, lay.prm.L
    p = asm.Program().vsetvl(prm.n).vsetq(prm.q[j])
    if lay.kind == "rotate":
        p.vle(4, asm.BASE_SRC0, (L + j) * rp).vintt(2, 4).vaut(4, 2)
        c = 4
    else:
        p.vle(4, asm.BASE_SRC0, (2 * L + j) * rp).vintt(2, 4)
        c = 2
    if prm.alpha == 1:
        p.vse(c, asm.BASE_SRC1, j * rp)
    else:
        p.vse(c, asm.BASE_RSLT, (2 * L + j) * rp)
        p.vfqmul(8, c, imm=prm.qhat_inv[j]).vse(8, asm.BASE_SRC1, j * rp)
    if lay.kind == "rotate":
        p.vle(3, asm.BASE_SRC0, j * rp).vintt(6, 3).vaut(3, 6).vntt(2, 3).vse(2, asm.BASE_SRC1, (lay.slots + j) * rp)
    return p.brk()


def phase2_stream(lay: Layout, t: int, groups: list[int] | None = None, first: bool = True, last: bool = True) -> asm.Program:
    """src0 = S (gathered digits), src1 = OUT (unscaled own digits), rslt = ACC, base 15 = KSK[t].
    `groups` restricts the stream to some digits (a machine can start on the digits it already holds while the
    others are in flight); `first` / `last` say whether the accumulators start here / are finished here."""
    prm, rp, L, K = lay.prm, lay.prm.rp, lay.prm.L, lay.prm.K
    mt = prm.moduli[t]
    groups = list(range(prm.dnum)) if groups is None else groups
    p = asm.Program().vsetvl(prm.n).vsetq(mt)
    acc_row = lambda c: (2 * t + c) * rp
    # A later chunk sums its own digits first and adds the accumulators of the earlier chunks at the end: a chain that
    # starts from a product is one sum-of-products for the batcher, one that starts from a loaded accumulator is a
    # serial multiply-add per digit.  (Every term is canonical, so the order of the modular additions does not show.)
    started = False
    for b in groups:
        g = prm.groups[b]
        if t in g:
            if prm.alpha == 1:
                p.vle(0, asm.BASE_SRC0, t * rp)
            else:
                p.vle(0, asm.BASE_SRC1, (2 * L + t) * rp)
            src = 0
        elif prm.alpha == 1:
            j = g[0]
            p.vle(0, asm.BASE_SRC0, j * rp)
            _extend(p, prm.q[j], mt, 8, 0)
            src = 8
        else:
            for n_, j in enumerate(g):
                p.vle(0, asm.BASE_SRC0, j * rp)
                _extend(p, prm.q[j], mt, 8, 0)
                if n_ == 0:
                    p.vfqmul(13, 8, imm=prm.qhat_mod[j, t])
                else:
                    p.vfqmul(10, 8, imm=prm.qhat_mod[j, t]).vfqadd(13, 13, 10)
            src = 13
        p.vntt(2, src)
        for c in (0, 1):
            k, prod_, acc = 1 + 2 * c, 5 + 2 * c, 4 + 2 * c          # odd key reg, odd product, even accumulator
            p.vle(k, asm.BASE_KSK, (2 * b + c) * rp)
            if not started:
                p.vfqmul(acc, 2, k)
            else:
                p.vfqmul(prod_, 2, k).vfqadd(acc, acc, prod_)
        started = True
    acc0, acc1 = 4, 6
    if not first:
        # (the sum lands in the register that held the earlier accumulator: were that register left aliasing the
        #  accumulator's rows, the store below would have to copy it out of the way first)
        p.vle(15, asm.BASE_RSLT, acc_row(0)).vfqadd(15, 15, 4).vle(17, asm.BASE_RSLT, acc_row(1)).vfqadd(17, 17, 6)
        acc0, acc1 = 15, 17
    if not last or t < L:
        p.vse(acc0, asm.BASE_RSLT, acc_row(0)).vse(acc1, asm.BASE_RSLT, acc_row(1))
    else:
        k = t - L
        half = prm.half % mt                                          # keyswitch.mem insts 79-82
        for c, (acc, tmp, out) in enumerate(((acc0, 8, 10), (acc1, 9, 11))):
            p.vintt(tmp, acc).vfqadd(out, tmp, imm=half)
            if K > 1:
                p.vfqmul(12 + c, out, imm=prm.phat_inv[k])
                out = 12 + c
            p.vse(out, asm.BASE_RSLT, lay.T_off + (2 * k + c) * rp)
    return p.brk()


def phase3_stream(lay: Layout, i: int) -> asm.Program:
    """src0 = ACC (+T), src1 = S (rotate: A') or IN (relinearise: the addends), rslt = OUT.  keyswitch.mem insts 85-120."""
    prm, rp, L, K = lay.prm, lay.prm.rp, lay.prm.L, lay.prm.K
    qi = prm.q[i]
    p = asm.Program().vsetvl(prm.n).vsetq(qi)
    for c in (0, 1):
        if K == 1:
            p.vle(0, asm.BASE_SRC0, lay.T_off + c * rp).vfqsub(2, 0, imm=prm.half % qi)
        else:
            for k in range(K):
                p.vle(0, asm.BASE_SRC0, lay.T_off + (2 * k + c) * rp)
                _extend(p, prm.p[k], qi, 8, 0)
                if k == 0:
                    p.vfqmul(13, 8, imm=prm.phat_mod[k, i])
                else:
                    p.vfqmul(10, 8, imm=prm.phat_mod[k, i]).vfqadd(13, 13, 10)
            p.vfqsub(2, 13, imm=prm.half % qi)
        p.vntt(4, 2)
        p.vle(1, asm.BASE_SRC0, (2 * i + c) * rp).vfqsub(6, 1, 4).vfqmul(8, 6, imm=prm.pinv[i])
        if lay.kind == "rotate":
            if c == 0:
                p.vle(3, asm.BASE_SRC1, (lay.slots + i) * rp).vfqadd(10, 3, 8).vse(10, asm.BASE_RSLT, i * rp)
            else:
                p.vse(8, asm.BASE_RSLT, (L + i) * rp)
        else:
            p.vle(3, asm.BASE_SRC1, (c * L + i) * rp).vfqadd(10, 3, 8).vse(10, asm.BASE_RSLT, (c * L + i) * rp)
    return p.brk()


SCRATCH_REGS = tuple(range(0, 14)) + (15, 17)      # every register the phase streams write


def retire_stream(prm: Params) -> asm.Program:
    """Last call of every batch of phase streams: reload the scratch registers from key memory.  A VLE is an alias for
    the batcher (no kernel), and it ends the lifetime of whatever the LAST phase stream of the batch left in those
    registers -- architecturally visible values the batcher would otherwise have to materialise, which keeps that
    stream's multiply / add / base-extension chains from fusing and trails a dozen one-job launches behind every
    batch.  (Key memory, because the VP never writes it: aliases of scratchpad rows would have to be copied out of
    the way by the next store to those rows.)"""
    p = asm.Program().vsetvl(prm.n)
    for r in SCRATCH_REGS:
        p.vle(r, asm.BASE_KSK, 0)
    return p.brk()


# ---------------------------------------------------------------------------------------------- exchange
class LocalComm:
    """world = 1: no exchange."""
    world, rank = 1, 0

    def all_gather(self, machine, row, rows_per_rank, count, stride, chunked=False):
        pass

    def broadcast(self, machine, row, nrows, root, count, stride):
        pass

    def wait(self, machine, source=-1):
        pass


class GroupComm:
    """The product path: NCCL inside the C library (aloha_group_*), transfers beside the kernels."""

    def __init__(self, group):
        self.group, self.world, self.rank = group, group.size, group.rank

    def all_gather(self, machine, row, rows_per_rank, count, stride, chunked=False):
        self.group.all_gather_rows(row, rows_per_rank, count, stride, chunked)

    def broadcast(self, machine, row, nrows, root, count, stride):
        for c in range(count):
            self.group.broadcast_rows(row + c * stride, nrows, root)

    def wait(self, machine, source=-1):
        self.group.wait(source)


class TorchComm:
    """torch.distributed exchange staged through host arrays: for machines without device memory (the CPU
    oracle under gloo in the tests)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)

    def all_gather(self, machine, row, rows_per_rank, count, stride, chunked=False):
        import torch
        for c in range(count):
            base = row + c * stride
            mine = torch.from_numpy(machine.dma_mem_d2h(base + self.rank * rows_per_rank, rows_per_rank * 128).view(np.int64))
            parts = [torch.empty_like(mine) for _ in range(self.world)]
            self.dist.all_gather(parts, mine, group=self.group)
            machine.dma_mem_h2d(base, np.concatenate([x.numpy() for x in parts]).view(np.uint64))

    def broadcast(self, machine, row, nrows, root, count, stride):
        import torch
        for c in range(count):
            t = torch.from_numpy(machine.dma_mem_d2h(row + c * stride, nrows * 128).view(np.int64).copy())
            self.dist.broadcast(t, src=root, group=self.group)
            machine.dma_mem_h2d(row + c * stride, t.numpy().view(np.uint64))

    def wait(self, machine, source=-1):
        pass


class KeySwitch:
    """Runs the three phases on one machine (anything with the Engine method set) of a group.
    overlap: False -- phase 2 waits for the whole all-gather;
             "own" -- phase 2 is cut in two: the digits this machine produced itself (no wait: they run under the
                      all-gather), then all the others;
             "chunks" (or True) -- phase 2 is cut by the rank that owns each digit's limbs; the all-gather moves one
                      source rank at a time and the machine consumes the blocks as they arrive (its own first)."""

    def __init__(self, machine, lay: Layout, comm=None, pc_base: int = 0, overlap: bool = False, lockstep: bool = False):
        """lockstep: every rank consumes the chunks in the same order and waits for each of them, its own
        included -- what a single host thread driving all machines of a local group needs (run_local_group)."""
        self.machine, self.lay, self.comm, self.lockstep = machine, lay, comm or LocalComm(), lockstep
        assert self.comm.world == lay.world and self.comm.rank == lay.rank
        prm = lay.prm
        self.overlap = ("chunks" if overlap is True else overlap) if lay.world > 1 else False
        # a digit can be consumed once the blocks of all the ranks that hold some of its limbs have arrived: the chunked
        # all-gather delivers the blocks in rank order, so a digit belongs to the chunk of the LAST such rank; it needs no
        # wait at all only if every one of its limbs is this rank's own
        self.digit_sources = [sorted({lay.owner(j) for j in g}) for g in prm.groups]
        self.digit_rank = [src[-1] for src in self.digit_sources]
        self.digit_local = [src == [lay.rank] for src in self.digit_sources]
        self.pc1, self.pc2, self.pc3 = {}, {}, {}
        pc = pc_base

        def put(table, key, prog):
            nonlocal pc
            words = prog.words()
            machine.load_isram(words, pc)
            table[key] = pc
            pc += len(words)
        self.pc_tail = {}
        put(self.pc_tail, 0, retire_stream(prm))
        for t in lay.owned():
            if t < prm.L:
                put(self.pc1, t, phase1_stream(lay, t))
                put(self.pc3, t, phase3_stream(lay, t))
            if not self.overlap:
                put(self.pc2, (t, None), phase2_stream(lay, t))
            else:
                order = [r for r in self.chunk_order() if self.chunk_digits(r)]      # (lockstep keeps empty chunks in the list)
                for n_, r in enumerate(order):
                    put(self.pc2, (t, r), phase2_stream(lay, t, self.chunk_digits(r), first=n_ == 0, last=n_ == len(order) - 1))
        self.pc_end = pc

    def chunk_order(self) -> list[int]:
        """Chunks of phase 2 in the order they run.  "chunks": one per source rank (the rank of a digit's last limb),
        this rank's first, then the order of arrival (lockstep: rank order for everyone).  "own": the digits made of
        this rank's limbs only (chunk = its rank), then -1 = all the others."""
        have = sorted({r for r in self.digit_rank})
        if self.overlap == "own":
            if self.lockstep:                   # every rank walks the same list, whether or not it holds a whole digit
                return [self.lay.rank, -1]
            mine = [self.lay.rank] if any(self.digit_local) else []
            return mine + ([-1] if not all(self.digit_local) else [])
        if self.lockstep:
            return have
        return [r for r in have if r == self.lay.rank] + [r for r in have if r != self.lay.rank]

    def chunk_digits(self, r: int) -> list[int]:
        dnum = self.lay.prm.dnum
        if self.overlap == "own":
            return [b for b in range(dnum) if self.digit_local[b] == (r != -1)]
        return [b for b in range(dnum) if self.digit_rank[b] == r]

    def chunk_waits(self, r: int) -> list[int]:
        """the transfers chunk r's digits must have seen arrive: ALOHA_GROUP_ALL for the unchunked all-gather, else the
        source ranks of its digits' limbs (lockstep: this rank's own block included -- one list for every rank)"""
        if self.overlap == "own":
            return [-1] if r == -1 else []
        need = sorted({s for b in self.chunk_digits(r) for s in self.digit_sources[b]})
        return need if self.lockstep else [s for s in need if s != self.lay.rank]

    def load_ksk(self, t: int, data: np.ndarray):
        """data: KSK[t] = 2 dnum polynomials ([b][c] order) under modulus t; only the owner stores it."""
        lay = self.lay
        assert lay.owner(t) == lay.rank and data.size == 2 * lay.prm.dnum * lay.prm.n
        self.machine.dma_ksk_h2d(lay.ksk_ptr(t), data.reshape(-1))

    def load_input(self, i: int, polys, b: int = 0):
        """polys: (a_i, b_i) for rotate, (d0_i, d1_i, d2_i) for relinearise -- limb i of batch element b"""
        lay, rp, L = self.lay, self.lay.prm.rp, self.lay.prm.L
        for c, x in enumerate(polys):
            self.machine.dma_mem_h2d(lay.region("IN", b) + (c * L + i) * rp, x)

    def program(self, galois_k: int = 1, only: list[int] | None = None) -> list[tuple]:
        """The key-switch as a list of machine-level ops, in issue order:
             ("run", [(pc, src0, src1, rslt, ksk_ptr, step), ...])        one aloha_run_vp_multi
             ("all_gather", row, rows_per_rank, count, stride, chunked)   aloha_group_all_gather_rows
             ("broadcast", row, nrows, root, count, stride)               aloha_group_broadcast_rows (x count)
             ("wait", source)                                             aloha_group_wait
        Every rank's list has the same collectives in the same order.
        only: restrict phases 2 and 3 to these output limbs (+ the special primes) -- for checks that want a
        few output limbs without paying for all of them."""
        lay, prm = self.lay, self.lay.prm
        B, rp = lay.batch, prm.rp
        mine = lay.owned()
        want = lambda t: only is None or t >= prm.L or t in only
        reg = lay.region
        ops = [("run", [(self.pc1[j], reg("IN", b), reg("S", b), reg("OUT", b), 0, galois_k)
                        for b in range(B) for j in mine if j < prm.L])]
        ops.append(("all_gather", lay.S, lay.per_rank * rp, B, lay.S_size, self.overlap == "chunks"))
        if not self.overlap:
            ops.append(("wait", -1))
            ops.append(("run", [(self.pc2[t, None], reg("S", b), reg("OUT", b), reg("ACC", b), lay.ksk_ptr(t), 0)
                                for b in range(B) for t in mine if want(t)]))
        else:
            for r in self.chunk_order():
                for source in self.chunk_waits(r):
                    ops.append(("wait", source))
                ops.append(("run", [(self.pc2[t, r], reg("S", b), reg("OUT", b), reg("ACC", b), lay.ksk_ptr(t), 0)
                                    for b in range(B) for t in mine if want(t)] if self.chunk_digits(r) else []))
        for r in range(lay.world):
            ks = [t - prm.L for t in lay.owned(r) if t >= prm.L]
            if ks:
                ops.append(("broadcast", lay.ACC + lay.T_off + 2 * ks[0] * rp, 2 * len(ks) * rp, r, B, lay.ACC_size))
        ops.append(("wait", -2))
        src1 = "S" if lay.kind == "rotate" else "IN"
        ops.append(("run", [(self.pc3[i], reg("ACC", b), reg(src1, b), reg("OUT", b), 0, 0)
                            for b in range(B) for i in mine if i < prm.L and want(i)]))
        tail = (self.pc_tail[0], 0, 0, 0, 0, 0)
        return [("run", op[1] + [tail]) if op[0] == "run" and op[1] else op for op in ops]

    def execute(self, ops):
        m, comm = self.machine, self.comm
        for op in ops:
            if op[0] == "run":
                if op[1]:
                    m.run_vp_multi(op[1])
            elif op[0] == "all_gather":
                comm.all_gather(m, *op[1:])
            elif op[0] == "broadcast":
                comm.broadcast(m, *op[1:])
            else:
                comm.wait(m, op[1])

    def run(self, galois_k: int = 1, only: list[int] | None = None):
        self.execute(self.program(galois_k, only))

    def read_output(self, i: int, b: int = 0):
        lay, rp, L = self.lay, self.lay.prm.rp, self.lay.prm.L
        out = lay.region("OUT", b)
        return (self.machine.dma_mem_d2h(out + i * rp, lay.prm.n), self.machine.dma_mem_d2h(out + (L + i) * rp, lay.prm.n))


def run_local_group(switches: list["KeySwitch"], group, galois_k: int = 1, only=None):
    """One process driving every machine of a local group (aloha_group_create_local): the ranks' op lists are
    walked in lockstep -- run ops go to each machine, every collective is issued once for the whole group."""
    progs = [ks.program(galois_k, only) for ks in switches]
    assert len({len(p) for p in progs}) == 1
    for step in zip(*progs):
        kind = step[0][0]
        assert all(op[0] == kind for op in step)
        if kind == "run":
            for ks, op in zip(switches, step):
                if op[1]:
                    ks.machine.run_vp_multi(op[1])
        elif kind == "all_gather":
            group.all_gather_rows(*step[0][1:5], chunked=step[0][5])
        elif kind == "broadcast":
            _, row, nrows, root, count, stride = step[0]
            for c in range(count):
                group.broadcast_rows(row + c * stride, nrows, root)
        else:
            group.wait(step[0][1])                  # overlap: the switches were built with lockstep=True


def write_replay_case(directory: str, switches: list["KeySwitch"], engine_cfg: dict, loads, dumps, galois_k: int = 1,
                      only=None):
    """Everything the C replay tool (aloha_b200/csrc/group_replay_main.cpp -- a caller that knows only
    include/aloha_b200.h) needs to run this key-switch without Python: per rank r
        rank<r>.cfg    vlmax_bits spm_rows ksk_rows pool_buffers isram_depth nmoduli, then q psi per modulus
        rank<r>.isram  the instruction ROM image, n x 12 bytes
        rank<r>.prog   one op per line: run <ncalls> then ncalls lines "pc src0 src1 rslt ksk step";
                       allgather row rows_per_rank count stride chunked; bcast row nrows root; wait source
        rank<r>.io     "spm|ksk <row> <file>" loads and "dump <row> <nwords> <file>" read-backs (raw u64 files)
    loads[r] = [(space, row, array)], dumps[r] = [(row, nwords, name)]."""
    import os
    os.makedirs(directory, exist_ok=True)
    for r, ks in enumerate(switches):
        m = ks.machine                     # a recorder: anything that kept the load_isram calls
        with open(os.path.join(directory, f"rank{r}.cfg"), "w") as f:
            mods = engine_cfg["moduli"]
            f.write(f"{engine_cfg['vlmax_bits']} {ks.lay.spm_rows} {max(ks.lay.ksk_rows, 1)} {engine_cfg.get('pool_buffers', 0)} "
                    f"{engine_cfg.get('isram_depth', 0)} {len(mods)}\n")
            for q, psi in mods:
                f.write(f"{q} {psi}\n")
        rom = np.zeros((ks.pc_end, 12), dtype=np.uint8)
        for words, pc in m.isram_loads:
            rom[pc:pc + len(words)] = words
        rom.tofile(os.path.join(directory, f"rank{r}.isram"))
        with open(os.path.join(directory, f"rank{r}.prog"), "w") as f:
            for op in ks.program(galois_k, only):
                if op[0] == "run":
                    f.write(f"run {len(op[1])}\n")
                    for c in op[1]:
                        f.write(" ".join(str(v) for v in c) + "\n")
                elif op[0] == "all_gather":
                    f.write(f"allgather {op[1]} {op[2]} {op[3]} {op[4]} {int(op[5])}\n")
                elif op[0] == "broadcast":
                    for c in range(op[4]):
                        f.write(f"bcast {op[1] + c * op[5]} {op[2]} {op[3]}\n")
                else:
                    f.write(f"wait {op[1]}\n")
        with open(os.path.join(directory, f"rank{r}.io"), "w") as f:
            for n_, (space, row, arr) in enumerate(loads[r]):
                name = f"rank{r}.in{n_}.u64"
                np.ascontiguousarray(arr, dtype=np.uint64).tofile(os.path.join(directory, name))
                f.write(f"{space} {row} {name}\n")
            for row, nwords, name in dumps[r]:
                f.write(f"dump {row} {nwords} {name}\n")


class Recorder:
    """A machine that only records what a KeySwitch loads into it (for write_replay_case)."""

    def __init__(self):
        self.isram_loads, self.loads = [], []

    def load_isram(self, words, pc):
        self.isram_loads.append((np.array(words), pc))

    def dma_mem_h2d(self, row, data):
        self.loads.append(("spm", row, np.array(data, dtype=np.uint64).reshape(-1)))

    def dma_ksk_h2d(self, row, data):
        self.loads.append(("ksk", row, np.array(data, dtype=np.uint64).reshape(-1)))


# ------------------------------------------------------------------------------------------------ rescale
class Rescale:
    """Drop the last ciphertext prime: per remaining limb i and component c
        out_c,i = (ct_c,i - NTT_i(ext_i(INTT(ct_c,last)) + round)) * q_last^-1  (mod q_i),
    the same instruction pattern as the key-switch's mod-down (keyswitch.mem insts 79-120) with P = q_last.
    Regions: IN = ct (2 L polys) , T (2 polys), OUT (2 (L-1) polys)."""

    def __init__(self, machine, n: int, q: list[int], world: int = 1, rank: int = 0, comm=None, pc_base: int = 0,
                 in_row: int | None = None, base: int = 0, per_rank: int | None = None, tail: tuple | None = None):
        """in_row: the ciphertext is already in the SPM at this row (component c, limb i at (c L + i) polys) --
        e.g. a key-switch's OUT region; T and OUT then start at `base`.  per_rank: limbs per machine when the
        ownership is another stage's (Layout.per_rank) rather than ceil(L / world)."""
        self.machine, self.n, self.q, self.rp = machine, n, list(q), n // 128
        self.comm = comm or LocalComm()
        self.world, self.rank = world, rank
        self.tail = [tail] if tail else []       # a retire_stream call to close each batch with (needs key memory)
        L, rp = len(self.q), self.rp
        self.L = L
        self.per_rank = per_rank or _ceil_div(L, world)
        if in_row is None:
            self.IN, self.T = base, base + 2 * L * rp
        else:
            self.IN, self.T = in_row, base
        self.OUT = self.T + 2 * rp
        self.spm_rows = self.OUT + 2 * (L - 1) * rp
        if 2 * L * rp > 65536:
            raise ValueError("region exceeds the 16-bit row offset of VLE/VSE")
        ql = self.q[-1]
        half = ql // 2
        self.pc_t, self.pc_out = None, {}
        pc = pc_base
        if self.owner(L - 1) == rank:
            p = asm.Program().vsetvl(n).vsetq(ql)
            for c in (0, 1):
                p.vle(4, asm.BASE_SRC0, (c * L + L - 1) * rp).vintt(8, 4).vfqadd(10, 8, imm=half).vse(10, asm.BASE_RSLT, c * rp)
            words = p.brk().words()
            machine.load_isram(words, pc)
            self.pc_t = pc
            pc += len(words)
        for i in range(L - 1):
            if self.owner(i) != rank:
                continue
            qi = self.q[i]
            p = asm.Program().vsetvl(n).vsetq(qi)
            for c in (0, 1):
                p.vle(0, asm.BASE_SRC1, c * rp).vfqsub(2, 0, imm=half % qi).vntt(4, 2)
                p.vle(1, asm.BASE_SRC0, (c * L + i) * rp).vfqsub(6, 1, 4).vfqmul(8, 6, imm=pow(ql, -1, qi))
                p.vse(8, asm.BASE_RSLT, (c * (L - 1) + i) * rp)
            words = p.brk().words()
            machine.load_isram(words, pc)
            self.pc_out[i] = pc
            pc += len(words)
        self.pc_end = pc

    def owner(self, i: int) -> int:
        return i // self.per_rank

    def load_input(self, i: int, c0: np.ndarray, c1: np.ndarray):
        self.machine.dma_mem_h2d(self.IN + i * self.rp, c0)
        self.machine.dma_mem_h2d(self.IN + (self.L + i) * self.rp, c1)

    def run(self):
        m = self.machine
        if self.pc_t is not None:
            m.run_vp_multi([(self.pc_t, self.IN, 0, self.T, 0, 0)] + self.tail)
        self.comm.broadcast(m, self.T, 2 * self.rp, self.owner(self.L - 1), 1, 0)
        self.comm.wait(m, -2)
        if self.pc_out:
            m.run_vp_multi([(pc, self.IN, self.T, self.OUT, 0, 0) for pc in self.pc_out.values()] + self.tail)

    def read_output(self, i: int):
        return (self.machine.dma_mem_d2h(self.OUT + i * self.rp, self.n),
                self.machine.dma_mem_d2h(self.OUT + (self.L - 1 + i) * self.rp, self.n))


# ----------------------------------------------------------------------------------------------- multiply
def tensor_stream(prm: Params, i: int) -> asm.Program:
    """Limb i of the degree-2 product of two ciphertexts (evaluation form): src0 = X (a0 | a1 | b0 | b1, L limbs
    each), rslt = the relinearise layout's IN (d0 | d1 | d2):
        d0 = a0 b0,  d1 = a0 b1 + a1 b0,  d2 = a1 b1.
    mul_plain.mem's VLE / VFQMUL.vv / VSE pattern with hom_add.mem's VFQADD.vv (B.3); the
    operands of every vv instruction sit in different register banks."""
    rp, L = prm.rp, prm.L
    p = asm.Program().vsetvl(prm.n).vsetq(prm.q[i])
    p.vle(0, asm.BASE_SRC0, i * rp).vle(2, asm.BASE_SRC0, (L + i) * rp)                 # a0, a1: even bank
    p.vle(1, asm.BASE_SRC0, (2 * L + i) * rp).vle(3, asm.BASE_SRC0, (3 * L + i) * rp)   # b0, b1: odd bank
    p.vfqmul(4, 0, 1).vse(4, asm.BASE_RSLT, i * rp)
    p.vfqmul(5, 0, 3).vfqmul(8, 2, 1).vfqadd(10, 5, 8).vse(10, asm.BASE_RSLT, (L + i) * rp)
    p.vfqmul(6, 2, 3).vse(6, asm.BASE_RSLT, (2 * L + i) * rp)
    return p.brk()


class Multiply:
    """Ciphertext x ciphertext on one machine of a group: tensor product, relinearise (hybrid key-switch of d2
    under the relinearisation key), rescale by the last prime -- three stages of per-limb streams over ONE
    scratchpad image, the output of each stage read by the next where it lies:
        X [4 L polys] | relinearise layout (IN = d0 d1 d2, S, ACC, OUT = c0 c1) | rescale T, OUT.
    Limb ownership is the key-switch's (Layout.owner) for all three stages, so the only exchanges are the
    key-switch's all-gather and broadcast and the rescale's broadcast of the dropped limb."""

    def __init__(self, machine, prm: Params, world: int = 1, rank: int = 0, comm=None, pc_base: int = 0,
                 overlap=False, rescale: bool = True):
        L, rp = prm.L, prm.rp
        self.machine, self.prm = machine, prm
        self.X = 0
        self.lay = Layout(prm, world, rank, 1, "relin", base=4 * L * rp)
        self.ks = KeySwitch(machine, self.lay, comm, pc_base, overlap)
        self.tail = (self.ks.pc_tail[0], 0, 0, 0, 0, 0)
        pc = self.ks.pc_end
        self.pc_tensor = {}
        for i in self.lay.owned():
            if i < L:
                words = tensor_stream(prm, i).words()
                machine.load_isram(words, pc)
                self.pc_tensor[i] = pc
                pc += len(words)
        self.rs = None
        if rescale:
            self.rs = Rescale(machine, prm.n, prm.q, world, rank, self.ks.comm, pc, in_row=self.lay.OUT,
                              base=self.lay.spm_rows, per_rank=self.lay.per_rank, tail=self.tail)
            pc = self.rs.pc_end
        self.pc_end = pc

    @staticmethod
    def spm_rows(prm: Params, world: int = 1, rescale: bool = True) -> int:
        end = Layout(prm, world, 0, 1, "relin", base=4 * prm.L * prm.rp).spm_rows
        return end + (2 * prm.L * prm.rp if rescale else 0)

    def load_input(self, i: int, a: tuple, b: tuple):
        """limb i of the two ciphertexts: a = (a0_i, a1_i), b = (b0_i, b1_i)"""
        L, rp = self.prm.L, self.prm.rp
        for c, x in enumerate((*a, *b)):
            self.machine.dma_mem_h2d(self.X + (c * L + i) * rp, x)

    def load_ksk(self, t: int, data: np.ndarray):
        self.ks.load_ksk(t, data)

    def run(self):
        lay = self.lay
        if self.pc_tensor:
            self.machine.run_vp_multi([(pc, self.X, 0, lay.IN, 0, 0) for pc in self.pc_tensor.values()] + [self.tail])
        self.ks.run(1)
        if self.rs is not None:
            self.rs.run()

    def read_output(self, i: int):
        """limb i (rescale: i < L - 1) of the product ciphertext"""
        return self.rs.read_output(i) if self.rs is not None else self.ks.read_output(i)
