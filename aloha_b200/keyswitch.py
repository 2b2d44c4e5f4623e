"""Limb-sharded key-switch (rotate) stream: the reference's `keyswitch` microcode generalised from
L = 2 ciphertext primes to any L (+ one special prime P), partitioned by RNS limb over G machines.

Reference: sim/vp/isram_file_generator/keyswitch.mem (122 instructions) -- digit = one
limb, K = 1 special prime.  For L = 2 and G = 1 the streams below perform, per coefficient, exactly
the arithmetic of that kernel (same opcodes, same operand order), so they reproduce the reference's
`ct_after_rotate` vectors bit for bit; tests pin that.

Phases (each is a set of independent per-limb instruction streams):
  1  limb j < L, on its owner:   d_j  = VAUT(INTT_qj(b_j))            -> D[j]      (coefficient form)
                                 a'_j = NTT_qj(VAUT(INTT_qj(a_j)))    -> A'[j]
     --- all-gather of D over the machines (the only bulk collective) ---
  2  modulus i <= L, on its owner: e_ji = NTT_qi(ext_qi(d_j)) for every j   (VCPY / VFQMOD base extension)
                                 acc_i,c = sum_j e_ji * KSK[i][j][c]  -> ACC[i][c]
     for i = L (the special prime): t_c = INTT_P(acc_L,c) + (P-1)/2    -> T[c]
     --- broadcast of T (2 polynomials) from the owner of P ---
  3  limb i < L: r_c = (acc_i,c - NTT_qi(t_c - (P-1)/2)) * P^-1 mod q_i ;  out_a[i] = a'_i + r_0, out_b[i] = r_1

Only the three pointer CSRs + ksk_ptr address memory, and a VLE/VSE immediate reaches 65536 rows, so
the SPM is laid out in regions of at most 65536 rows and each phase picks three of them.
"""
from __future__ import annotations

import numpy as np

from . import asm


def _ceil_div(a: int, b: int) -> int:
    return -(-a // b)


class KeySwitchLayout:
    """SPM / KSK row map and limb ownership for one machine of `world`."""

    def __init__(self, n: int, ct_moduli: list[int], special: int, world: int = 1, rank: int = 0):
        self.n, self.rp = n, n // 128
        self.q = list(ct_moduli)
        self.P = special
        self.L = len(self.q)
        self.world, self.rank = world, rank
        self.per_rank = _ceil_div(self.L + 1, world)          # limbs (incl. P = index L) per machine
        self.slots = self.per_rank * world                      # padded limb count (uniform blocks)
        L, rp = self.L, self.rp
        # regions (rows)
        self.IN = 0                                              # a_0..a_{L-1}, b_0..b_{L-1}
        self.S = self.IN + 2 * L * rp                            # D[0..slots) then A'[0..L)
        self.ACC = self.S + (self.slots + L) * rp                # ACC[i][c] at (2i + c); T[c] = slot 2L + c
        self.OUT = self.ACC + (2 * L + 2) * rp                   # out_a[0..L), out_b[0..L)
        self.spm_rows = self.OUT + 2 * L * rp
        for region_rows in (2 * L * rp, (self.slots + L) * rp, (2 * L + 2) * rp):
            if region_rows > 65536:
                raise ValueError("a region exceeds the 16-bit row offset of VLE/VSE; reduce L or N")
        self.ksk_slice_rows = 2 * L * rp                         # KSK[i] = [j][c] polys of one modulus
        self.ksk_rows = self.per_rank * self.ksk_slice_rows

    def owner(self, limb: int) -> int:
        return limb // self.per_rank

    def owned(self, rank: int | None = None) -> list[int]:
        r = self.rank if rank is None else rank
        return [i for i in range(self.L + 1) if self.owner(i) == r]

    def modulus(self, i: int) -> int:
        return self.q[i] if i < self.L else self.P

    def ksk_ptr(self, i: int) -> int:
        """local KSK row of modulus i's slice on its owner"""
        return (i - self.owner(i) * self.per_rank) * self.ksk_slice_rows


def phase1_stream(lay: KeySwitchLayout, j: int) -> asm.Program:
    """src0 = IN, src1 = S (stores!), rslt unused.  Register choice follows keyswitch.mem insts 3-5, 17-20."""
    L, rp = lay.L, lay.rp
    p = asm.Program().vsetvl(lay.n).vsetq(lay.q[j])
    p.vle(4, asm.BASE_SRC0, (L + j) * rp).vintt(2, 4).vaut(4, 2).vse(4, asm.BASE_SRC1, j * rp)
    p.vle(3, asm.BASE_SRC0, j * rp).vintt(6, 3).vaut(3, 6).vntt(2, 3).vse(2, asm.BASE_SRC1, (lay.slots + j) * rp)
    return p.brk()


def phase2_stream(lay: KeySwitchLayout, i: int) -> asm.Program:
    """src0 = S (digits), rslt = ACC, base 15 = this modulus' KSK slice."""
    L, rp = lay.L, lay.rp
    qi = lay.modulus(i)
    p = asm.Program().vsetvl(lay.n).vsetq(qi)
    for j in range(L):
        p.vle(0, asm.BASE_SRC0, j * rp)
        if j == i:
            src = 0
        else:
            # keyswitch.mem: VFQMOD when the source modulus is larger than the target (insts 28),
            # VCPY when the target is larger (insts 8, 12, 32)
            (p.vfqmod if lay.q[j] > qi else p.vcpy)(8, 0)
            src = 8
        p.vntt(2, src)
        for c in (0, 1):
            k, prod, acc = 1 + 2 * c, 5 + 2 * c, 4 + 2 * c          # odd key reg, odd product, even accumulator
            p.vle(k, asm.BASE_KSK, (2 * j + c) * rp)
            if j == 0:
                p.vfqmul(acc, 2, k)
            else:
                p.vfqmul(prod, 2, k).vfqadd(acc, acc, prod)
    if i < L:
        p.vse(4, asm.BASE_RSLT, (2 * i) * rp).vse(6, asm.BASE_RSLT, (2 * i + 1) * rp)
    else:
        half = (lay.P - 1) // 2                                       # keyswitch.mem insts 79-82
        p.vintt(8, 4).vfqadd(10, 8, imm=half).vse(10, asm.BASE_RSLT, (2 * L) * rp)
        p.vintt(9, 6).vfqadd(11, 9, imm=half).vse(11, asm.BASE_RSLT, (2 * L + 1) * rp)
    return p.brk()


def phase3_stream(lay: KeySwitchLayout, i: int) -> asm.Program:
    """src0 = ACC (+T), src1 = S (A'), rslt = OUT.  keyswitch.mem insts 85-120."""
    L, rp = lay.L, lay.rp
    qi = lay.q[i]
    half = (lay.P - 1) // 2
    pinv = pow(lay.P, -1, qi)
    p = asm.Program().vsetvl(lay.n).vsetq(qi)
    for c in (0, 1):
        p.vle(0, asm.BASE_SRC0, (2 * L + c) * rp).vfqsub(2, 0, imm=half).vntt(4, 2)
        p.vle(1, asm.BASE_SRC0, (2 * i + c) * rp).vfqsub(6, 1, 4).vfqmul(8, 6, imm=pinv)
        if c == 0:
            p.vle(3, asm.BASE_SRC1, (lay.slots + i) * rp).vfqadd(10, 3, 8).vse(10, asm.BASE_RSLT, i * rp)
        else:
            p.vse(8, asm.BASE_RSLT, (L + i) * rp)
    return p.brk()


class LocalComm:
    """world = 1: no exchange."""
    world, rank = 1, 0

    def all_gather_digits(self, ks):
        pass

    def broadcast_t(self, ks):
        pass


class TorchComm:
    """torch.distributed exchange.  On CUDA machines the collectives run in place on the SPM (NCCL over
    NVLink, on the machine's stream); on CPU machines (gloo, tests) they stage through host arrays."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)

    def _bind(self, machine):
        """The collectives below run on a torch stream, directly on SPM device memory: the engine must launch
        on that same stream or its kernels race with them (the engine's default is a private non-blocking
        stream).  Returns the stream to run the collective on: torch's current one, or -- when that is the
        legacy default stream, whose handle 0 would select the engine's own -- a stream this object owns."""
        import torch
        cur = torch.cuda.current_stream()
        if cur.cuda_stream == 0:
            if getattr(self, "_own_stream", None) is None:
                self._own_stream = torch.cuda.Stream()
            cur = self._own_stream
        if hasattr(machine, "set_stream") and getattr(machine, "_bound_stream", None) != cur.cuda_stream:
            machine.set_stream(cur.cuda_stream)
            machine._bound_stream = cur.cuda_stream
        return cur

    def _spm_tensor(self, machine, row, nrows):
        import torch

        class _View:
            pass
        v = _View()
        v.__cuda_array_interface__ = {"shape": (nrows * 128,), "typestr": "<i8", "version": 2,
                                      "data": (machine.spm_device_ptr(row), False)}
        return torch.as_tensor(v, device="cuda")

    def all_gather_digits(self, ks):
        import torch
        lay, m = ks.lay, ks.machine
        block = lay.per_rank * lay.rp
        if hasattr(m, "spm_device_ptr"):
            st = self._bind(m)
            m.spm_mark_written(lay.S, lay.slots * lay.rp)      # copy-on-write for aliased registers
            full = self._spm_tensor(m, lay.S, lay.slots * lay.rp)
            mine = full[self.rank * block * 128:(self.rank + 1) * block * 128]
            with torch.cuda.stream(st):
                self.dist.all_gather_into_tensor(full, mine, group=self.group)
        else:
            mine = torch.from_numpy(m.dma_mem_d2h(lay.S + self.rank * block, block * 128).view(np.int64))
            parts = [torch.empty_like(mine) for _ in range(self.world)]
            self.dist.all_gather(parts, mine, group=self.group)
            m.dma_mem_h2d(lay.S, np.concatenate([p.numpy() for p in parts]).view(np.uint64))

    def broadcast_t(self, ks):
        import torch
        lay, m = ks.lay, ks.machine
        row, nrows, src = lay.ACC + 2 * lay.L * lay.rp, 2 * lay.rp, lay.owner(lay.L)
        if hasattr(m, "spm_device_ptr"):
            st = self._bind(m)
            m.spm_mark_written(row, nrows)
            with torch.cuda.stream(st):
                self.dist.broadcast(self._spm_tensor(m, row, nrows), src=src, group=self.group)
        else:
            t = torch.from_numpy(m.dma_mem_d2h(row, nrows * 128).view(np.int64).copy())
            self.dist.broadcast(t, src=src, group=self.group)
            m.dma_mem_h2d(row, t.numpy().view(np.uint64))


class ShardedKeySwitch:
    """Runs the three phases on one machine (anything with the Engine method set) of a group."""

    def __init__(self, machine, lay: KeySwitchLayout, comm=None, pc_base: int = 0):
        self.machine, self.lay, self.comm = machine, lay, comm or LocalComm()
        assert self.comm.world == lay.world and self.comm.rank == lay.rank
        self.pc1, self.pc2, self.pc3 = {}, {}, {}
        pc = pc_base
        for i in lay.owned():
            for table, gen in ((self.pc1, phase1_stream), (self.pc2, phase2_stream), (self.pc3, phase3_stream)):
                if i == lay.L and gen is not phase2_stream:
                    continue
                words = gen(lay, i).words()
                machine.load_isram(words, pc)
                table[i] = pc
                pc += len(words)
        self.pc_end = pc

    def load_ksk(self, i: int, data: np.ndarray):
        """data: KSK[i] = 2L polynomials ([j][c] order) under modulus i; only the owner stores it."""
        assert self.lay.owner(i) == self.lay.rank and data.size == 2 * self.lay.L * self.lay.n
        self.machine.dma_ksk_h2d(self.lay.ksk_ptr(i), data.reshape(-1))

    def load_input(self, i: int, a: np.ndarray, b: np.ndarray):
        lay = self.lay
        self.machine.dma_mem_h2d(lay.IN + i * lay.rp, a)
        self.machine.dma_mem_h2d(lay.IN + (lay.L + i) * lay.rp, b)

    def run(self, galois_k: int):
        lay, m = self.lay, self.machine
        mine = lay.owned()
        m.run_vp_multi([(self.pc1[j], lay.IN, lay.S, 0, 0, galois_k) for j in mine if j < lay.L])
        self.comm.all_gather_digits(self)
        m.run_vp_multi([(self.pc2[i], lay.S, 0, lay.ACC, lay.ksk_ptr(i), 0) for i in mine])
        self.comm.broadcast_t(self)
        m.run_vp_multi([(self.pc3[i], lay.ACC, lay.S, lay.OUT, 0, 0) for i in mine if i < lay.L])

    def read_output(self, i: int):
        lay = self.lay
        return (self.machine.dma_mem_d2h(lay.OUT + i * lay.rp, lay.n),
                self.machine.dma_mem_d2h(lay.OUT + (lay.L + i) * lay.rp, lay.n))


def transform_count(L: int) -> int:
    """limb-(I)NTTs in one key-switch: 2L INTT + L NTT (phase 1), L(L+1) NTT + 2 INTT (phase 2), 2L NTT (phase 3)."""
    return 3 * L + L * (L + 1) + 2 + 2 * L
