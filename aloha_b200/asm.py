"""Assembler for ALOHA's 96-bit R-type HE instructions and generators for limb-batched streams.

Encoding (reference: src/vp/sequncer/expander.v:65-107,123-130): inst32 = funct6<<26 | m<<25 |
vs2<<20 | vs1<<15 | funct3<<12 | vd<<7 | 0x0b, followed by imm64.  The shipped microcode always
sets m = 1 and leaves unused register fields as the assembler that produced it did; this assembler
zeroes them (the decoder ignores them).

The reference ships four fixed kernels for L = 2 ciphertext primes, K = 1 special prime
(sim/vp/isram_file_generator/*.mem).  The generators below emit the same instruction patterns for
any number of RNS limbs -- SURVEY 8(f)1 -- and are what bench.py and the synthetic parity tests run.
"""
from __future__ import annotations

import numpy as np

F6 = dict(NOP=0x00, VFQMUL=0x01, VNTT=0x02, VLE=0x03, VSETVL=0x04, VFQADD=0x05, VINTT=0x06, VSE=0x07,
          VSETQ=0x08, VFQSUB=0x09, VSETIQ=0x0c, VFQMOD=0x0d, BREAK=0x10, VCPY=0x11, VAUT=0x15,
          VROLI=0x19)
VV, VS, SV = 0, 1, 2
BASE_SRC0, BASE_SRC1, BASE_RSLT, BASE_KSK = 0, 1, 2, 15


def word(funct6: int, vd=0, vs1=0, vs2=0, funct3=0, imm=0, m=1) -> bytes:
    inst = (funct6 << 26) | (m << 25) | (vs2 << 20) | (vs1 << 15) | (funct3 << 12) | (vd << 7) | 0x0b
    return inst.to_bytes(4, "big") + (imm & (2**64 - 1)).to_bytes(8, "big")


def barrett_iq(q: int) -> int:
    return (1 << 121) // q        # modmul_tb.sv:30-36


class Program:
    """Accumulates instruction words; `words()` gives the (n, 12) uint8 array load_isram takes."""

    def __init__(self):
        self.buf: list[bytes] = []

    def __len__(self):
        return len(self.buf)

    def words(self) -> np.ndarray:
        return np.frombuffer(b"".join(self.buf), dtype=np.uint8).reshape(-1, 12).copy()

    def hex(self) -> str:
        return "\n".join(w.hex() for w in self.buf)

    # config
    def vsetvl(self, n: int): self.buf.append(word(F6["VSETVL"], imm=n * 64)); return self
    def vsetq(self, q: int):
        self.buf.append(word(F6["VSETQ"], imm=q))
        self.buf.append(word(F6["VSETIQ"], imm=barrett_iq(q)))
        return self
    def brk(self): self.buf.append(word(F6["BREAK"])); return self
    # memory: row offset lives in imm[25:10], base selector in imm[63:48] (vp_top_full.sv:105-117)
    def vle(self, vd, base, row): self.buf.append(word(F6["VLE"], vd=vd, imm=(base << 48) | (row << 10))); return self
    def vse(self, vs1, base, row): self.buf.append(word(F6["VSE"], vs1=vs1, imm=(base << 48) | (row << 10))); return self
    # transforms / permutations
    def vntt(self, vd, vs1): self.buf.append(word(F6["VNTT"], vd=vd, vs1=vs1)); return self
    def vintt(self, vd, vs1): self.buf.append(word(F6["VINTT"], vd=vd, vs1=vs1)); return self
    def vaut(self, vd, vs1, imm=0): self.buf.append(word(F6["VAUT"], vd=vd, vs1=vs1, imm=imm)); return self
    def vroli(self, vd, vs1, imm): self.buf.append(word(F6["VROLI"], vd=vd, vs1=vs1, imm=imm)); return self
    # ALU
    def _alu(self, name, vd, vs1, vs2=None, imm=None, sv=False):
        if sv:
            self.buf.append(word(F6[name], vd=vd, vs2=vs2, funct3=SV, imm=imm))
        elif imm is not None:
            self.buf.append(word(F6[name], vd=vd, vs1=vs1, funct3=VS, imm=imm))
        else:
            assert (vs1 ^ vs2) & 1, "vv operands must sit in different banks (expander.v:183-200)"
            self.buf.append(word(F6[name], vd=vd, vs1=vs1, vs2=vs2, funct3=VV))
        return self
    def vfqmul(self, vd, vs1, vs2=None, imm=None): return self._alu("VFQMUL", vd, vs1, vs2, imm)
    def vfqadd(self, vd, vs1, vs2=None, imm=None): return self._alu("VFQADD", vd, vs1, vs2, imm)
    def vfqsub(self, vd, vs1, vs2=None, imm=None): return self._alu("VFQSUB", vd, vs1, vs2, imm)
    def vfqsub_sv(self, vd, imm, vs2): return self._alu("VFQSUB", vd, 0, vs2, imm, sv=True)
    def vfqmod(self, vd, vs1): self.buf.append(word(F6["VFQMOD"], vd=vd, vs1=vs1)); return self
    def vcpy(self, vd, vs1): self.buf.append(word(F6["VCPY"], vd=vd, vs1=vs1)); return self


def rows_per_poly(n: int) -> int:
    return n // 128


def transform_stream(n: int, moduli: list[int], inverse: bool = False) -> Program:
    """encode_post generalised to L limbs: limb l of src0 -> (I)NTT -> limb l of rslt.
    Pattern of encode_post.mem: VSETQ/IQ; VLE v0; VNTT v2 <- v0; VSE v2."""
    p = Program().vsetvl(n)
    rp = rows_per_poly(n)
    for l, q in enumerate(moduli):
        p.vsetq(q).vle(0, BASE_SRC0, l * rp)
        (p.vintt if inverse else p.vntt)(2, 0)
        p.vse(2, BASE_RSLT, l * rp)
    return p.brk()


def rotate_mac_stream(n: int, moduli: list[int]) -> Program:
    """Rotate-and-sum inner step per limb: rslt[l] = src1[l + L] + aut_k(src0[l]) * src1[l].
    src1 holds L plaintext limbs followed by L accumulator limbs; k comes from the step CSR."""
    p = Program().vsetvl(n)
    rp, L = rows_per_poly(n), len(moduli)
    for l, q in enumerate(moduli):
        p.vsetq(q)
        p.vle(0, BASE_SRC0, l * rp).vaut(2, 0)          # v2 = aut_k(x)
        p.vle(1, BASE_SRC1, l * rp).vfqmul(4, 2, 1)     # v4 = v2 * p
        p.vle(3, BASE_SRC1, (L + l) * rp).vfqadd(6, 4, 3)   # v6 = v4 + acc
        p.vse(6, BASE_RSLT, l * rp)
    # Retire the last limb's temporaries: what v2 and v4 hold at BREAK is architecturally visible, so the batcher
    # would have to materialise them -- three separate kernels for that limb instead of the fused one.  Reloading
    # them from the row just stored is an alias (no kernel), and the next call redefines both before it stores there.
    p.vle(2, BASE_RSLT, (L - 1) * rp).vle(4, BASE_RSLT, (L - 1) * rp)
    return p.brk()


def vaut_stream(n: int, moduli: list[int]) -> Program:
    """Standalone automorphism per limb: rslt[l] = aut_k(src0[l]); k comes from the step CSR
    (the VAUT steps of keyswitch.mem, insts 5, 20, ... on their own)."""
    p = Program().vsetvl(n)
    rp = rows_per_poly(n)
    for l, q in enumerate(moduli):
        p.vsetq(q).vle(0, BASE_SRC0, l * rp).vaut(2, 0).vse(2, BASE_RSLT, l * rp)
    return p.brk()


def elementwise_stream(n: int, moduli: list[int], op: str, scalar: int | None = None) -> Program:
    """mul_plain / hom_add generalised: rslt[l] = src0[l] (op) src1[l]  (or scalar)."""
    p = Program().vsetvl(n)
    rp = rows_per_poly(n)
    for l, q in enumerate(moduli):
        p.vsetq(q).vle(0, BASE_SRC0, l * rp)
        if scalar is None:
            p.vle(1, BASE_SRC1, l * rp)
            getattr(p, op)(2, 0, 1)
        else:
            getattr(p, op)(2, 0, imm=scalar)
        p.vse(2, BASE_RSLT, l * rp)
    return p.brk()
