"""ctypes wrapper around oracle/libaloha_oracle.so plus the host-driver replay loop.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs.  The product package (aloha_b200/) never imports this module.

Host flow restated from the reference testbench sim/top/top_noaxilite_tb.sv:
  parse_op :249-298, run_vp :396-417, run_encode :419-448, run_load_cipher :450-472,
  run_store_cipher :474-496, run_rotate :530-532, dump_poly :536-565, run :596-638.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libaloha_oracle.so")

# reference machine constants
LANES = 128
VLMAX_BITS = 524288            # src/vp/include/vp_defines.vh:24
SPM_ROWS = 16384               # src/mem_buf/spm.sv:47-178   (16 MiB)
KSK_ROWS = 9216                # src/mem_buf/ksk_mem.sv:12-16
N_TV = 8192
Q0, Q1, Q2 = 576460825317867521, 576460924102115329, 576462951330889729
PSI0, PSI1, PSI2 = 3825716582911, 79932510954937, 101017252977188  # tf_rom_generator.sv:75
ISRAM_ENCODE_POST, ISRAM_MUL_PLAIN, ISRAM_HOM_ADD, ISRAM_KEYSWITCH = 0, 64, 160, 256  # tb:63-66
DRAM_VP_BASE = 10485760        # tb:45
KSK_DRAM_BASE = 524288         # vivado_prj/top_noaxilite.xpr:1448-1450


def build(force: bool = False) -> str:
    """ALOHA_ORACLE_NATIVE=1 (set by bench.py's CPU legs): the -march=native build, compiled on this machine."""
    native = bool(os.environ.get("ALOHA_ORACLE_NATIVE"))
    path = os.path.join(_HERE, "libaloha_oracle_native.so") if native else _LIB_PATH
    src = [os.path.join(_HERE, f) for f in ("golden_model.cpp", "golden_model.h", "Makefile")]
    if force or not os.path.exists(path) or any(os.path.getmtime(s) > os.path.getmtime(path) for s in src):
        if not native:
            subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []))
        else:
            # several ranks of one job may get here at once: each builds into its own file and renames it into
            # place (atomic), so nobody ever loads a half-written library
            tmp = f"libaloha_oracle_native.{os.getpid()}.tmp.so"
            try:
                subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "native", f"NATIVE_OUT={tmp}"])
                os.replace(os.path.join(_HERE, tmp), path)
            except Exception:
                path = build_portable()
    return path


def build_portable() -> str:
    src = [os.path.join(_HERE, f) for f in ("golden_model.cpp", "golden_model.h", "Makefile")]
    if not os.path.exists(_LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        u64, u32, p64, p8, p32, vp = (C.c_uint64, C.c_uint32, C.POINTER(C.c_uint64),
                                      C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.c_void_p)
        sig = {
            "gm_create": (vp, [u64, u32, u32]),
            "gm_destroy": (None, [vp]),
            "gm_set_moduli": (C.c_int, [vp, p64, p64, u32]),
            "gm_load_isram": (C.c_int, [vp, p8, u32, u32]),
            "gm_dma_mem_h2d": (C.c_int, [vp, u32, p64, u64]),
            "gm_dma_mem_d2h": (C.c_int, [vp, p64, u32, u64]),
            "gm_dma_ksk_h2d": (C.c_int, [vp, u32, p64, u64]),
            "gm_spm_written": (C.c_int, [vp, u32, u64, p8]),
            "gm_run_vp": (C.c_int, [vp, u32, u32, u32, u32, u32, u32]),
            "gm_last_inst_count": (u32, [vp]),
            "gm_vreg_read": (C.c_int, [vp, u32, p64, u64]),
            "gm_vreg_write": (C.c_int, [vp, u32, p64, u64]),
            "gm_get_csr": (C.c_int, [vp, p64, p64, p64]),
            "gm_decode": (C.c_int, [p8, u64, p64]),
            "gm_barrett": (u64, [u64, u64, u64, u64]),
            "gm_half": (u64, [u64, u64]),
            "gm_alu": (u64, [u32, u64, u64, u64, u64, u64, p64]),
            "gm_barrett_iq": (u64, [u64]),
            "gm_powmod": (u64, [u64, u64, u64]),
            "gm_min_primitive_root": (u64, [u64, u64]),
            "gm_ntt": (C.c_int, [p64, p64, u64, u64, u64, u64, C.c_int]),
            "gm_ntt_tables_create": (vp, [u64, p64, p64, u32]),
            "gm_ntt_tables_destroy": (None, [vp]),
            "gm_ntt_batch": (C.c_int, [vp, p64, p32, u64, C.c_int, u32]),
            "gm_automorph": (C.c_int, [p64, p64, u64, u64, u64]),
            "gm_aut_mac_batch": (C.c_int, [p64, p64, p64, u64, u64, p64, p32, u64, u32]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def _p64(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_uint64))


def _p8(a: np.ndarray):
    assert a.dtype == np.uint8 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def _p32(a: np.ndarray):
    assert a.dtype == np.uint32 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_uint32))


class OracleError(RuntimeError):
    pass


def _ck(rc: int, what: str):
    if rc != 0:
        raise OracleError(f"{what}: oracle error {rc}")


def parse_mem_words(text: str) -> np.ndarray:
    """24-hex-digit-per-line $readmemh text -> (n, 12) uint8, byte 0 most significant."""
    rows = []
    for line in text.split():
        line = line.strip()
        if not line or line.startswith("//"):
            continue
        assert len(line) == 24, line
        rows.append(bytes.fromhex(line))
    return np.frombuffer(b"".join(rows), dtype=np.uint8).reshape(-1, 12).copy()


def barrett_iq(q: int) -> int:
    return (1 << 121) // q


class GoldenModel:
    """One ALOHA VP instance (SPM + KSK memory + 32 vregs + CSRs)."""

    def __init__(self, vlmax_bits=VLMAX_BITS, spm_rows=SPM_ROWS, ksk_rows=KSK_ROWS,
                 moduli=((Q0, PSI0), (Q1, PSI1), (Q2, PSI2))):
        self.L = lib()
        self.h = self.L.gm_create(vlmax_bits, spm_rows, ksk_rows)
        if not self.h:
            raise OracleError("gm_create failed")
        self.nmax = vlmax_bits // 64
        q = np.array([m[0] for m in moduli], dtype=np.uint64)
        psi = np.array([m[1] for m in moduli], dtype=np.uint64)
        _ck(self.L.gm_set_moduli(self.h, _p64(q), _p64(psi), len(q)), "set_moduli")

    def __del__(self):
        if getattr(self, "h", None):
            self.L.gm_destroy(self.h)
            self.h = None

    def load_isram(self, words: np.ndarray, at_pc: int):
        words = np.ascontiguousarray(words, dtype=np.uint8)
        _ck(self.L.gm_load_isram(self.h, _p8(words), len(words), at_pc), "load_isram")

    def dma_mem_h2d(self, spm_row: int, data: np.ndarray):
        data = np.ascontiguousarray(data, dtype=np.uint64)
        _ck(self.L.gm_dma_mem_h2d(self.h, spm_row, _p64(data), data.nbytes), "dma_mem_h2d")

    def dma_mem_d2h(self, spm_row: int, nwords: int) -> np.ndarray:
        out = np.empty(nwords, dtype=np.uint64)
        _ck(self.L.gm_dma_mem_d2h(self.h, _p64(out), spm_row, out.nbytes), "dma_mem_d2h")
        return out

    def dma_ksk_h2d(self, ksk_row: int, data: np.ndarray):
        data = np.ascontiguousarray(data, dtype=np.uint64)
        _ck(self.L.gm_dma_ksk_h2d(self.h, ksk_row, _p64(data), data.nbytes), "dma_ksk_h2d")

    def spm_written(self, spm_row: int, nwords: int) -> np.ndarray:
        out = np.empty(nwords, dtype=np.uint8)
        _ck(self.L.gm_spm_written(self.h, spm_row, nwords, _p8(out)), "spm_written")
        return out.astype(bool)

    def run_vp(self, pc, src0=0, src1=0, rslt=0, ksk_ptr=0, step=0):
        _ck(self.L.gm_run_vp(self.h, pc, src0, src1, rslt, ksk_ptr, step), f"run_vp(pc={pc})")

    def run_vp_multi(self, calls):
        for c in calls:
            self.run_vp(*c)

    def run_vp_batch(self, pc, calls):
        for c in calls:
            self.run_vp(pc, *c)

    def sync(self):
        pass

    def vreg_read(self, reg: int, n: int) -> np.ndarray:
        out = np.empty(n, dtype=np.uint64)
        _ck(self.L.gm_vreg_read(self.h, reg, _p64(out), n), "vreg_read")
        return out

    def vreg_write(self, reg: int, data: np.ndarray):
        data = np.ascontiguousarray(data, dtype=np.uint64)
        _ck(self.L.gm_vreg_write(self.h, reg, _p64(data), len(data)), "vreg_write")


# ------------------------------------------------------------------ stateless helpers
def decode(word12: bytes | np.ndarray, csr_step: int = 0) -> list[int]:
    w = np.frombuffer(bytes(word12), dtype=np.uint8).copy()
    out = np.zeros(17, dtype=np.uint64)
    _ck(lib().gm_decode(_p8(w), csr_step, _p64(out)), "decode")
    return [int(x) for x in out]


def alu(op: int, a: int, b: int, s: int, q: int, iq: int) -> tuple[int, int]:
    r1 = C.c_uint64(0)
    r0 = lib().gm_alu(op, a, b, s, q, iq, C.byref(r1))
    return int(r0), int(r1.value)


def ntt(a: np.ndarray, q: int, psi: int, inverse: bool = False) -> np.ndarray:
    a = np.array(a, dtype=np.uint64, copy=True)
    scratch = np.empty_like(a)
    _ck(lib().gm_ntt(_p64(a), _p64(scratch), len(a), q, barrett_iq(q), psi, int(inverse)), "ntt")
    return a


class NttTables:
    def __init__(self, n: int, q: list[int], psi: list[int]):
        self.n = n
        self.q = np.array(q, dtype=np.uint64)
        self.psi = np.array(psi, dtype=np.uint64)
        self.h = lib().gm_ntt_tables_create(n, _p64(self.q), _p64(self.psi), len(q))
        if not self.h:
            raise OracleError("ntt_tables_create failed")

    def __del__(self):
        if getattr(self, "h", None):
            lib().gm_ntt_tables_destroy(self.h)
            self.h = None

    def batch(self, a: np.ndarray, mod_idx: np.ndarray, inverse=False, nthreads=1) -> np.ndarray:
        """a: (count, n) uint64, transformed in place and returned."""
        assert a.dtype == np.uint64 and a.flags.c_contiguous and a.shape[-1] == self.n
        mod_idx = np.ascontiguousarray(mod_idx, dtype=np.uint32)
        count = a.size // self.n
        assert len(mod_idx) == count
        _ck(lib().gm_ntt_batch(self.h, _p64(a), _p32(mod_idx), count, int(inverse), nthreads),
            "ntt_batch")
        return a


def automorph(x: np.ndarray, k: int, q: int) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.uint64)
    out = np.empty_like(x)
    _ck(lib().gm_automorph(_p64(out), _p64(x), len(x), k, q), "automorph")
    return out


def aut_mac_batch(acc, x, p, k, q: np.ndarray, mod_idx: np.ndarray, nthreads=1):
    n = acc.shape[-1]
    count = acc.size // n
    q = np.ascontiguousarray(q, dtype=np.uint64)
    mod_idx = np.ascontiguousarray(mod_idx, dtype=np.uint32)
    _ck(lib().gm_aut_mac_batch(_p64(acc), _p64(x), _p64(p), n, k, _p64(q), _p32(mod_idx), count,
                               nthreads), "aut_mac_batch")
    return acc


def synthetic_primes(count: int, two_n: int, below: int = 1 << 60) -> list[int]:
    """the first `count` primes scanning downward from 2^60 with q = 1 mod 2N."""
    def is_prime(n: int) -> bool:
        if n < 2:
            return False
        for p in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
            if n % p == 0:
                return n == p
        d, s = n - 1, 0
        while d % 2 == 0:
            d //= 2
            s += 1
        for a in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
            x = pow(a, d, n)
            if x in (1, n - 1):
                continue
            for _ in range(s - 1):
                x = x * x % n
                if x == n - 1:
                    break
            else:
                return False
        return True

    out = []
    q = (below - 1) // two_n * two_n + 1
    while len(out) < count:
        if q < below and is_prime(q):
            out.append(q)
        q -= two_n
    return out


def min_primitive_root(q: int, two_n: int) -> int:
    return int(lib().gm_min_primitive_root(q, two_n))


# ------------------------------------------------------------------ host driver (testbench flow)
@dataclass
class HostOp:
    kind: str            # load_cipher store_cipher encode mul_plain hom_add rotate
    spm_addr: int
    dram_addr: int = 0
    src1: int = 0
    src2: int = 0
    step: int = 0


_OPS = {1: "load_cipher", 2: "store_cipher", 3: "encode", 4: "encode_post", 5: "mul_plain",
        6: "hom_add", 7: "rotate"}


def parse_program(text: str) -> list[HostOp]:
    """PROGRAM file: one op per line, three hex u32 (top_noaxilite_tb.sv:249-298,350-370)."""
    ops = []
    for line in text.splitlines():
        line = line.strip()
        if not line:
            continue
        a0, a1, a2 = (int(x, 16) for x in line.split(","))
        kind = _OPS[(a0 >> 28) & 0xF]
        spm = a0 & 0x3FFF
        if kind in ("load_cipher", "store_cipher", "encode"):
            ops.append(HostOp(kind, spm, dram_addr=(a1 << 32) | a2))
        elif kind == "rotate":
            ops.append(HostOp(kind, spm, step=a1 & 0x3FFF, src1=a2 & 0x3FFF))
        else:
            ops.append(HostOp(kind, spm, src1=a1 & 0x3FFF, src2=a2 & 0x3FFF))
    return ops


def clog2(x: int) -> int:
    return max(0, (x - 1).bit_length())


def replay(model, ops: list[HostOp], dram: np.ndarray, encoder_out: dict[int, np.ndarray],
           n: int = N_TV):
    """Run a host program on `model` (anything exposing the GoldenModel methods).  Yields
    (op_index, sub_id_or_None, data[4n] uint64, written[4n] bool) per dump, in the TB's order.

    dram: uint64 view of the DDR image.  encoder_out[i]: the 2n words the (out-of-scope) encoder
    wrote for op i -- injected from rtl_result/inst_<i>_0_out.txt ('Encoder not
    reproducible')."""
    words = 4 * n
    for i, op in enumerate(ops):
        if op.kind == "load_cipher":
            base = (DRAM_VP_BASE + op.dram_addr) // 8
            model.dma_mem_h2d(op.spm_addr, dram[base:base + words])
        elif op.kind == "store_cipher":
            base = (DRAM_VP_BASE + op.dram_addr) // 8
            dram[base:base + words] = model.dma_mem_d2h(op.spm_addr, words)
            yield i, None, dram[base:base + words].copy(), model.spm_written(op.spm_addr, words)
            continue
        elif op.kind == "encode":
            model.dma_mem_h2d(op.spm_addr, encoder_out[i])
            yield (i, 0, model.dma_mem_d2h(op.spm_addr, words),
                   model.spm_written(op.spm_addr, words))
            model.run_vp(ISRAM_ENCODE_POST, op.spm_addr, 0, op.spm_addr, 0, 0)
        elif op.kind == "mul_plain":
            model.run_vp(ISRAM_MUL_PLAIN, op.src1, op.src2, op.spm_addr, 0, 0)
        elif op.kind == "hom_add":
            model.run_vp(ISRAM_HOM_ADD, op.src1, op.src2, op.spm_addr, 0, 0)
        elif op.kind == "rotate":
            step = pow(3, op.step, 2 * n)
            ksk_ptr = (clog2(op.step) - 1) * n * 12 // LANES
            model.run_vp(ISRAM_KEYSWITCH, op.src1, 0, op.spm_addr, ksk_ptr, step)
        else:
            raise OracleError(f"op {op.kind} not supported")
        yield i, None, model.dma_mem_d2h(op.spm_addr, words), model.spm_written(op.spm_addr, words)
