"""aloha_b200 -- B200-native execution engine for ALOHA's leveled-FHE vector datapath.

Python face of the C-ABI in include/aloha_b200.h (ctypes; no torch types cross the boundary).
`Engine` mirrors the reference testbench's host tasks one to one
(sim/top/top_noaxilite_tb.sv: run_vp :396-417, run_load_cipher :450-472, run_store_cipher :474-496,
load_ksk :372-394) and `HostDriver` its op-list replay (:249-298, :596-638).

There is no CPU fallback: importing works anywhere (so the build check can run), but creating an
Engine without the CUDA library or without an sm_100 device raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

__all__ = ["Engine", "Group", "HostDriver", "AlohaError", "load_library", "decode", "REFERENCE_MODULI",
           "LANES", "VLMAX_BITS", "SPM_ROWS", "KSK_ROWS"]

LANES = 128
VLMAX_BITS = 524288          # src/vp/include/vp_defines.vh:24
SPM_ROWS = 16384             # src/mem_buf/spm.sv:47-178
KSK_ROWS = 9216              # src/mem_buf/ksk_mem.sv:12-16
# sim/vp/tf_rom_generator/tf_rom_generator.sv:75-77  (q, psi)
REFERENCE_MODULI = ((576460825317867521, 3825716582911), (576460924102115329, 79932510954937),
                    (576462951330889729, 101017252977188))

F_NO_BATCH, F_NO_ALIAS, F_GRAPHS, F_NO_FUSE, F_STRICT, F_DEFER, F_GENERIC_MODMUL, F_AUT_GATHER, F_AUT_TILED = 1, 2, 4, 8, 16, 32, 64, 128, 256

_ERRORS = {-1: "E_ARG", -2: "E_RANGE", -3: "E_OPCODE", -4: "E_STATE", -5: "E_ILLEGAL", -6: "E_NOBREAK",
           -7: "E_UNDEFINED", -8: "E_CUDA", -9: "E_NOMEM"}


class AlohaError(RuntimeError):
    def __init__(self, code: int, what: str, detail: str = ""):
        self.code = code
        self.name = _ERRORS.get(code, str(code))
        super().__init__(f"{what}: {self.name} {detail}".strip())


class _Cfg(C.Structure):
    _fields_ = [("vlmax_bits", C.c_uint64), ("spm_rows", C.c_uint32), ("ksk_rows", C.c_uint32),
                ("device", C.c_int32), ("flags", C.c_uint32), ("pool_buffers", C.c_uint32),
                ("l2_chunk_bytes", C.c_uint64), ("isram_depth", C.c_uint32), ("reserved", C.c_uint32)]


class VpArgs(C.Structure):
    _fields_ = [("src0", C.c_uint32), ("src1", C.c_uint32), ("rslt", C.c_uint32),
                ("ksk_ptr", C.c_uint32), ("step", C.c_uint32)]


class _Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("kernel_launches", "instructions", "plans_built",
                                          "plans_reused", "copies_elided", "copies_emitted",
                                          "limb_ntts", "ops_fused")]


EXPORTS = ["aloha_create", "aloha_destroy", "aloha_strerror", "aloha_last_error", "aloha_load_isram",
           "aloha_load_tf_rom", "aloha_dma_mem_h2d", "aloha_dma_mem_d2h", "aloha_dma_ksk_h2d",
           "aloha_dma_mem_h2d_async", "aloha_dma_mem_d2h_async",
           "aloha_spm_written", "aloha_run_vp", "aloha_run_vp_batch", "aloha_run_vp_multi", "aloha_sync",
           "aloha_spm_device_ptr", "aloha_ksk_device_ptr", "aloha_spm_mark_written", "aloha_set_stream",
           "aloha_get_stats", "aloha_get_csr", "aloha_decode", "aloha_host_create",
           "aloha_host_destroy", "aloha_host_num_ops", "aloha_host_dram_write", "aloha_host_dram_read",
           "aloha_host_set_encoder_output", "aloha_host_run_op", "aloha_write_dump_text",
           "aloha_flush", "aloha_pinned_alloc", "aloha_pinned_free", "aloha_host_run_op_async", "aloha_host_run_range_async", "aloha_host_sync",
           "aloha_group_unique_id", "aloha_group_create", "aloha_group_create_local",
           "aloha_group_destroy", "aloha_group_size", "aloha_group_rank", "aloha_group_last_error",
           "aloha_group_all_gather_rows", "aloha_group_broadcast_rows", "aloha_group_wait"]

_lib = None


def load_library(rebuild: bool = False) -> C.CDLL:
    """dlopen aloha_b200/libaloha_b200.so (building it with nvcc first if it is missing/stale)."""
    global _lib
    if _lib is not None and not rebuild:
        return _lib
    path = _build.LIB
    if rebuild or not os.path.exists(path) or (_build.stale() and os.path.exists(_build.NVCC)):
        path = _build.build(force=rebuild)
    if not os.path.exists(path):
        raise AlohaError(-8, "load_library", f"{path} is missing and cannot be built (no nvcc); "
                         "the engine has no CPU fallback")
    _lib = bind_library(path)
    return _lib


def bind_library(path: str) -> C.CDLL:
    """dlopen `path` and declare the C-ABI of include/aloha_b200.h on it; every symbol must be there."""
    L = C.CDLL(path)
    u64, u32, vp = C.c_uint64, C.c_uint32, C.c_void_p
    p64, p8 = C.POINTER(C.c_uint64), C.POINTER(C.c_uint8)
    sig = {
        "aloha_create": (C.c_int, [C.POINTER(_Cfg), C.POINTER(vp)]),
        "aloha_destroy": (None, [vp]),
        "aloha_strerror": (C.c_char_p, [C.c_int]),
        "aloha_last_error": (C.c_char_p, [vp]),
        "aloha_load_isram": (C.c_int, [vp, p8, u32, u32]),
        "aloha_load_tf_rom": (C.c_int, [vp, p64, p64, u32]),
        "aloha_dma_mem_h2d": (C.c_int, [vp, u32, vp, u64]),
        "aloha_dma_mem_d2h": (C.c_int, [vp, vp, u32, u64]),
        "aloha_dma_ksk_h2d": (C.c_int, [vp, u32, vp, u64]),
        "aloha_dma_mem_h2d_async": (C.c_int, [vp, u32, vp, u64]),
        "aloha_dma_mem_d2h_async": (C.c_int, [vp, vp, u32, u64]),
        "aloha_spm_written": (C.c_int, [vp, u32, u64, p8]),
        "aloha_run_vp": (C.c_int, [vp, u32, u32, u32, u32, u32, u32]),
        "aloha_run_vp_batch": (C.c_int, [vp, u32, u32, C.POINTER(VpArgs)]),
        "aloha_run_vp_multi": (C.c_int, [vp, u32, C.POINTER(u32), C.POINTER(VpArgs)]),
        "aloha_sync": (C.c_int, [vp]),
        "aloha_spm_device_ptr": (C.c_int, [vp, u32, C.POINTER(vp)]),
        "aloha_ksk_device_ptr": (C.c_int, [vp, u32, C.POINTER(vp)]),
        "aloha_spm_mark_written": (C.c_int, [vp, u32, u32]),
        "aloha_set_stream": (C.c_int, [vp, vp]),
        "aloha_get_stats": (C.c_int, [vp, C.POINTER(_Stats)]),
        "aloha_get_csr": (C.c_int, [vp, p64, p64, p64]),
        "aloha_decode": (C.c_int, [p8, u64, p64]),
        "aloha_host_create": (C.c_int, [vp, C.c_char_p, u64, u32, C.POINTER(vp)]),
        "aloha_host_destroy": (None, [vp]),
        "aloha_host_num_ops": (C.c_int, [vp]),
        "aloha_host_dram_write": (C.c_int, [vp, u64, vp, u64]),
        "aloha_host_dram_read": (C.c_int, [vp, u64, vp, u64]),
        "aloha_host_set_encoder_output": (C.c_int, [vp, u32, p64, u64]),
        "aloha_host_run_op": (C.c_int, [vp, u32, p64, p8, p64, p8, C.POINTER(C.c_int)]),
        "aloha_write_dump_text": (C.c_int, [C.c_char_p, p64, p8, u64]),
        "aloha_flush": (C.c_int, [vp]),
        "aloha_pinned_alloc": (C.c_int, [u64, C.POINTER(vp)]),
        "aloha_pinned_free": (None, [vp]),
        "aloha_host_run_op_async": (C.c_int, [vp, u32, p64, p8, p64, p8, C.POINTER(C.c_int)]),
        "aloha_host_run_range_async": (C.c_int, [vp, u32, u32, p64, p8, p64, p8, C.POINTER(C.c_int)]),
        "aloha_host_sync": (C.c_int, [vp]),
        "aloha_group_unique_id": (C.c_int, [p8]),
        "aloha_group_create": (C.c_int, [vp, p8, C.c_int, C.c_int, C.POINTER(vp)]),
        "aloha_group_create_local": (C.c_int, [C.POINTER(vp), C.c_int, C.POINTER(vp)]),
        "aloha_group_destroy": (None, [vp]),
        "aloha_group_size": (C.c_int, [vp]),
        "aloha_group_rank": (C.c_int, [vp]),
        "aloha_group_last_error": (C.c_char_p, [vp]),
        "aloha_group_all_gather_rows": (C.c_int, [vp, u32, u32, u32, u32, u32]),
        "aloha_group_broadcast_rows": (C.c_int, [vp, u32, u32, C.c_int]),
        "aloha_group_wait": (C.c_int, [vp, C.c_int]),
    }
    assert sorted(sig) == sorted(EXPORTS)
    for name, (res, args) in sig.items():
        fn = getattr(L, name)   # raises AttributeError if the library lacks a declared symbol
        fn.restype, fn.argtypes = res, args
    return L


def _p64(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_uint64))


def _p8(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def decode(word12: bytes, csr_step: int = 0) -> list[int]:
    """17 micro-op fields of one 96-bit instruction word (sim/vp/sequncer/seq_top_tb.sv:138-160)."""
    w = np.frombuffer(bytes(word12), dtype=np.uint8).copy()
    out = np.zeros(17, dtype=np.uint64)
    rc = load_library().aloha_decode(_p8(w), csr_step, _p64(out))
    if rc:
        raise AlohaError(rc, "decode")
    return [int(x) for x in out]


class Engine:
    """One ALOHA vector processor on one GPU: SPM, KSK memory, 32 vregs and CSRs live in HBM."""

    def __init__(self, vlmax_bits=VLMAX_BITS, spm_rows=SPM_ROWS, ksk_rows=KSK_ROWS, device=0, flags=0,
                 moduli=REFERENCE_MODULI, pool_buffers=0, l2_chunk_bytes=0, isram_depth=0):
        self.L = load_library()
        self.h = C.c_void_p()
        self.nmax = vlmax_bits // 64
        cfg = _Cfg(vlmax_bits, spm_rows, ksk_rows, device, flags, pool_buffers, l2_chunk_bytes, isram_depth, 0)
        rc = self.L.aloha_create(C.byref(cfg), C.byref(self.h))
        if rc:
            detail = self.L.aloha_last_error(self.h).decode() if self.h else ""
            if self.h:
                self.L.aloha_destroy(self.h)
                self.h = None
            raise AlohaError(rc, "aloha_create", detail)
        if moduli:
            self.load_tf_rom([m[0] for m in moduli], [m[1] for m in moduli])

    def close(self):
        if getattr(self, "h", None):
            self.L.aloha_destroy(self.h)
            self.h = None

    __del__ = close

    def _ck(self, rc: int, what: str):
        if rc:
            raise AlohaError(rc, what, self.L.aloha_last_error(self.h).decode())

    # ---- provisioning
    def load_isram(self, words: np.ndarray, at_pc: int):
        words = np.ascontiguousarray(words, dtype=np.uint8)
        assert words.ndim == 2 and words.shape[1] == 12
        self._ck(self.L.aloha_load_isram(self.h, _p8(words), len(words), at_pc), "load_isram")

    def load_tf_rom(self, q, psi):
        q = np.array(q, dtype=np.uint64)
        psi = np.array(psi, dtype=np.uint64)
        self._ck(self.L.aloha_load_tf_rom(self.h, _p64(q), _p64(psi), len(q)), "load_tf_rom")

    # ---- DMA
    def dma_mem_h2d(self, spm_row: int, data):
        if isinstance(data, np.ndarray):
            data = np.ascontiguousarray(data, dtype=np.uint64)
            ptr, nbytes = data.ctypes.data, data.nbytes
        else:                      # (address, nbytes) of a pinned host buffer
            ptr, nbytes = data
        self._ck(self.L.aloha_dma_mem_h2d(self.h, spm_row, ptr, nbytes), "dma_mem_h2d")

    def dma_mem_d2h(self, spm_row: int, nwords: int, out=None) -> np.ndarray:
        if out is None:
            out = np.empty(nwords, dtype=np.uint64)
            ptr = out.ctypes.data
        else:
            ptr = out if isinstance(out, int) else out.ctypes.data
        self._ck(self.L.aloha_dma_mem_d2h(self.h, ptr, spm_row, nwords * 8), "dma_mem_d2h")
        return out

    def dma_mem_h2d_async(self, spm_row: int, host_ptr: int, nbytes: int):
        """host_ptr: address of a page-locked buffer that stays valid until sync()."""
        self._ck(self.L.aloha_dma_mem_h2d_async(self.h, spm_row, host_ptr, nbytes), "dma_mem_h2d_async")

    def dma_mem_d2h_async(self, host_ptr: int, spm_row: int, nbytes: int):
        self._ck(self.L.aloha_dma_mem_d2h_async(self.h, host_ptr, spm_row, nbytes), "dma_mem_d2h_async")

    def dma_ksk_h2d(self, ksk_row: int, data: np.ndarray):
        data = np.ascontiguousarray(data, dtype=np.uint64)
        self._ck(self.L.aloha_dma_ksk_h2d(self.h, ksk_row, data.ctypes.data, data.nbytes), "dma_ksk_h2d")

    def spm_written(self, spm_row: int, nwords: int) -> np.ndarray:
        out = np.empty(nwords, dtype=np.uint8)
        self._ck(self.L.aloha_spm_written(self.h, spm_row, nwords, _p8(out)), "spm_written")
        return out.astype(bool)

    # ---- execution
    def run_vp(self, pc, src0=0, src1=0, rslt=0, ksk_ptr=0, step=0):
        self._ck(self.L.aloha_run_vp(self.h, pc, src0, src1, rslt, ksk_ptr, step), f"run_vp(pc={pc})")

    @staticmethod
    def make_args(calls) -> "C.Array[VpArgs]":
        """calls: iterable of (src0, src1, rslt, ksk_ptr, step)."""
        calls = list(calls)
        arr = (VpArgs * len(calls))()
        for i, c in enumerate(calls):
            arr[i] = VpArgs(*c)
        return arr

    def run_vp_batch(self, pc: int, args):
        if not isinstance(args, C.Array):
            args = self.make_args(args)
        self._ck(self.L.aloha_run_vp_batch(self.h, pc, len(args), args), f"run_vp_batch(pc={pc})")

    def run_vp_multi(self, calls):
        """calls: iterable of (pc, src0, src1, rslt, ksk_ptr, step) -- one batch, one plan."""
        calls = list(calls)
        pcs = (C.c_uint32 * len(calls))(*[c[0] for c in calls])
        args = self.make_args([c[1:] for c in calls])
        self._ck(self.L.aloha_run_vp_multi(self.h, len(calls), pcs, args), "run_vp_multi")

    def sync(self):
        self._ck(self.L.aloha_sync(self.h), "sync")

    def flush(self):
        self._ck(self.L.aloha_flush(self.h), "flush")

    # ---- device-side access
    def spm_device_ptr(self, spm_row: int) -> int:
        p = C.c_void_p()
        self._ck(self.L.aloha_spm_device_ptr(self.h, spm_row, C.byref(p)), "spm_device_ptr")
        return p.value

    def ksk_device_ptr(self, ksk_row: int) -> int:
        p = C.c_void_p()
        self._ck(self.L.aloha_ksk_device_ptr(self.h, ksk_row, C.byref(p)), "ksk_device_ptr")
        return p.value

    def spm_mark_written(self, spm_row: int, nrows: int):
        self._ck(self.L.aloha_spm_mark_written(self.h, spm_row, nrows), "spm_mark_written")

    def set_stream(self, cuda_stream: int | None):
        self._ck(self.L.aloha_set_stream(self.h, cuda_stream), "set_stream")

    def stats(self) -> dict:
        s = _Stats()
        self._ck(self.L.aloha_get_stats(self.h, C.byref(s)), "get_stats")
        return {n: int(getattr(s, n)) for n, _ in _Stats._fields_}

    def csr(self) -> tuple[int, int, int]:
        vl, q, iq = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._ck(self.L.aloha_get_csr(self.h, C.byref(vl), C.byref(q), C.byref(iq)), "get_csr")
        return vl.value, q.value, iq.value


GROUP_CHUNKED, GROUP_ALL, GROUP_BCAST = 1, -1, -2


class Group:
    """Several engines (one per GPU) as one limb-sharded machine: NCCL transfers between their SPMs, inside
    the C library (aloha_group_*).  Group.create(engine, id, rank, nranks) for one process per GPU -- the
    128-byte id comes from Group.unique_id() on rank 0 and travels by whatever channel the ranks share --
    or Group.local([engines...]) for one process driving several GPUs."""

    def __init__(self, handle, lib, engines):
        self.h, self.L, self.engines = handle, lib, engines

    @staticmethod
    def unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        L = load_library()
        rc = L.aloha_group_unique_id(buf)
        if rc:
            raise AlohaError(rc, "group_unique_id", L.aloha_group_last_error(None).decode())
        return bytes(buf)

    @classmethod
    def create(cls, engine: Engine, uid: bytes, rank: int, nranks: int) -> "Group":
        L, h = engine.L, C.c_void_p()
        buf = (C.c_uint8 * 128).from_buffer_copy(uid)
        rc = L.aloha_group_create(engine.h, buf, rank, nranks, C.byref(h))
        if rc:
            detail = L.aloha_group_last_error(h).decode() if h else ""
            if h:
                L.aloha_group_destroy(h)
            raise AlohaError(rc, "group_create", detail)
        return cls(h, L, [engine])

    @classmethod
    def local(cls, engines) -> "Group":
        engines = list(engines)
        L, h = engines[0].L, C.c_void_p()
        arr = (C.c_void_p * len(engines))(*[e.h for e in engines])
        rc = L.aloha_group_create_local(arr, len(engines), C.byref(h))
        if rc:
            detail = L.aloha_group_last_error(h).decode() if h else ""
            if h:
                L.aloha_group_destroy(h)
            raise AlohaError(rc, "group_create_local", detail)
        return cls(h, L, engines)

    def close(self):
        if getattr(self, "h", None):
            self.L.aloha_group_destroy(self.h)
            self.h = None

    def _ck(self, rc, what):
        if rc:
            raise AlohaError(rc, what, self.L.aloha_group_last_error(self.h).decode())

    @property
    def size(self) -> int:
        return self.L.aloha_group_size(self.h)

    @property
    def rank(self) -> int:
        return self.L.aloha_group_rank(self.h)

    def all_gather_rows(self, spm_row: int, rows_per_rank: int, count: int = 1, stride_rows: int = 0, chunked=False):
        self._ck(self.L.aloha_group_all_gather_rows(self.h, spm_row, rows_per_rank, count, stride_rows,
                                                    GROUP_CHUNKED if chunked else 0), "group_all_gather_rows")

    def broadcast_rows(self, spm_row: int, nrows: int, root: int):
        self._ck(self.L.aloha_group_broadcast_rows(self.h, spm_row, nrows, root), "group_broadcast_rows")

    def wait(self, source: int = GROUP_ALL):
        self._ck(self.L.aloha_group_wait(self.h, source), "group_wait")


class HostDriver:
    """The testbench's host program runner (PROGRAM file + DDR image -> per-op dumps)."""

    def __init__(self, engine: Engine, program_text: str, n: int = 8192, dram_bytes: int = 64 << 20):
        self.eng, self.n, self.L = engine, n, engine.L
        self.h = C.c_void_p()
        rc = self.L.aloha_host_create(engine.h, program_text.encode(), dram_bytes, n, C.byref(self.h))
        if rc:
            raise AlohaError(rc, "host_create")

    def close(self):
        if getattr(self, "h", None):
            self.L.aloha_host_destroy(self.h)
            self.h = None
            for p in getattr(self, "_pinned_blocks", []):
                self.L.aloha_pinned_free(p)
            self._pinned_blocks, self._dump_pool = [], None

    __del__ = close

    def __len__(self):
        return self.L.aloha_host_num_ops(self.h)

    def dram_write(self, byte_addr: int, data: np.ndarray):
        data = np.ascontiguousarray(data)
        rc = self.L.aloha_host_dram_write(self.h, byte_addr, data.ctypes.data, data.nbytes)
        if rc:
            raise AlohaError(rc, "dram_write")

    def dram_read(self, byte_addr: int, nwords: int) -> np.ndarray:
        out = np.empty(nwords, dtype=np.uint64)
        rc = self.L.aloha_host_dram_read(self.h, byte_addr, out.ctypes.data, out.nbytes)
        if rc:
            raise AlohaError(rc, "dram_read")
        return out

    def set_encoder_output(self, op_index: int, data: np.ndarray):
        data = np.ascontiguousarray(data, dtype=np.uint64)
        rc = self.L.aloha_host_set_encoder_output(self.h, op_index, _p64(data), len(data))
        if rc:
            raise AlohaError(rc, "set_encoder_output")

    def run_op_nodump(self, i: int):
        """Run host op i without the testbench's per-op dump read-back (asynchronous)."""
        rc = self.L.aloha_host_run_op(self.h, i, None, None, None, None, None)
        if rc:
            raise AlohaError(rc, f"host_run_op({i})", self.L.aloha_last_error(self.eng.h).decode())

    def run_op(self, i: int):
        """-> [(sub_id or None, data[4n], written[4n])] in the order the TB writes its dump files."""
        w = 4 * self.n
        dump, sub = np.empty(w, np.uint64), np.empty(w, np.uint64)
        wr, swr = np.empty(w, np.uint8), np.empty(w, np.uint8)
        has_sub = C.c_int(0)
        rc = self.L.aloha_host_run_op(self.h, i, _p64(dump), _p8(wr), _p64(sub), _p8(swr), C.byref(has_sub))
        if rc:
            raise AlohaError(rc, f"host_run_op({i})", self.L.aloha_last_error(self.eng.h).decode())
        out = []
        if has_sub.value:
            out.append((0, sub, swr.astype(bool)))
        out.append((None, dump, wr.astype(bool)))
        return out

    def _pinned(self, nwords: int) -> np.ndarray:
        """a page-locked uint64 array owned by this driver (freed in close())"""
        p = C.c_void_p()
        rc = self.L.aloha_pinned_alloc(nwords * 8, C.byref(p))
        if rc:
            raise AlohaError(rc, "pinned_alloc")
        self._pinned_blocks.append(p)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint64)), shape=(nwords,))

    def run_all_async(self, first: int = 0, count: int | None = None):
        """Ops [first, first+count) with every dump the testbench would write, read-backs overlapped with the
        following ops, one synchronisation at the end.  -> per op, the same list run_op returns.  The dump
        arrays are page-locked, owned by this driver and reused by the next call."""
        count = len(self) - first if count is None else count
        w = 4 * self.n
        if not hasattr(self, "_pinned_blocks"):
            self._pinned_blocks, self._dump_pool = [], None
        if self._dump_pool is None or self._dump_pool[0].shape[0] < count:
            self._dump_pool = (self._pinned(count * w).reshape(count, w), self._pinned(count * w).reshape(count, w),
                               np.empty((count, w), np.uint8), np.empty((count, w), np.uint8), (C.c_int * count)())
            d, s_, wr_, swr_, _ = self._dump_pool
            # the per-op views handed back below, made once: the arrays are reused call after call
            self._dump_views = [((0, s_[j], swr_[j].view(bool)), (None, d[j], wr_[j].view(bool))) for j in range(count)]
            self._dump_ptrs = (_p64(d), _p8(wr_), _p64(s_), _p8(swr_))
        dumps, subs, wr, swr, has_sub = self._dump_pool
        rc = self.L.aloha_host_run_range_async(self.h, first, count, *self._dump_ptrs, has_sub)
        if not rc:
            rc = self.L.aloha_host_sync(self.h)
        if rc:
            raise AlohaError(rc, "host_run_range_async", self.L.aloha_last_error(self.eng.h).decode())
        views = self._dump_views
        return [list(views[j]) if has_sub[j] else [views[j][1]] for j in range(count)]

    @staticmethod
    def write_dump_text(path: str, data: np.ndarray, written: np.ndarray):
        data = np.ascontiguousarray(data, dtype=np.uint64)
        wr = np.ascontiguousarray(written, dtype=np.uint8)
        rc = load_library().aloha_write_dump_text(path.encode(), _p64(data), _p8(wr), len(data))
        if rc:
            raise AlohaError(rc, "write_dump_text")
