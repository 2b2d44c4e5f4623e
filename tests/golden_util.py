"""Shared helpers for the parity tests: committed golden fixtures (tests/golden/, built by
tools/make_golden.py from the reference's tv/ data) and the replay checker that works on any
machine object exposing the GoldenModel method set (the oracle, or the CUDA engine wrapper)."""
from __future__ import annotations

import hashlib
import json
import os

import numpy as np

from oracle import oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_manifest = None
_pool = None


def manifest():
    global _manifest
    if _manifest is None:
        _manifest = json.load(open(os.path.join(GOLDEN, "manifest.json")))
    return _manifest


def pool(key: str) -> np.ndarray:
    global _pool
    if _pool is None:
        _pool = np.load(os.path.join(GOLDEN, "pool.npz"))
    return _pool[key]


def microcode():
    """[(words (n,12) uint8, pc)] for the four shipped kernels."""
    mc = json.load(open(os.path.join(GOLDEN, "microcode.json")))
    return [(O.parse_mem_words("\n".join(k["words"])), k["pc"]) for k in mc.values()]


def microcode_words(name: str):
    mc = json.load(open(os.path.join(GOLDEN, "microcode.json")))
    return O.parse_mem_words("\n".join(mc[name]["words"]))


def write_microcode_dir(path: str):
    """Materialise the kernels as <name>.mem files ($readmemh text), as the replay CLI reads them."""
    os.makedirs(path, exist_ok=True)
    for name, k in json.load(open(os.path.join(GOLDEN, "microcode.json"))).items():
        with open(os.path.join(path, name + ".mem"), "w") as f:
            f.write("\n".join(k["words"]) + "\n")
    return path


def poly_hashes(data: np.ndarray, written: np.ndarray, n: int):
    out = []
    for p in range(len(data) // n):
        d, m = data[p * n:(p + 1) * n], written[p * n:(p + 1) * n]
        if not m.any():
            out.append("x")
        elif not m.all():
            out.append("partial")
        else:
            out.append(hashlib.sha256(np.ascontiguousarray(d).tobytes()).hexdigest())
    return out


def case_inputs(case: str):
    """(ops, dram image, encoder injections, {ksk_row: array}) for one tv case."""
    m = manifest()
    e = m["cases"][case]
    n = m["n"]
    ops = O.parse_program("\n".join(e["program"]))
    dram = np.zeros(64 * 1024 * 1024 // 8, dtype=np.uint64)
    for i, key in e["loads"].items():
        base = (O.DRAM_VP_BASE + ops[int(i)].dram_addr) // 8
        dram[base:base + 4 * n] = pool(key)
    enc = {int(i): pool(key) for i, key in e["encoder"].items()}
    ksk = {int(row): pool(key) for row, key in e["ksk"].items()}
    return ops, dram, enc, ksk


def check_case(model, case: str):
    """Replay a tv case on `model`; assert every per-op dump matches the reference hashes."""
    m = manifest()
    n = m["n"]
    ops, dram, enc, ksk = case_inputs(case)
    for row, data in ksk.items():
        model.dma_ksk_h2d(row, data)
    want = m["cases"][case]["dumps"]
    seen = 0
    for i, sub, data, wr in O.replay(model, ops, dram, enc, n):
        name = f"inst_{i}_out" if sub is None else f"inst_{i}_{sub}_out"
        got = poly_hashes(data, wr, n)
        assert got == want[name], f"{case}/{name}: polys differing = " \
            f"{[p for p in range(4) if got[p] != want[name][p]]}"
        seen += 1
    assert seen == len(want)
    return seen


def run_kernel_vector(model, item):
    """One kernel-level software-model vector; returns (got_hashes, want_hashes)."""
    n = manifest()["n"]
    model.dma_mem_h2d(0, pool(item["src0"]))
    if "src1" in item:
        model.dma_mem_h2d(256, pool(item["src1"]))
    op = item["op"]
    if op == "rotate":
        model.dma_ksk_h2d(item["ksk_row"], pool(item["ksk"]))
        model.run_vp(O.ISRAM_KEYSWITCH, 0, 0, 512, item["ksk_row"], pow(3, item["step"], 2 * n))
    elif op == "mul_plain":
        model.run_vp(O.ISRAM_MUL_PLAIN, 0, 256, 512)
    elif op == "hom_add":
        model.run_vp(O.ISRAM_HOM_ADD, 0, 256, 512)
    elif op == "encode_post":
        model.run_vp(O.ISRAM_ENCODE_POST, 0, 0, 512)
    npoly = len(item["want"])
    data = model.dma_mem_d2h(512, npoly * n)
    return poly_hashes(data, np.ones(len(data), bool), n), item["want"]
