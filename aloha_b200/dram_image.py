"""`$readmemh` images of the testbench's DDR model.

The reference testbench loads its 64 MiB DDR from a text image with one 512-bit word per line
(`top_noaxilite_tb.sv:339-346`, DDR_DATA_WIDTH = 512: `$readmemh(DRAM_INPUT_FILE, mem_bank)`), and the
DMA moves it as a linear array of u64: u64 number j of a 512-bit word sits in bits [64j+63 : 64j]
(`dump_poly`, :536-565, reads `mem_bank[addr] >> bit_offset`).  In the hex text the most significant
digit comes first, so the LAST 16 digits of a line are u64 number 0.  `$readmemh` also accepts
`@hexaddr` lines (word address), `x`/`z` digits, `_` separators and // comments; unspecified words
stay at their previous value (zero here).

The shipped project references `dram_input_case3.mem` but does not ship it; `build_image` assembles
one from the tv/ text files the way the testbench's memory map lays them out (cleartexts from byte 0,
KSKs from 524 288 in 786 432-byte slots for steps 2 / 4 / 8, ciphertexts from DRAM_VP_BASE).
"""
from __future__ import annotations

import numpy as np

WORD_BYTES = 64                  # one 512-bit DDR word
U64_PER_WORD = 8
DRAM_VP_BASE = 10485760          # top_noaxilite_tb.sv:45
KSK_DRAM_BASE = 524288           # vivado_prj/top_noaxilite.xpr:1448-1450
KSK_SLOT_BYTES = 3 * 12 * 8192 * 8 // 3   # 12 polynomials of 8192 u64 per rotation key


def write_readmemh(path: str, data: np.ndarray, start_word: int = 0) -> None:
    """data: uint64 array, length a multiple of 8.  One 128-hex-digit line per 512-bit word."""
    data = np.ascontiguousarray(data, dtype=np.uint64)
    if len(data) % U64_PER_WORD:
        raise ValueError("length must be a multiple of 8 u64 (one 512-bit DDR word)")
    words = data.reshape(-1, U64_PER_WORD)
    with open(path, "w") as f:
        if start_word:
            f.write(f"@{start_word:x}\n")
        for w in words:
            f.write("".join(f"{int(v):016x}" for v in w[::-1]) + "\n")


def read_readmemh(path: str, total_words: int | None = None) -> np.ndarray:
    """-> uint64 array covering words [0, total_words) (or up to the last word the file sets)."""
    chunks: dict[int, np.ndarray] = {}
    addr = 0
    with open(path) as f:
        for raw in f:
            line = raw.split("//")[0].strip().replace("_", "")
            if not line:
                continue
            for tok in line.split():
                if tok.startswith("@"):
                    addr = int(tok[1:], 16)
                    continue
                tok = tok.lower().replace("x", "0").replace("z", "0").rjust(128, "0")
                if len(tok) > 128:
                    raise ValueError("word wider than 512 bits")
                vals = [int(tok[16 * (7 - j):16 * (8 - j)], 16) for j in range(U64_PER_WORD)]
                chunks[addr] = np.array(vals, dtype=np.uint64)
                addr += 1
    n = total_words if total_words is not None else (max(chunks) + 1 if chunks else 0)
    out = np.zeros(n * U64_PER_WORD, dtype=np.uint64)
    for a, v in chunks.items():
        if a < n:
            out[a * U64_PER_WORD:(a + 1) * U64_PER_WORD] = v
    return out


def build_image(dram_bytes: int, ciphertexts: dict[int, np.ndarray] | None = None,
                ksks: dict[int, np.ndarray] | None = None,
                cleartexts: dict[int, np.ndarray] | None = None) -> np.ndarray:
    """ciphertexts: {byte offset from DRAM_VP_BASE: 4N u64}; ksks: {rotation step: 12N u64};
    cleartexts: {byte address: raw 8-byte words of the encoder input}."""
    img = np.zeros(dram_bytes // 8, dtype=np.uint64)

    def put(byte_addr: int, arr: np.ndarray):
        arr = np.ascontiguousarray(arr, dtype=np.uint64)
        if byte_addr % 8 or byte_addr // 8 + len(arr) > len(img):
            raise ValueError("placement outside the DDR model")
        img[byte_addr // 8:byte_addr // 8 + len(arr)] = arr
    for addr, a in (cleartexts or {}).items():
        put(addr, a)
    for step, a in (ksks or {}).items():
        slot = max(0, (step - 1).bit_length() - 1)           # clog2(step) - 1: step 2 -> 0, 4 -> 1, 8 -> 2
        put(KSK_DRAM_BASE + slot * KSK_SLOT_BYTES, a)
    for off, a in (ciphertexts or {}).items():
        put(DRAM_VP_BASE + off, a)
    return img
