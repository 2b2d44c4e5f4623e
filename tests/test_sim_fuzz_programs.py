"""Random host PROGRAMs (the testbench's op lists: load / encode / mul_plain / hom_add / rotate / store over the
shipped microcode, top_noaxilite_tb.sv:249-298,596-638) through the C host driver on the simulated device
(tests/sim_engine.py), blocking op by op and as one asynchronous range, under the PROGRAM-level flags; every dump
and its written-mask against the oracle's replay of the same list.  The reference ships three such programs; this
makes more, with operands that overlap, results written in place, and stores read back by later loads."""
import os
import random

import numpy as np
import pytest

import golden_util as G
import sim_engine
from oracle import oracle as O

N = 8192
CT = 256                      # rows of a ciphertext (4 polynomials of 64 rows)
STEPS = [2, 4, 8]


def gen_program(rng: random.Random, n_ops: int):
    lines, loads, enc = [], {}, {}
    cts, pts = [], []
    slot = lambda: rng.randrange(0, 40) * CT + (128 if rng.random() < 0.15 else 0)       # some half-ciphertext offsets
    dram_slot = lambda: rng.randrange(0, 16) * 4 * N * 8
    for i in range(n_ops):
        kinds = ["load"] + (["encode"] if len(cts) else []) + (["mul", "add", "rotate", "store"] if cts and i > 1 else [])
        kind = rng.choice(kinds) if cts else "load"
        if kind == "mul" and not pts:
            kind = "encode"
        if kind == "load":
            s, d = slot(), dram_slot()
            lines.append(f"{0x10000000 | s:08x},{d >> 32:08x},{d & 0xFFFFFFFF:08x}")
            loads[i] = d
            cts.append(s)
        elif kind == "encode":
            s = slot()
            lines.append(f"{0x30000000 | s:08x},00000000,00000000")
            enc[i] = None
            pts.append(s)
        elif kind == "mul":
            dst = rng.choice(cts + [slot()])
            lines.append(f"{0x50000000 | dst:08x},{rng.choice(cts):08x},{rng.choice(pts):08x}")
            cts.append(dst)
        elif kind == "add":
            dst = rng.choice(cts + [slot()])
            lines.append(f"{0x60000000 | dst:08x},{rng.choice(cts):08x},{rng.choice(cts):08x}")
            cts.append(dst)
        elif kind == "rotate":
            src = rng.choice(cts)
            dst = rng.choice([c for c in cts if abs(c - src) >= CT] + [slot()])
            if abs(dst - src) < CT:
                dst = (src + CT * 2) % (40 * CT)
            lines.append(f"{0x70000000 | dst:08x},{rng.choice(STEPS):08x},{src:08x}")
            cts.append(dst)
        else:
            d = dram_slot()
            lines.append(f"{0x20000000 | rng.choice(cts):08x},{d >> 32:08x},{d & 0xFFFFFFFF:08x}")
    return lines, loads, enc


def run_case(A, seed):
    rng = random.Random(seed)
    data = np.random.default_rng(seed)
    lines, loads, enc = gen_program(rng, rng.randrange(4, 12))
    text = "\n".join(lines)
    ops = O.parse_program(text)
    flags = rng.choice([0, A.F_DEFER, A.F_GRAPHS, A.F_DEFER | A.F_GRAPHS, A.F_NO_FUSE])
    dram0 = np.zeros(64 * 1024 * 1024 // 8, dtype=np.uint64)
    for d in set(loads.values()):
        base = (O.DRAM_VP_BASE + d) // 8
        dram0[base:base + 4 * N] = data.integers(0, O.Q0, 4 * N, dtype=np.uint64)
    enc = {i: data.integers(0, O.Q0, 2 * N, dtype=np.uint64) for i in enc}
    ksk = data.integers(0, O.Q0, 3 * 12 * N, dtype=np.uint64)                   # the keys of steps 2, 4, 8

    model = O.GoldenModel()
    for words, pc in G.microcode():
        model.load_isram(words, pc)
    model.dma_ksk_h2d(0, ksk)
    dram = dram0.copy()
    want = [(i, sub, d.copy(), w.copy()) for i, sub, d, w in O.replay(model, ops, dram, enc, N)]

    for mode in ("blocking", "async", "mixed"):
        eng = A.Engine(flags=flags)
        for words, pc in G.microcode():
            eng.load_isram(words, pc)
        eng.dma_ksk_h2d(0, ksk)
        host = A.HostDriver(eng, text, N)
        for i, e in enc.items():
            host.set_encoder_output(i, e)
        for d in set(loads.values()):
            base = (O.DRAM_VP_BASE + d) // 8
            host.dram_write(O.DRAM_VP_BASE + d, dram0[base:base + 4 * N])
        if mode == "blocking":
            per_op = [host.run_op(i) for i in range(len(ops))]
        elif mode == "async":
            per_op = host.run_all_async()
        else:                                             # ranges of random length, some asynchronous, some op by op
            per_op, i = [], 0
            while i < len(ops):
                cnt = min(len(ops) - i, rng.randrange(1, 5))
                if rng.random() < 0.5:
                    per_op += [[(s_, d_.copy(), w_.copy()) for s_, d_, w_ in dumps] for dumps in host.run_all_async(i, cnt)]
                else:
                    per_op += [host.run_op(j) for j in range(i, i + cnt)]
                i += cnt
        got = [(i, sub, d, w) for i, dumps in enumerate(per_op) for sub, d, w in dumps]
        assert [(i, s) for i, s, _, _ in got] == [(i, s) for i, s, _, _ in want], (seed, mode)
        for (i, sub, gd, gw), (_, _, wd, ww) in zip(got, want):
            assert (np.asarray(gw, bool) == ww).all(), (seed, mode, hex(flags), i, sub, "written-mask")
            assert (gd[ww] == wd[ww]).all(), (seed, mode, hex(flags), i, sub, text)
        for i, op in enumerate(ops):
            if op.kind == "store_cipher":
                base = (O.DRAM_VP_BASE + op.dram_addr) // 8
                assert (host.dram_read(O.DRAM_VP_BASE + op.dram_addr, 4 * N) == dram[base:base + 4 * N]).all(), (seed, mode, i)
        host.close()
        eng.close()


@pytest.mark.parametrize("block", range(4))
def test_random_host_programs(block):
    per = int(os.environ.get("ALOHA_PROGRAM_SEEDS", "3"))
    with sim_engine.simulated() as A:
        for seed in range(block * per, (block + 1) * per):
            run_case(A, seed)


def test_independent_ops_share_their_launches():
    """PROGRAM-level scheduling in the host driver: four encodes, four mul_plains and two rotates on rows of their own,
    issued as one asynchronous range, take the launches of ONE of each (the engine levels the batch), and every dump
    is still what the testbench would have written after each op."""
    lines = ["10000000,00000000,00000000"]                                                  # load ct -> 0x000
    # (encodes a whole ciphertext apart: the testbench dumps 4 polynomials after an encode, so plaintexts packed two
    #  polynomials apart, as in the tv cases, lie inside each other's dumps and cannot share a batch)
    lines += [f"{0x30000000 | (0x100 + 0x100 * i):08x},00000000,00000000" for i in range(4)]  # 4 encodes
    lines += [f"{0x50000000 | (0x600 + 0x100 * i):08x},00000000,{0x100 + 0x100 * i:08x}" for i in range(4)]   # 4 mul_plains
    lines += [f"{0x70000000 | (0xa00 + 0x100 * i):08x},{2 << i:08x},{0x600 + 0x100 * i:08x}" for i in range(2)]  # 2 rotates
    text = "\n".join(lines)
    ops = O.parse_program(text)
    data = np.random.default_rng(5)
    dram = np.zeros(64 * 1024 * 1024 // 8, dtype=np.uint64)
    base = O.DRAM_VP_BASE // 8
    dram[base:base + 4 * N] = data.integers(0, O.Q0, 4 * N, dtype=np.uint64)
    enc = {i: data.integers(0, O.Q0, 2 * N, dtype=np.uint64) for i in range(1, 5)}
    ksk = data.integers(0, O.Q0, 3 * 12 * N, dtype=np.uint64)
    model = O.GoldenModel()
    for words, pc in G.microcode():
        model.load_isram(words, pc)
    model.dma_ksk_h2d(0, ksk)
    want = [(i, sub, d.copy(), w.copy()) for i, sub, d, w in O.replay(model, ops, dram.copy(), enc, N)]
    with sim_engine.simulated() as A:
        launches = {}
        for mode in ("op by op", "range"):
            eng = A.Engine()
            for words, pc in G.microcode():
                eng.load_isram(words, pc)
            eng.dma_ksk_h2d(0, ksk)
            host = A.HostDriver(eng, text, N)
            for i, e in enc.items():
                host.set_encoder_output(i, e)
            host.dram_write(O.DRAM_VP_BASE, dram[base:base + 4 * N])
            per_op = [host.run_op(i) for i in range(len(ops))] if mode == "op by op" else host.run_all_async()
            got = [(i, sub, d, w) for i, dumps in enumerate(per_op) for sub, d, w in dumps]
            assert [(i, s) for i, s, _, _ in got] == [(i, s) for i, s, _, _ in want]
            for (i, sub, gd, gw), (_, _, wd, ww) in zip(got, want):
                assert (np.asarray(gw, bool) == ww).all() and (gd[ww] == wd[ww]).all(), (mode, i, sub)
            launches[mode] = eng.stats()["kernel_launches"]
            host.close()
            eng.close()
    assert launches["range"] * 2 < launches["op by op"], launches
