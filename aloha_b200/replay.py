"""Command-line replay of a host PROGRAM on the GPU engine -- the software twin of the reference's
full-system testbench run (sim/top/top_noaxilite_tb.sv:718-724: init(); run(); check_result()).

    python -m aloha_b200.replay --program case3.txt --isram isram_dir --ksk 2:ksk_step2.txt 8:ksk_step8.txt \\
        --cipher 0:ct0.txt --encoder 1:inst_1_0_out.txt 2:inst_2_0_out.txt --dump-dir out/ [--expect expected.txt]

Text inputs are the reference's formats: one decimal u64 (or `x`) per line for polynomials / keys
(tv/README.md), 24 hex digits per line for microcode (.mem), `a0,a1,a2` hex per line for the PROGRAM
(top_noaxilite_tb.sv:249-298).  Outputs are `inst_<i>_out.txt` / `inst_<i>_0_out.txt` exactly as
dump_poly / dump_sub_poly write them (:536-593).  `--cipher ADDR:file` places a 4-polynomial
ciphertext at DRAM_VP_BASE + ADDR; `--ksk STEP:file` loads a rotation key where the testbench's
load_ksk puts it ((clog2(step)-1)*12N/128 rows); `--encoder OP:file` injects the (out-of-scope)
encoder's SPM output for host op OP.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

from . import Engine, HostDriver

DRAM_VP_BASE = 10485760
KERNEL_PCS = (("encode_post", 0), ("mul_plain", 64), ("hom_add", 160), ("keyswitch", 256))


def read_poly_text(path: str) -> np.ndarray:
    with open(path) as f:
        return np.array([0 if t == "x" else int(t) for t in f.read().split()], dtype=np.uint64)


def read_mem_words(path: str) -> np.ndarray:
    rows = [bytes.fromhex(t) for t in open(path).read().split() if not t.startswith("//")]
    return np.frombuffer(b"".join(rows), dtype=np.uint8).reshape(-1, 12).copy()


def clog2(x: int) -> int:
    return max(0, (x - 1).bit_length())


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--program", required=True)
    ap.add_argument("--isram", required=True, help="directory with encode_post/mul_plain/hom_add/keyswitch .mem, or one isram_file.mem")
    ap.add_argument("--ksk", nargs="*", default=[], metavar="STEP:FILE")
    ap.add_argument("--cipher", nargs="*", default=[], metavar="DRAMADDR:FILE")
    ap.add_argument("--encoder", nargs="*", default=[], metavar="OP:FILE")
    ap.add_argument("--dram-image", help="$readmemh image of the whole DDR (512-bit words, top_noaxilite_tb.sv:339-346); "
                    "rotation keys are then DMA'd from it as the testbench's load_ksk does")
    ap.add_argument("--dump-dir", required=True)
    ap.add_argument("--expect", help="expected final result (case3_expected_result.txt format)")
    ap.add_argument("-n", type=int, default=8192)
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args(argv)

    eng = Engine(device=a.device)
    if os.path.isdir(a.isram):
        for name, pc in KERNEL_PCS:
            eng.load_isram(read_mem_words(os.path.join(a.isram, name + ".mem")), pc)
    else:
        eng.load_isram(read_mem_words(a.isram), 0)
    for item in a.ksk:
        step, path = item.split(":", 1)
        eng.dma_ksk_h2d((clog2(int(step)) - 1) * a.n * 12 // 128, read_poly_text(path))
    host = HostDriver(eng, open(a.program).read(), a.n)
    if a.dram_image:
        from . import dram_image as D
        img = D.read_readmemh(a.dram_image, total_words=(64 << 20) // 64)
        host.dram_write(0, img)
        if not a.ksk:      # load_ksk(KSK_DRAM_BASE): 3 keys x 12 N u64 in one DMA (tb:372-394, 715)
            n_ksk = 3 * 12 * a.n
            eng.dma_ksk_h2d(0, img[D.KSK_DRAM_BASE // 8:D.KSK_DRAM_BASE // 8 + n_ksk])
    for item in a.cipher:
        addr, path = item.split(":", 1)
        host.dram_write(DRAM_VP_BASE + int(addr, 0), read_poly_text(path))
    for item in a.encoder:
        op, path = item.split(":", 1)
        host.set_encoder_output(int(op), read_poly_text(path)[:2 * a.n])
    os.makedirs(a.dump_dir, exist_ok=True)
    last = None
    for i in range(len(host)):
        for sub, data, written in host.run_op(i):
            name = f"inst_{i}_out.txt" if sub is None else f"inst_{i}_{sub}_out.txt"
            HostDriver.write_dump_text(os.path.join(a.dump_dir, name), data, written)
            if sub is None:
                last = data
    print(f"replayed {len(host)} ops; {eng.stats()['kernel_launches']} kernel launches; dumps in {a.dump_dir}")
    if a.expect:
        want = read_poly_text(a.expect)
        ok = last is not None and len(want) == len(last) and bool((want == last).all())
        print("check_result:", "PASS" if ok else "FAIL")    # top_noaxilite_tb.sv:663-683
        return 0 if ok else 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
