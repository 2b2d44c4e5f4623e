"""CPU-only checks of the product's host side: the C-ABI library builds, loads and exports every
symbol include/aloha_b200.h declares; its decoder matches the reference's decode goldens and the
oracle's decoder; the assembler round-trips; and creating an engine without a GPU fails loudly
(there is no CPU fallback)."""
import json
import os
import random
import re

import numpy as np
import pytest

import golden_util as G
import aloha_b200 as A
from aloha_b200 import asm
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "aloha_b200.h")).read()
    declared = set(re.findall(r"\b(aloha_[a-z0-9_]+)\s*\(", header))
    declared -= {"aloha_t", "aloha_host_t"}
    lib = A.load_library()
    for name in sorted(declared):
        assert hasattr(lib, name), f"libaloha_b200.so does not export {name}"
    assert declared == set(A.EXPORTS)


@pytest.mark.parametrize("name", ["homo_add", "mul_plain", "inst_issue_test"])
def test_product_decoder_against_reference_goldens(name):
    g = json.load(open(os.path.join(G.GOLDEN, "decode", name + ".json")))
    bad = [i for i, (w, row) in enumerate(zip(g["words"], g["fields"]))
           if A.decode(bytes.fromhex(w)) != [int(x, 16) for x in row.split(",")]]
    # rows 15-16 of inst_issue_test: VFQSUB.sv, reference RTL vs reference golden conflict
    assert bad == ([15, 16] if name == "inst_issue_test" else [])


def test_product_decoder_equals_oracle_decoder_on_random_words():
    rng = random.Random(7)
    f6s = list(asm.F6.values()) + [0x3f, 0x12, 0x0a]
    for _ in range(5000):
        w = asm.word(rng.choice(f6s), vd=rng.randrange(32), vs1=rng.randrange(32), vs2=rng.randrange(32),
                     funct3=rng.randrange(4), imm=rng.getrandbits(64), m=rng.randrange(2))
        step = rng.getrandbits(14)
        assert A.decode(w, step) == O.decode(w, step), w.hex()


def test_assembler_reproduces_shipped_microcode_semantics():
    """Re-assemble encode_post from its disassembly; decoded fields must equal the shipped words'."""
    q0, q1 = O.Q0, O.Q1
    p = asm.Program().vsetvl(8192)
    for l, q in enumerate((q0, q1)):
        p.vsetq(q).vle(0, asm.BASE_SRC0, 64 * l).vntt(2, 0).vse(2, asm.BASE_RSLT, 64 * l)
    p.brk()
    shipped = G.microcode_words("encode_post")
    mine = p.words()
    assert len(mine) == len(shipped)
    for a, b in zip(mine, shipped):
        assert A.decode(bytes(a)) == A.decode(bytes(b))


def test_generated_streams_run_on_the_oracle():
    """transform_stream on 3 limbs at N=512 == per-limb oracle NTT (host logic, no GPU)."""
    n = 512
    primes = O.synthetic_primes(3, 1 << 17)
    psis = [O.min_primitive_root(q, 2 * n) for q in primes]
    m = O.GoldenModel(vlmax_bits=n * 64, spm_rows=64, ksk_rows=0, moduli=list(zip(primes, psis)))
    m.load_isram(asm.transform_stream(n, primes).words(), 0)
    rng = np.random.default_rng(5)
    x = np.stack([rng.integers(0, q, n, dtype=np.uint64) for q in primes])
    m.dma_mem_h2d(0, x.reshape(-1))
    m.run_vp(0, 0, 0, 32)
    got = m.dma_mem_d2h(32, 3 * n).reshape(3, n)
    for l, (q, psi) in enumerate(zip(primes, psis)):
        assert (got[l] == O.ntt(x[l], q, psi)).all()


def test_engine_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(A.AlohaError) as e:
        A.Engine()
    assert e.value.name == "E_CUDA" and "no CPU fallback" in str(e.value)


def test_dump_text_format(tmp_path):
    data = np.array([5, 0, 2**63 + 1, 7], dtype=np.uint64)
    wr = np.array([1, 0, 1, 1], dtype=np.uint8)
    path = str(tmp_path / "inst_0_out.txt")
    A.HostDriver.write_dump_text(path, data, wr)
    assert open(path).read() == "5\nx\n9223372036854775809\n7\n"


def test_readmemh_dram_image_roundtrip(tmp_path):
    """512-bit-per-line $readmemh images (top_noaxilite_tb.sv:339-346): u64 number 0 of a DDR word is the
    LAST 16 hex digits of its line; @address records, x digits and comments are honoured."""
    from aloha_b200 import dram_image as D
    rng = np.random.default_rng(0)
    data = rng.integers(0, 2**64, 64, dtype=np.uint64)
    path = str(tmp_path / "img.mem")
    D.write_readmemh(path, data)
    first = open(path).readline().strip()
    assert len(first) == 128 and int(first[-16:], 16) == int(data[0]) and int(first[:16], 16) == int(data[7])
    assert (D.read_readmemh(path) == data).all()
    with open(path, "a") as f:
        f.write("// a sparse record\n@10\n" + "0" * 112 + "00000000_0000xx2a\n")
    back = D.read_readmemh(path, total_words=20)
    assert (back[:64] == data).all() and back[16 * 8] == 0x2a and not back[64:16 * 8].any()
    ct = rng.integers(0, 2**60, 4 * 8192, dtype=np.uint64)
    ksk = rng.integers(0, 2**60, 12 * 8192, dtype=np.uint64)
    img = D.build_image(64 << 20, ciphertexts={0: ct}, ksks={8: ksk})
    assert (img[D.DRAM_VP_BASE // 8:D.DRAM_VP_BASE // 8 + len(ct)] == ct).all()
    base = (D.KSK_DRAM_BASE + 2 * D.KSK_SLOT_BYTES) // 8              # step 8 -> third slot
    assert (img[base:base + len(ksk)] == ksk).all() and D.KSK_SLOT_BYTES == 786432
