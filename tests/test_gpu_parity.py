"""GPU parity tests (run on the B200 box): the CUDA engine, called through the C-ABI, against
  (1) the committed reference golden fixtures -- every per-op RTL dump of the three tv/ cases and the
      shipped kernel-level vectors -- bit-exact;
  (2) the oracle (CPU golden model) on seeded synthetic inputs at sizes it finishes in seconds;
  (3) size-independent properties at the BASELINE.json sizes (N = 2^16, 32 limbs).
Nothing here reads /root/reference."""
import os

import numpy as np
import pytest

import golden_util as G
import aloha_b200 as A
from aloha_b200 import asm
from oracle import oracle as O

pytestmark = pytest.mark.gpu

FLAG_SETS = [0, A.F_NO_BATCH, A.F_NO_ALIAS, A.F_NO_BATCH | A.F_NO_ALIAS, A.F_GRAPHS, A.F_NO_FUSE, A.F_STRICT, A.F_DEFER,
             A.F_AUT_GATHER, A.F_AUT_TILED]


def tv_engine(flags=0):
    e = A.Engine(flags=flags)
    for words, pc in G.microcode():
        e.load_isram(words, pc)
    return e


# ------------------------------------------------------------------ (1) reference golden vectors
@pytest.mark.parametrize("flags", [0, A.F_DEFER, A.F_GRAPHS])
@pytest.mark.parametrize("case", ["case0_4_4", "case1_8_8", "case2_16_16"])
def test_tv_replay_with_asynchronous_dumps(case, flags):
    """PROGRAM-level scheduling: the whole op list issued without a blocking read-back per op
    (aloha_host_run_op_async + one aloha_host_sync); every dump the testbench writes, still bit-exact --
    twice, so the second pass runs on cached plans and a recycled ring."""
    m = G.manifest()
    n = m["n"]
    entry = m["cases"][case]
    eng = tv_engine(flags)
    ops, dram, enc, ksk = G.case_inputs(case)
    for row, data in ksk.items():
        eng.dma_ksk_h2d(row, data)
    host = A.HostDriver(eng, "\n".join(entry["program"]), n)
    for i, data in enc.items():
        host.set_encoder_output(i, data)
    first_masks = {}
    for rnd in range(2):
        for i, key in entry["loads"].items():          # the programs store their result over their input
            host.dram_write(O.DRAM_VP_BASE + ops[int(i)].dram_addr, G.pool(key))
        seen = 0
        for i, dumps in enumerate(host.run_all_async()):
            for sub, data, wr in dumps:
                name = f"inst_{i}_out" if sub is None else f"inst_{i}_{sub}_out"
                if rnd == 0:
                    first_masks[name] = wr.copy()
                # never-written words ('x') only exist the first time round: the second pass checks the values
                # under the first pass's masks
                assert G.poly_hashes(data, first_masks[name], n) == entry["dumps"][name], f"{case}/{name}/{rnd}"
                seen += 1
        assert seen == len(entry["dumps"])



@pytest.mark.parametrize("flags", FLAG_SETS)
@pytest.mark.parametrize("case,ndumps", [("case0_4_4", 10), ("case1_8_8", 19), ("case2_16_16", 37)])
def test_tv_replay_bit_exact(case, ndumps, flags):
    """Replay through the generic driver loop (oracle.replay) with the engine as the machine."""
    assert G.check_case(tv_engine(flags), case) == ndumps


@pytest.mark.parametrize("case", ["case0_4_4", "case1_8_8", "case2_16_16"])
def test_tv_replay_through_c_host_driver(case):
    """Same replay, driven by the product's own C++ host driver (aloha_host_*)."""
    m = G.manifest()
    n = m["n"]
    entry = m["cases"][case]
    eng = tv_engine()
    ops, dram, enc, ksk = G.case_inputs(case)
    for row, data in ksk.items():
        eng.dma_ksk_h2d(row, data)
    host = A.HostDriver(eng, "\n".join(entry["program"]), n)
    assert len(host) == len(ops)
    for i, key in entry["loads"].items():
        host.dram_write(O.DRAM_VP_BASE + ops[int(i)].dram_addr, G.pool(key))
    for i, data in enc.items():
        host.set_encoder_output(i, data)
    seen = 0
    for i in range(len(host)):
        for sub, data, wr in host.run_op(i):
            name = f"inst_{i}_out" if sub is None else f"inst_{i}_{sub}_out"
            assert G.poly_hashes(data, wr, n) == entry["dumps"][name], f"{case}/{name}"
            seen += 1
    assert seen == len(entry["dumps"])
    st = eng.stats()
    assert st["kernel_launches"] > 0 and st["copies_elided"] > 0


def test_kernel_level_vectors():
    eng = tv_engine()
    for item in G.manifest()["kernels"]:
        got, want = G.run_kernel_vector(eng, item)
        assert got == want, (item["case"], item["kernel"])


# ------------------------------------------------------------------ (2) oracle, synthetic
def synth(n, nlimbs):
    primes = O.synthetic_primes(nlimbs, 1 << 17)          # SURVEY 8(d)3 prime rule
    psis = [O.min_primitive_root(q, 2 * n) for q in primes]
    return primes, psis


@pytest.mark.parametrize("logn", [8, 9, 10, 11, 12, 13, 14, 15, 16])
def test_ntt_intt_vs_oracle(logn):
    n, L = 1 << logn, 3
    rp = n // 128
    primes, psis = synth(n, L)
    eng = A.Engine(vlmax_bits=n * 64, spm_rows=4 * L * rp, ksk_rows=0, moduli=list(zip(primes, psis)))
    eng.load_isram(asm.transform_stream(n, primes).words(), 0)
    eng.load_isram(asm.transform_stream(n, primes, inverse=True).words(), 1024)
    rng = np.random.default_rng(logn)
    x = np.stack([rng.integers(0, q, n, dtype=np.uint64) for q in primes])
    # adversarial rows: all q-1 (worst case for the lazy bounds), zeros, and values in [q, 2q)
    x[0, : n // 4] = primes[0] - 1
    x[1, : n // 4] = 0
    x[2, : n // 4] += np.uint64(primes[2])
    eng.dma_mem_h2d(0, x.reshape(-1))
    eng.run_vp(0, 0, 0, L * rp)
    f = eng.dma_mem_d2h(L * rp, L * n).reshape(L, n)
    tabs = O.NttTables(n, primes, psis)
    want = tabs.batch(x.copy(), np.arange(L))
    assert (f == want).all()
    eng.run_vp(1024, L * rp, 0, 2 * L * rp)
    back = eng.dma_mem_d2h(2 * L * rp, L * n).reshape(L, n)
    assert (back == tabs.batch(want.copy(), np.arange(L), inverse=True)).all()
    assert (back == x % np.array(primes, dtype=np.uint64)[:, None]).all()


def test_ntt_all_max_inputs_full_size():
    """Every coefficient q-1 at N = 2^16: the input that maximises every lazy intermediate."""
    n, L = 65536, 4
    rp = n // 128
    primes, psis = synth(n, L)
    eng = A.Engine(vlmax_bits=n * 64, spm_rows=2 * L * rp, ksk_rows=0, moduli=list(zip(primes, psis)))
    eng.load_isram(asm.transform_stream(n, primes).words(), 0)
    x = np.stack([np.full(n, q - 1, dtype=np.uint64) for q in primes])
    eng.dma_mem_h2d(0, x.reshape(-1))
    eng.run_vp(0, 0, 0, L * rp)
    got = eng.dma_mem_d2h(L * rp, L * n).reshape(L, n)
    assert (got == O.NttTables(n, primes, psis).batch(x.copy(), np.arange(L))).all()


@pytest.mark.parametrize("logn", [8, 9, 12, 13, 16])
@pytest.mark.parametrize("flags", [0, A.F_GENERIC_MODMUL])
def test_grouped_row_pass_and_modulus_forms(logn, flags):
    """19 polynomials x 3 moduli in one launch: per modulus 16 take the TMA-staged row pass and 3 the plain
    one; two moduli are pseudo-Mersenne (2^60 - d), one is a generic 60-bit prime, so both arithmetic
    forms run side by side (ALOHA_F_GENERIC_MODMUL forces the generic form for all three).  Every output
    word against the oracle, inputs include q-1 runs and words in [q, 2q)."""
    n, B = 1 << logn, 19
    rp = n // 128
    primes = O.synthetic_primes(2, 2 * n) + O.synthetic_primes(1, 2 * n, below=(1 << 60) - (1 << 40))
    assert (1 << 60) - primes[1] <= 1 << 27 < (1 << 60) - primes[2]
    psis = [O.min_primitive_root(q, 2 * n) for q in primes]
    L = len(primes)
    rows = B * L * rp
    eng = A.Engine(vlmax_bits=n * 64, spm_rows=2 * rows, ksk_rows=0, moduli=list(zip(primes, psis)), flags=flags)
    eng.load_isram(asm.transform_stream(n, primes).words(), 0)
    eng.load_isram(asm.transform_stream(n, primes, inverse=True).words(), 1024)
    rng = np.random.default_rng(1000 + logn)
    qv = np.array(primes, dtype=np.uint64)[None, :, None]
    x = rng.integers(0, 1 << 59, (B, L, n), dtype=np.uint64) % qv
    x[0, :, : n // 2] = qv[0, :, :] - np.uint64(1)
    x[1, :, n // 2:] += qv[0, :, :]
    x[18, :, ::3] = qv[0, :, :] - np.uint64(1)
    eng.dma_mem_h2d(0, x.reshape(-1))
    s0 = eng.stats()["kernel_launches"]
    eng.run_vp_batch(0, [(b * L * rp, 0, rows + b * L * rp, 0, 0) for b in range(B)])
    # columns (if any) + TMA-staged rows + plain rows, once per arithmetic form present in the launch
    forms = 1 if flags else 2
    assert eng.stats()["kernel_launches"] - s0 == forms * ((1 if logn > 8 else 0) + 2)
    F = eng.dma_mem_d2h(rows, B * L * n).reshape(B, L, n)
    tabs = O.NttTables(n, primes, psis)
    for b in range(B):
        assert (F[b] == tabs.batch(x[b].copy(), np.arange(L))).all(), b
    eng.run_vp_batch(1024, [(rows + b * L * rp, 0, b * L * rp, 0, 0) for b in range(B)])
    back = eng.dma_mem_d2h(0, B * L * n).reshape(B, L, n)
    assert (back == x % qv).all()


@pytest.mark.parametrize("n", [256, 2048])
@pytest.mark.parametrize("pre", ["vcpy", "vfqmod"])
def test_grouped_row_pass_with_folded_load_op(n, pre):
    """Base-extension pattern (VCPY / VFQMOD feeding VNTT) on 20 polynomials per modulus: the element-wise op
    is folded into the transform's load, and at N = 256 that load sits in the row pass itself -- here the
    TMA-staged one for 16 of the 20 and the plain one for the other 4.  Raw 64-bit input words."""
    B = 20
    rp = n // 128
    primes = O.synthetic_primes(1, 2 * n) + O.synthetic_primes(1, 2 * n, below=(1 << 60) - (1 << 40))
    psis = [O.min_primitive_root(q, 2 * n) for q in primes]
    L = len(primes)
    rows = B * L * rp
    prog = asm.Program().vsetvl(n)
    for l, q in enumerate(primes):
        prog.vsetq(q).vle(0, 0, l * rp)
        getattr(prog, pre)(3, 0)
        prog.vntt(2, 3).vse(2, 2, l * rp)
    prog.brk()
    machines = [O.GoldenModel(vlmax_bits=n * 64, spm_rows=2 * rows, ksk_rows=0, moduli=list(zip(primes, psis))),
                A.Engine(vlmax_bits=n * 64, spm_rows=2 * rows, ksk_rows=0, moduli=list(zip(primes, psis)))]
    rng = np.random.default_rng(n)
    # VFQMOD accepts any word; VCPY's two conditional subtracts leave words below 3q under 2q ... keep to < 2q there
    x = rng.integers(0, 1 << 64, B * L * n, dtype=np.uint64) if pre == "vfqmod" else \
        rng.integers(0, 2 * min(primes), B * L * n, dtype=np.uint64)
    out = []
    for m in machines:
        m.load_isram(prog.words(), 0)
        m.dma_mem_h2d(0, x)
        m.run_vp_batch(0, [(b * L * rp, 0, rows + b * L * rp, 0, 0) for b in range(B)])
        out.append(m.dma_mem_d2h(rows, B * L * n))
    assert (out[0] == out[1]).all()
    assert machines[1].stats()["ops_fused"] >= B          # the last limb's temporary stays live and is not folded


ALU_STREAMS = [("vfqmul", None), ("vfqadd", None), ("vfqsub", None), ("vfqmul", 0x123456789abcdef),
               ("vfqadd", (1 << 60) + 5), ("vfqsub", 77)]
ALU_CODE = {("vfqmul", False): 0x00, ("vfqadd", False): 0x01, ("vfqsub", False): 0x02,
            ("vfqmul", True): 0x04, ("vfqadd", True): 0x05, ("vfqsub", True): 0x06}


@pytest.mark.parametrize("op,scalar", ALU_STREAMS)
def test_elementwise_alu_vs_oracle_raw_words(op, scalar):
    """Inputs are RAW 64-bit words (not reduced): the engine must store what the RTL ALU would."""
    n, L = 1024, 2
    rp = n // 128
    primes = [O.Q0, O.Q2]
    eng = A.Engine(vlmax_bits=n * 64, spm_rows=3 * L * rp, ksk_rows=0, moduli=())
    eng.load_isram(asm.elementwise_stream(n, primes, op, scalar).words(), 0)
    rng = np.random.default_rng(11)
    a = rng.integers(0, 2**64, (L, n), dtype=np.uint64)
    b = rng.integers(0, 2**64, (L, n), dtype=np.uint64)
    a[:, :256] %= np.uint64(primes[0])
    b[:, :256] %= np.uint64(primes[0])
    eng.dma_mem_h2d(0, a.reshape(-1))
    eng.dma_mem_h2d(L * rp, b.reshape(-1))
    eng.run_vp(0, 0, L * rp, 2 * L * rp)
    got = eng.dma_mem_d2h(2 * L * rp, L * n).reshape(L, n)
    code = ALU_CODE[(op, scalar is not None)]
    for l, q in enumerate(primes):
        iq = O.barrett_iq(q)
        idx = list(range(0, n, 7))
        want = [O.alu(code, int(a[l, i]), int(b[l, i]), scalar or 0, q, iq)[0] for i in idx]
        assert [int(got[l, i]) for i in idx] == want


def run_single(eng, prog, n, inputs, out_rows):
    eng.load_isram(prog.words(), 0)
    for row, data in inputs:
        eng.dma_mem_h2d(row, data)
    eng.run_vp(0, 0, 0, out_rows)
    return eng.dma_mem_d2h(out_rows, n)


@pytest.mark.parametrize("n,kbits_n", [(8192, 8192), (4096, 8192), (65536, 65536), (65536, 131072)])
def test_vaut_vroli_vcpy_vfqmod_vs_oracle(n, kbits_n):
    """Permutation + base-extension primitives, incl. Galois elements >= N (k is truncated
    to log2(vlmax/64) bits, so on a vlmax = 64 N machine odd-i signs differ from mathematics)."""
    rp = n // 128
    q = O.Q0
    rng = np.random.default_rng(n)
    x = rng.integers(0, q, n, dtype=np.uint64)
    x[:16] = 0                                        # VAUT turns 0 into q (Q2)
    for step in (1, 2, 8, n // 4):
        k_csr = pow(3, step, 2 * n)
        eng = A.Engine(vlmax_bits=kbits_n * 64, spm_rows=4 * rp, ksk_rows=0, moduli=())
        p = asm.Program().vsetvl(n).vsetq(q).vle(0, 0, 0).vaut(2, 0).vse(2, 2, 0).brk()
        eng.load_isram(p.words(), 0)
        eng.dma_mem_h2d(0, x)
        eng.run_vp(0, 0, 0, rp, 0, k_csr)
        got = eng.dma_mem_d2h(rp, n)
        k_seen = k_csr & (kbits_n - 1)
        assert (got == O.automorph(x, k_seen, q)).all(), (n, step)
    eng = A.Engine(vlmax_bits=kbits_n * 64, spm_rows=4 * rp, ksk_rows=0, moduli=())
    for rot in (1, 129, n - 1):
        p = asm.Program().vsetvl(n).vsetq(q).vle(0, 0, 0).vroli(2, 0, rot).vse(2, 2, 0).brk()
        got = run_single(eng, p, n, [(0, x)], rp)
        assert (got == np.roll(x, -rot)).all()
    raw = rng.integers(0, 2**64, n, dtype=np.uint64)
    iq = O.barrett_iq(q)
    for name, code in (("vcpy", 0x05), ("vfqmod", 0x03)):
        p = asm.Program().vsetvl(n).vsetq(q).vle(0, 0, 0)
        getattr(p, name)(2, 0)
        p.vse(2, 2, 0).brk()
        got = run_single(eng, p, n, [(0, raw)], rp)
        idx = list(range(0, n, 97))
        assert [int(got[i]) for i in idx] == [O.alu(code, int(raw[i]), 0, 0, q, iq)[0] for i in idx]


def test_rotate_mac_stream_vs_oracle():
    n, L = 8192, 3
    rp = n // 128
    primes, psis = synth(n, L)
    k = pow(3, 4, 2 * n)
    eng = A.Engine(vlmax_bits=n * 64, spm_rows=4 * L * rp, ksk_rows=0, moduli=())
    eng.load_isram(asm.rotate_mac_stream(n, primes).words(), 0)
    rng = np.random.default_rng(2)
    x = np.stack([rng.integers(0, q, n, dtype=np.uint64) for q in primes])
    p = np.stack([rng.integers(0, q, n, dtype=np.uint64) for q in primes])
    acc = np.stack([rng.integers(0, q, n, dtype=np.uint64) for q in primes])
    eng.dma_mem_h2d(0, x.reshape(-1))
    eng.dma_mem_h2d(L * rp, np.concatenate([p.reshape(-1), acc.reshape(-1)]))
    eng.run_vp(0, 0, L * rp, 3 * L * rp, 0, k)
    got = eng.dma_mem_d2h(3 * L * rp, L * n).reshape(L, n)
    want = O.aut_mac_batch(acc.copy(), x, p, k, np.array(primes, dtype=np.uint64), np.arange(L))
    assert (got == want).all()
    # the batcher folds VAUT + VFQMUL + VFQADD of every limb into the gather-multiply-add kernel: the stream ends
    # by reloading v2 / v4, so not even the last limb's temporaries are architecturally visible
    assert eng.stats()["ops_fused"] == 2 * L
    assert eng.stats()["kernel_launches"] == 1


@pytest.mark.parametrize("flags", [0, A.F_AUT_GATHER, A.F_AUT_TILED])
@pytest.mark.parametrize("n", [256, 2048, 65536])
def test_vaut_every_kind_of_galois_element(n, flags):
    """The tiled permutation (aut_plan.hpp) and the 8-byte gather against the oracle for Galois elements
    whose lattices are balanced, extremely skewed (3, N/2+1, 1/3 mod N), the reversal 2N-1 and the identity
    permutation N+1 -- several in ONE launch, so tile counts differ between the launch's jobs."""
    rp = n // 128
    q = O.Q0
    rng = np.random.default_rng(n + flags)
    ks = [3, 9, pow(3, 8, 2 * n), pow(3, n // 8, 2 * n), 2 * n - 1, n + 1, n // 2 + 1, pow(3, -1, n), 12345 | 1, n - 3]
    x = rng.integers(0, q, (len(ks), n), dtype=np.uint64)
    x[:, :3] = 0
    eng = A.Engine(vlmax_bits=2 * n * 64, spm_rows=2 * len(ks) * rp, ksk_rows=0, moduli=(), flags=flags)
    eng.load_isram(asm.Program().vsetvl(n).vsetq(q).vle(0, 0, 0).vaut(2, 0).vse(2, 2, 0).brk().words(), 0)
    eng.dma_mem_h2d(0, x.reshape(-1))
    eng.run_vp_batch(0, [(i * rp, 0, (len(ks) + i) * rp, 0, k) for i, k in enumerate(ks)])
    got = eng.dma_mem_d2h(len(ks) * rp, len(ks) * n).reshape(len(ks), n)
    for i, k in enumerate(ks):
        assert (got[i] == O.automorph(x[i], k & (2 * n - 1), q)).all(), (n, k)
    assert eng.stats()["kernel_launches"] == 1


@pytest.mark.parametrize("flags", [0, A.F_AUT_GATHER, A.F_AUT_TILED])
def test_rotate_mac_full_size_vs_oracle(flags):
    """BASELINE.json configs[3] at its real size: N = 2^16, 32 limbs, fused aut-mul-add, every word of two
    polynomials' worth of limbs against the oracle, for a small, a pseudo-random and a >= N Galois element."""
    n, L, B = 65536, 32, 2
    rp = n // 128
    primes, psis = synth(n, L)
    qv = np.array(primes, dtype=np.uint64)
    per_poly = L * rp
    eng = A.Engine(vlmax_bits=n * 64, spm_rows=4 * B * per_poly, ksk_rows=0, moduli=(), pool_buffers=160, flags=flags)
    eng.load_isram(asm.rotate_mac_stream(n, primes).words(), 0)
    rng = np.random.default_rng(11)
    x, p, acc = (rng.integers(0, 1 << 59, (B, L, n), dtype=np.uint64) % qv[None, :, None] for _ in range(3))
    eng.dma_mem_h2d(0, x.reshape(-1))
    eng.dma_mem_h2d(B * per_poly, np.concatenate([p, acc], axis=1).reshape(-1))
    for step in (1, n // 8, 1000):
        k = pow(3, step, 2 * n)
        if step == 1000:
            k |= n                                      # a Galois element >= N with k mod N != 1 (SURVEY Q5 truncation)
        eng.run_vp_batch(0, [(b * per_poly, B * per_poly + 2 * b * per_poly, 3 * B * per_poly + b * per_poly, 0, k)
                             for b in range(B)])
        got = eng.dma_mem_d2h(3 * B * per_poly, B * L * n).reshape(B, L, n)
        k_seen = k & (n - 1)                            # vlmax/64 = n: the RTL keeps log2(n) bits of k
        for b in range(B):
            want = O.aut_mac_batch(acc[b].copy(), x[b], p[b], k_seen, qv, np.arange(L), nthreads=8)
            assert (got[b] == want).all(), (step, b)
    assert eng.stats()["ops_fused"] > 0


def test_fusion_keeps_in_place_rotate_accumulate_exact():
    """VLE v1 <- M; VAUT; (v1 redefined); MUL; ADD; VSE -> M in a batch of calls.  The fused
    gather-multiply-add must not read x from the range it writes."""
    n, L = 4096, 1
    rp = n // 128
    q = synth(n, 1)[0][0]
    calls = 3
    prog = (asm.Program().vsetvl(n).vsetq(q)
            .vle(1, 0, 0).vaut(2, 1)                    # v2 = aut(M)
            .vle(1, 1, 0)                               # v1 redefined: the alias of M is gone
            .vfqmul(4, 2, 1).vle(3, 1, rp).vfqadd(6, 4, 3)
            .vse(6, 0, 0).brk())                        # result back into M
    rng = np.random.default_rng(4)
    m0 = rng.integers(0, q, (calls, n), dtype=np.uint64)
    pa = rng.integers(0, q, (2, n), dtype=np.uint64)
    k = pow(3, 5, 2 * n)
    outs = []
    for flags in (0, A.F_NO_FUSE | A.F_NO_BATCH, A.F_AUT_GATHER, A.F_AUT_TILED):
        eng = A.Engine(vlmax_bits=n * 64, spm_rows=(calls + 2) * rp, ksk_rows=0, moduli=(), flags=flags, pool_buffers=34)
        eng.load_isram(prog.words(), 0)
        eng.dma_mem_h2d(0, m0.reshape(-1))
        eng.dma_mem_h2d(calls * rp, pa.reshape(-1))
        for _ in range(2):                              # second round: temporaries of the first are dead
            eng.run_vp_batch(0, [(c * rp, calls * rp, 0, 0, k) for c in range(calls)])
        outs.append(eng.dma_mem_d2h(0, calls * n))
    qv = np.array([q], dtype=np.uint64)
    want = m0.copy()
    for _ in range(2):
        for c in range(calls):
            want[c] = O.aut_mac_batch(pa[1:2].copy(), want[c:c + 1], pa[0:1], k & (n - 1), qv, np.zeros(1, dtype=np.int64))[0]
    for o in outs:
        assert (o.reshape(calls, n) == want).all()


# ------------------------------------------------------------------ batcher / architectural state
def test_batch_equals_loop_and_plan_cache():
    n, L, B = 4096, 2, 5
    rp = n // 128
    primes, psis = synth(n, L)
    prog = asm.transform_stream(n, primes).words()
    rng = np.random.default_rng(9)
    x = np.stack([rng.integers(0, primes[l % L], n, dtype=np.uint64) for l in range(B * L)])
    outs = []
    for batched in (False, True):
        eng = A.Engine(vlmax_bits=n * 64, spm_rows=2 * B * L * rp, ksk_rows=0, moduli=list(zip(primes, psis)))
        eng.load_isram(prog, 0)
        eng.dma_mem_h2d(0, x.reshape(-1))
        calls = [(b * L * rp, 0, (B + b) * L * rp, 0, 0) for b in range(B)]
        for _ in range(4):   # entry state differs between the first two runs (v0/v2 start undefined)
            if batched:
                eng.run_vp_batch(0, calls)
            else:
                for c in calls:
                    eng.run_vp(0, *c)
        outs.append(eng.dma_mem_d2h(B * L * rp, B * L * n))
        st = eng.stats()
        assert st["plans_reused"] >= 2, st
        if batched:   # all B*L transforms of one batch share the two launches of one forward NTT
            assert st["kernel_launches"] == 4 * 2, st
            assert st["copies_emitted"] == 0
    assert (outs[0] == outs[1]).all()


def test_register_alias_survives_memory_overwrite_and_undefined_reads_fail():
    n = 256
    q = O.Q0
    eng = A.Engine(vlmax_bits=n * 64, spm_rows=16, ksk_rows=0, moduli=())
    load = asm.Program().vsetvl(n).vsetq(q).vle(0, 0, 0).brk()
    store = asm.Program().vse(0, 2, 0).brk()
    use_undef = asm.Program().vcpy(2, 5).brk()
    eng.load_isram(load.words(), 0)
    eng.load_isram(store.words(), 100)
    eng.load_isram(use_undef.words(), 200)
    a = np.arange(n, dtype=np.uint64)
    eng.dma_mem_h2d(0, a)
    eng.run_vp(0, 0, 0, 0)                 # v0 aliases SPM rows 0-1
    eng.dma_mem_h2d(0, a + np.uint64(1000))    # overwrite the aliased rows: v0 must keep the OLD data
    eng.run_vp(100, 0, 0, 4)
    assert (eng.dma_mem_d2h(4, n) == a).all()
    assert (eng.dma_mem_d2h(0, n) == a + np.uint64(1000)).all()
    with pytest.raises(A.AlohaError) as e:
        eng.run_vp(200)
    assert e.value.name == "E_UNDEFINED"


def test_illegal_streams_error_codes():
    n = 256
    eng = A.Engine(vlmax_bits=n * 64, spm_rows=16, ksk_rows=0, moduli=((O.Q0, pow(O.PSI0, 8192 // n, O.Q0)),))
    eng.dma_mem_h2d(0, np.zeros(n, dtype=np.uint64))
    cases = {
        "E_ILLEGAL": asm.Program().vsetvl(n).vsetq(O.Q0).vle(2, 0, 0).vntt(2, 2).brk(),
        "E_STATE": asm.Program().vsetvl(n).vsetq(O.Q1).vle(0, 0, 0).vntt(2, 0).brk(),   # no ROM for q1
        "E_RANGE": asm.Program().vsetvl(n).vle(0, 0, 100).brk(),
        "E_NOBREAK": asm.Program().vsetvl(n),
    }
    for name, prog in cases.items():
        eng.load_isram(np.zeros((4096, 12), dtype=np.uint8), 0)   # NOPs
        eng.load_isram(prog.words(), 0)
        with pytest.raises(A.AlohaError) as e:
            eng.run_vp(0)
        assert e.value.name == name, name


# ------------------------------------------------------------------ (3) properties at full size
def test_full_size_roundtrip_and_linearity():
    """N = 2^16, 32 limbs, 4 polynomials: INTT(NTT(x)) == x, NTT(x + y) == NTT(x) + NTT(y)."""
    n, L, B = 65536, 32, 4
    rp = n // 128
    primes, psis = synth(n, L)
    rows = B * L * rp
    eng = A.Engine(vlmax_bits=n * 64, spm_rows=4 * rows, ksk_rows=0, moduli=list(zip(primes, psis)))
    eng.load_isram(asm.transform_stream(n, primes).words(), 0)
    eng.load_isram(asm.transform_stream(n, primes, inverse=True).words(), 1024)
    eng.load_isram(asm.elementwise_stream(n, primes, "vfqadd").words(), 2048)
    rng = np.random.default_rng(1)
    qv = np.array(primes, dtype=np.uint64)[None, :, None]
    x = rng.integers(0, 1 << 59, (B, L, n), dtype=np.uint64) % qv
    eng.dma_mem_h2d(0, x.reshape(-1))
    calls = lambda src, dst, src1=0: [(src + b * L * rp, src1 + b * L * rp, dst + b * L * rp, 0, 0) for b in range(B)]
    eng.run_vp_batch(0, calls(0, rows))                 # F = NTT(x)         at rows
    eng.run_vp_batch(1024, calls(rows, 2 * rows))       # INTT(F)            at 2*rows
    back = eng.dma_mem_d2h(2 * rows, B * L * n).reshape(B, L, n)
    assert (back == x).all()
    # linearity: polys 0,1 -> NTT(x0 + x1) vs NTT(x0) + NTT(x1)
    F = eng.dma_mem_d2h(rows, B * L * n).reshape(B, L, n)
    eng.run_vp(2048, 0, L * rp, 3 * rows)               # x0 + x1 at 3*rows
    eng.run_vp(0, 3 * rows, 0, 3 * rows + L * rp)       # NTT of that
    lhs = eng.dma_mem_d2h(3 * rows + L * rp, L * n).reshape(L, n)
    rhs = (F[0] + F[1]) % qv[0]
    assert (lhs == rhs).all()
    # spot-check two limbs of one polynomial against the oracle
    tabs = O.NttTables(n, primes, psis)
    sel = np.array([0, 31])
    assert (F[2, sel] == tabs.batch(x[2, sel].copy(), sel)).all()


def test_bench_workload_single_launch_parity():
    """The bench shape (64 polys x 32 limbs, N = 2^16) with the whole batch in ONE launch pair:
    eight randomly chosen limb-polys against the oracle, the rest by the inverse round trip."""
    n, L, B = 65536, 32, 64
    rp = n // 128
    primes, psis = synth(n, L)
    rows = B * L * rp
    eng = A.Engine(vlmax_bits=n * 64, spm_rows=2 * rows, ksk_rows=0, moduli=list(zip(primes, psis)),
                   l2_chunk_bytes=1 << 40)
    eng.load_isram(asm.transform_stream(n, primes).words(), 0)
    eng.load_isram(asm.transform_stream(n, primes, inverse=True).words(), 1024)
    rng = np.random.default_rng(77)
    qv = np.array(primes, dtype=np.uint64)[None, :, None]
    x = rng.integers(0, 1 << 59, (B, L, n), dtype=np.uint64) % qv
    eng.dma_mem_h2d(0, x.reshape(-1))
    s0 = eng.stats()
    eng.run_vp_batch(0, [(b * L * rp, 0, rows + b * L * rp, 0, 0) for b in range(B)])
    assert eng.stats()["kernel_launches"] - s0["kernel_launches"] == 2
    F = eng.dma_mem_d2h(rows, B * L * n).reshape(B, L, n)
    tabs = O.NttTables(n, primes, psis)
    for b, l in [(0, 0), (63, 31), (17, 5), (40, 20), (1, 30), (62, 1), (33, 16), (8, 8)]:
        assert (F[b, l] == tabs.batch(x[b, l][None].copy(), np.array([l]))[0]).all(), (b, l)
    eng.run_vp_batch(1024, [(rows + b * L * rp, 0, b * L * rp, 0, 0) for b in range(B)])
    back = eng.dma_mem_d2h(0, B * L * n).reshape(B, L, n)
    assert (back == x).all()


# ------------------------------------------------------------------ limb-sharded key-switch stream
def _ks_engine_factory(lay, moduli_psi):
    return A.Engine(vlmax_bits=lay.n * 64, spm_rows=lay.spm_rows, ksk_rows=lay.ksk_rows, moduli=moduli_psi,
                    pool_buffers=256, isram_depth=16384)


def test_generalised_keyswitch_stream_reference_vectors_on_gpu():
    import test_keyswitch_sharded as T
    for item in [i for i in G.manifest()["kernels"] if i["op"] == "rotate"]:
        got, want = T.run_reference_rotate_vector(item, _ks_engine_factory)
        assert got == want, (item["case"], item["kernel"])


@pytest.mark.parametrize("n,L", [(1024, 5), (65536, 3)])
def test_generalised_keyswitch_stream_vs_oracle(n, L):
    import test_keyswitch_sharded as T
    from aloha_b200 import keyswitch as KS
    gpu = T.run_sharded(n, L, 1, 0, KS.LocalComm(), _ks_engine_factory)
    cpu = T.run_sharded(n, L, 1, 0, KS.LocalComm())
    for i in range(L):
        assert (gpu[i][0] == cpu[i][0]).all() and (gpu[i][1] == cpu[i][1]).all(), i


def test_async_dma_pipeline_equals_blocking_dma():
    """Chunked upload / transform / download through the asynchronous DMA channels gives the same
    bytes as the blocking calls, including when the same rows are reused right away."""
    import torch
    n, L, B = 4096, 2, 8
    rp = n // 128
    primes, psis = synth(n, L)
    per_poly = L * rp
    eng = A.Engine(vlmax_bits=n * 64, spm_rows=2 * B * per_poly, ksk_rows=0, moduli=list(zip(primes, psis)))
    eng.load_isram(asm.transform_stream(n, primes).words(), 0)
    rng = np.random.default_rng(4)
    pin = (lambda t: t.pin_memory()) if torch.cuda.is_available() else (lambda t: t)     # (no driver on the simulated device)
    hin = pin(torch.empty(B * L * n, dtype=torch.int64))
    hout = pin(torch.empty(B * L * n, dtype=torch.int64))
    for rep in range(3):                      # reuse the same SPM rows and host buffers three times
        x = np.stack([rng.integers(0, primes[i % L], n, dtype=np.uint64) for i in range(B * L)])
        hin.numpy().view(np.uint64)[:] = x.reshape(-1)
        hout.zero_()
        cb = 2 * L * n * 8                    # 2 polys per chunk
        for c in range(B // 2):
            eng.dma_mem_h2d_async(2 * c * per_poly, hin.data_ptr() + c * cb, cb)
            eng.run_vp_batch(0, [((2 * c + b) * per_poly, 0, (B + 2 * c + b) * per_poly, 0, 0) for b in range(2)])
            eng.dma_mem_d2h_async(hout.data_ptr() + c * cb, (B + 2 * c) * per_poly, cb)
        eng.sync()
        want = O.NttTables(n, primes, psis).batch(x.copy(), np.arange(B * L) % L)
        assert (hout.numpy().view(np.uint64).reshape(B * L, n) == want).all(), rep
        assert (eng.dma_mem_d2h(B * per_poly, B * L * n).reshape(B * L, n) == want).all()


def test_replay_cli_writes_reference_format_dumps(tmp_path):
    """python -m aloha_b200.replay on case0_4_4 rebuilt as text files: the dump files it writes parse
    back to the reference's golden hashes ('x' lines included)."""
    from aloha_b200 import replay
    m = G.manifest()
    n, entry = m["n"], m["cases"]["case0_4_4"]
    ops, dram, enc, ksk = G.case_inputs("case0_4_4")

    def write(path, arr):
        with open(path, "w") as f:
            f.write("\n".join(str(int(v)) for v in arr) + "\n")
    (tmp_path / "prog.txt").write_text("\n".join(entry["program"]) + "\n")
    write(tmp_path / "ksk2.txt", ksk[0])
    args = ["--program", str(tmp_path / "prog.txt"), "--isram", G.write_microcode_dir(str(tmp_path / "isram")),
            "--ksk", f"2:{tmp_path / 'ksk2.txt'}", "--dump-dir", str(tmp_path / "out"), "--cipher"]
    for i, key in entry["loads"].items():
        write(tmp_path / f"ct{i}.txt", G.pool(key))
        args.append(f"{ops[int(i)].dram_addr}:{tmp_path / f'ct{i}.txt'}")
    args.append("--encoder")
    for i, data in enc.items():
        write(tmp_path / f"enc{i}.txt", data)
        args.append(f"{i}:{tmp_path / f'enc{i}.txt'}")
    assert replay.main(args) == 0
    for name, want in entry["dumps"].items():
        toks = (tmp_path / "out" / (name + ".txt")).read_text().split()
        assert len(toks) == 4 * n
        wr = np.array([t != "x" for t in toks])
        data = np.array([int(t) if t != "x" else 0 for t in toks], dtype=np.uint64)
        assert G.poly_hashes(data, wr, n) == want, name


def test_edge_cases_empty_batch_mixed_vl_and_last_rows():
    """count = 0 batches, a stream that changes VL midway (ragged polynomial sizes in one plan), and
    accesses that end exactly at the last SPM row."""
    primes, psis = synth(1024, 1)
    q = primes[0]
    rows = 32
    eng = A.Engine(vlmax_bits=1024 * 64, spm_rows=rows, ksk_rows=0, moduli=list(zip(primes, psis)))
    eng.run_vp_batch(0, [])                                   # nothing to do, nothing to fail
    assert eng.stats()["instructions"] == 0
    p = asm.Program().vsetvl(256).vsetq(q).vle(0, 0, 0).vntt(2, 0).vse(2, 2, 0)
    p.vsetvl(1024).vle(1, 0, 8).vntt(3, 1).vse(3, 2, 8).brk()     # second half at N = 1024
    eng.load_isram(p.words(), 0)
    rng = np.random.default_rng(8)
    a, b = rng.integers(0, q, 256, dtype=np.uint64), rng.integers(0, q, 1024, dtype=np.uint64)
    eng.dma_mem_h2d(0, a)
    eng.dma_mem_h2d(8, b)
    eng.run_vp(0, 0, 0, 16)                                   # outputs at rows 16..17 and 24..31 (the last row)
    psi256, psi1024 = pow(psis[0], 4, q), psis[0]
    assert (eng.dma_mem_d2h(16, 256) == O.ntt(a, q, psi256)).all()
    assert (eng.dma_mem_d2h(24, 1024) == O.ntt(b, q, psi1024)).all()
    assert eng.spm_written(16, 256).all() and not eng.spm_written(18, 6 * 128).any()
    with pytest.raises(A.AlohaError) as e:                    # one row further is out of range
        eng.run_vp(0, 0, 0, 17)
    assert e.value.name == "E_RANGE"
    with pytest.raises(A.AlohaError):
        eng.dma_mem_h2d(rows - 1, np.zeros(256, dtype=np.uint64))


@pytest.mark.parametrize("n", [1024, 8192])
def test_strict_mode_matches_rtl_on_garbage_inputs_and_clobbered_source(n):
    """ALOHA_F_STRICT: raw 64-bit input words (far above 2q), the ping-pong intermediate the RTL leaves
    in the source register, and the ROM fallback for an unknown modulus (Q6) -- all three
    are outside the fast path's contract and must equal the oracle word for word here."""
    rp = n // 128
    q, psi = O.Q0, pow(O.PSI0, 8192 // n, O.Q0)
    odd_q = O.Q1                                   # no ROM provisioned for it: falls back to the last table
    rng = np.random.default_rng(n)
    raw = rng.integers(0, 2**64, n, dtype=np.uint64)
    prog = asm.Program().vsetvl(n).vsetq(q).vle(0, 0, 0).vntt(2, 0).vse(2, 2, 0).vse(0, 2, rp)   # result, clobbered source
    prog.vle(4, 0, 0).vintt(6, 4).vse(6, 2, 2 * rp).vse(4, 2, 3 * rp)
    prog.vsetq(odd_q).vle(1, 0, 0).vntt(3, 1).vse(3, 2, 4 * rp).brk()
    machines = [A.Engine(vlmax_bits=n * 64, spm_rows=8 * rp, ksk_rows=0, moduli=[(q, psi)], flags=A.F_STRICT),
                O.GoldenModel(vlmax_bits=n * 64, spm_rows=8 * rp, ksk_rows=0, moduli=[(q, psi)])]
    outs = []
    for m in machines:
        m.load_isram(prog.words(), 0)
        m.dma_mem_h2d(0, raw)
        m.run_vp(0, 0, 0, rp)
        outs.append(m.dma_mem_d2h(rp, 5 * n))
    assert (outs[0] == outs[1]).all(), [int((outs[0][i * n:(i + 1) * n] != outs[1][i * n:(i + 1) * n]).sum()) for i in range(5)]
    # the fast path refuses the same stream instead of guessing
    eng = A.Engine(vlmax_bits=n * 64, spm_rows=8 * rp, ksk_rows=0, moduli=[(q, psi)])
    eng.load_isram(prog.words(), 0)
    eng.dma_mem_h2d(0, raw)
    with pytest.raises(A.AlohaError) as e:
        eng.run_vp(0, 0, 0, rp)
    assert e.value.name == "E_UNDEFINED"


def test_replay_cli_from_readmemh_dram_image(tmp_path):
    """case0_4_4 again, this time from a single $readmemh DDR image (ciphertext at DRAM_VP_BASE, the
    rotation key in its KSK slot) as the reference's DRAM_INPUT_FILE would provide it."""
    from aloha_b200 import dram_image as D, replay
    m = G.manifest()
    n, entry = m["n"], m["cases"]["case0_4_4"]
    ops, dram, enc, ksk = G.case_inputs("case0_4_4")
    img = D.build_image(64 << 20, ciphertexts={ops[int(i)].dram_addr: G.pool(k) for i, k in entry["loads"].items()},
                        ksks={2: ksk[0]})
    D.write_readmemh(str(tmp_path / "dram.mem"), img[:(D.DRAM_VP_BASE + 4 * n * 8) // 8])
    (tmp_path / "prog.txt").write_text("\n".join(entry["program"]) + "\n")
    args = ["--program", str(tmp_path / "prog.txt"), "--isram", G.write_microcode_dir(str(tmp_path / "isram")),
            "--dram-image", str(tmp_path / "dram.mem"), "--dump-dir", str(tmp_path / "out"), "--encoder"]
    for i, data in enc.items():
        with open(tmp_path / f"enc{i}.txt", "w") as f:
            f.write("\n".join(str(int(v)) for v in data) + "\n")
        args.append(f"{i}:{tmp_path / f'enc{i}.txt'}")
    assert replay.main(args) == 0
    toks = (tmp_path / "out" / "inst_7_out.txt").read_text().split()
    data = np.array([int(t) for t in toks], dtype=np.uint64)
    assert G.poly_hashes(data, np.ones(len(data), bool), n) == entry["dumps"]["inst_7_out"]


def test_vfqsub_sv_follows_the_rtl_operand():
    """VFQSUB.sv computes imm - vs2 with the vector operand taken from vs2 (expander.v:342-363), not vs1 as
    the reference's own decode golden expects -- GPU, oracle and exact arithmetic agree."""
    n, q = 256, O.Q0
    imm = 0x123456789
    eng = A.Engine(vlmax_bits=n * 64, spm_rows=8, ksk_rows=0, moduli=())
    p = asm.Program().vsetvl(n).vsetq(q).vle(4, 0, 0).vle(7, 0, 2).vfqsub_sv(2, imm, 4).vse(2, 2, 0).brk()
    rng = np.random.default_rng(1)
    a, b = rng.integers(0, q, n, dtype=np.uint64), rng.integers(0, q, n, dtype=np.uint64)
    gm = O.GoldenModel(vlmax_bits=n * 64, spm_rows=8, ksk_rows=0, moduli=())
    outs = []
    for mach in (eng, gm):
        mach.load_isram(p.words(), 0)
        mach.dma_mem_h2d(0, np.concatenate([a, b]))
        mach.run_vp(0, 0, 0, 4)
        outs.append(mach.dma_mem_d2h(4, n))
    want = np.array([(imm - int(v)) % q for v in a], dtype=np.uint64)
    assert (outs[0] == want).all() and (outs[1] == want).all()


def test_deferred_queue_batches_across_calls_and_keeps_error_semantics():
    """ALOHA_F_DEFER: sixteen separate run_vp calls become one plan (two launches) at the next DMA; a
    malformed call in the queue still lets the calls before it take effect."""
    n, L, B = 4096, 2, 8
    rp = n // 128
    primes, psis = synth(n, L)
    per_poly = L * rp
    rng = np.random.default_rng(12)
    x = np.stack([rng.integers(0, primes[i % L], n, dtype=np.uint64) for i in range(B * L)])
    want = O.NttTables(n, primes, psis).batch(x.copy(), np.arange(B * L) % L)
    eng = A.Engine(vlmax_bits=n * 64, spm_rows=2 * B * per_poly, ksk_rows=0, moduli=list(zip(primes, psis)), flags=A.F_DEFER)
    eng.load_isram(asm.transform_stream(n, primes).words(), 0)
    eng.load_isram(asm.Program().vsetvl(n).vle(0, 0, 60000).brk().words(), 512)     # E_RANGE when it runs
    eng.dma_mem_h2d(0, x.reshape(-1))
    for b in range(B):
        eng.run_vp(0, b * per_poly, 0, (B + b) * per_poly)                           # queued, not launched
    got = eng.dma_mem_d2h(B * per_poly, B * L * n).reshape(B * L, n)                 # flush point
    assert (got == want).all()
    st = eng.stats()
    assert st["kernel_launches"] == 2 and st["plans_built"] == 1
    # error semantics
    eng.dma_mem_h2d(B * per_poly, np.zeros(B * L * n, dtype=np.uint64))
    eng.run_vp(0, 0, 0, B * per_poly)            # fine
    eng.run_vp(512)                              # malformed; not noticed yet
    eng.run_vp(0, per_poly, 0, (B + 1) * per_poly)
    with pytest.raises(A.AlohaError) as e:
        eng.sync()
    assert e.value.name == "E_RANGE"
    out = eng.dma_mem_d2h(B * per_poly, 2 * L * n).reshape(2 * L, n)
    assert (out[:L] == want[:L]).all()           # the call before the offender ran
    assert not out[L:].any()                     # the one after it did not
