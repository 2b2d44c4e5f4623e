"""GPU parity of the hybrid key-switch / relinearise / rescale streams (aloha_b200.hks) through the C-ABI:
the CUDA engine against the oracle on the same streams and inputs -- small shapes on every output word, the
BASELINE.json config-5 shapes (N = 2^16; 47 + 1 limbs with one-limb digits, and 40 + 8 limbs with dnum = 5) on
four output limbs -- plus the NCCL group path: the C replay tool (one process, n engines, no Python on the
data path) and the Python face of aloha_group_*."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

import aloha_b200 as A
from aloha_b200 import asm, hks, params
from oracle import oracle as O
import test_hks as T

pytestmark = pytest.mark.gpu


SIMULATED = os.environ.get("ALOHA_TEST_DEVICE") == "sim"       # tests/conftest.py: the engine's host code on a simulated device


def engine_for(lay, prm, psi, device=0, **kw):
    return A.Engine(vlmax_bits=prm.n * 64, spm_rows=lay.spm_rows, ksk_rows=max(lay.ksk_rows, 1), device=device,
                    moduli=[(m, psi[m]) for m in prm.moduli], pool_buffers=kw.pop("pool_buffers", 512),
                    isram_depth=kw.pop("isram_depth", 65536), **kw)


def fill(ks, lay, prm, ct, ksk, only_moduli=None):
    for b in range(lay.batch):
        for i in lay.owned():
            if i < prm.L:
                ks.load_input(i, [np.roll(ct[c][i], b) for c in range(len(ct))], b)
    for t in lay.owned():
        if only_moduli is None or t in only_moduli:
            ks.load_ksk(t, np.stack([ksk[t][b][c] for b in range(prm.dnum) for c in (0, 1)]))


@pytest.mark.parametrize("n,L,K,dnum,kind,batch", [(1024, 6, 2, 3, "rotate", 1), (1024, 6, 2, 3, "relin", 2),
                                                   (4096, 4, 1, 4, "rotate", 1), (256, 5, 3, 2, "rotate", 3)])
def test_engine_equals_oracle_on_generated_streams(n, L, K, dnum, kind, batch):
    prm, psi, ct, ksk = T.make_problem(n, L, K, dnum, kind)
    k = pow(3, 7, 2 * n)
    want = T.run_machine(prm, psi, ct, ksk, k, kind, batch=batch)
    lay = hks.Layout(prm, 1, 0, batch, kind)
    eng = engine_for(lay, prm, psi)
    ks = hks.KeySwitch(eng, lay)
    fill(ks, lay, prm, ct, ksk)
    for _ in range(2):                         # second pass: cached plans
        ks.run(k if kind == "rotate" else 1)
    for (b, i), (x, y) in want.items():
        gx, gy = ks.read_output(i, b)
        assert (gx == x).all() and (gy == y).all(), (b, i)
    assert eng.stats()["ops_fused"] > 0


@pytest.mark.parametrize("L,K,dnum", [(47, 1, 47), (40, 8, 5)])
def test_config5_shapes_four_output_limbs_vs_oracle(L, K, dnum):
    """N = 2^16 at the bench's two shapes.  Phases 2 and 3 run for four output limbs (+ the special primes) on
    both machines: every limb's stream is independent of the other limbs' streams, so this is the full
    computation of those limbs."""
    n = 65536
    only = [0, 13, 29, L - 1]
    primes = params.synthetic_primes(L + K, 2 * n)
    p, q = primes[:K], primes[K:]
    psi = {m: params.min_primitive_root(m, 2 * n) for m in primes}
    prm = hks.Params(n, q, p, dnum)
    rng = np.random.default_rng(48)
    ct = [[rng.integers(0, qi, n, dtype=np.uint64) for qi in q] for _ in range(2)]
    need = set(only) | set(range(L, L + K))
    ksk = {t: [[rng.integers(0, prm.moduli[t], n, dtype=np.uint64) for _ in (0, 1)] for _ in range(prm.dnum)] for t in need}
    k = pow(3, 18, 2 * n) & (n - 1)
    lay = hks.Layout(prm)
    outs = []
    for make in (lambda: T.oracle_machine(lay, n, [(m, psi[m]) for m in prm.moduli]), lambda: engine_for(lay, prm, psi, pool_buffers=1024)):
        m = make()
        ks = hks.KeySwitch(m, lay)
        fill(ks, lay, prm, ct, ksk, only_moduli=need)
        ks.run(k, only=only)
        outs.append({i: ks.read_output(i) for i in only})
        del ks, m
    for i in only:
        assert (outs[0][i][0] == outs[1][i][0]).all() and (outs[0][i][1] == outs[1][i][1]).all(), i


@pytest.mark.parametrize("bad_iq", [False, True])
def test_base_extension_chain_on_raw_words(bad_iq):
    """The fused base-extension kernel against the oracle on words the generated streams never produce: raw
    64-bit garbage behind VCPY / no pre-op (its lazy fast path must hand those elements to the RTL chain),
    and a VSETIQ that is not q's Barrett constant (the whole job takes the RTL chain).  Nine terms, so the
    mid-chain reduction of the lazy sums runs too."""
    n, rp = 1024, 8
    q = O.Q0
    rng = np.random.default_rng(77)
    nterms = 9
    x = rng.integers(0, q, (nterms, n), dtype=np.uint64)
    x[0] = rng.integers(0, 2**64, n, dtype=np.uint64)          # behind VFQMOD: any word is in the domain
    x[1, ::7] = rng.integers(0, 2**64, len(x[1, ::7]), dtype=np.uint64)   # behind VCPY: some words >= 4q
    x[2, ::5] += np.uint64(3 * q)                               # behind VCPY: [3q, 4q) is still in the domain
    x[3, 3::11] = rng.integers(2 * q, 2**64, len(x[3, 3::11]), dtype=np.uint64)  # no pre-op: some words >= 2q
    scal = [int(v) for v in rng.integers(1, q, nterms)]
    pre = ["vfqmod", "vcpy", "vcpy", None] + ["vcpy", "vfqmod"] * 3
    p = asm.Program().vsetvl(n).vsetq(q)
    if bad_iq:
        p.buf.append(asm.word(asm.F6["VSETIQ"], imm=asm.barrett_iq(q) ^ 0x5555))
    for t in range(nterms):
        p.vle(0, 0, t * rp)
        src = 0
        if pre[t]:
            getattr(p, pre[t])(8, 0)
            src = 8
        if t == 0:
            p.vfqmul(13, src, imm=scal[t])
        else:
            p.vfqmul(10, src, imm=scal[t]).vfqadd(13, 13, 10)
    p.vfqsub(2, 13, imm=12345).vse(2, 2, 0)
    # a second, throw-away chain redefines v8 / v10 / v13 so that the first one's temporaries are dead
    p.vle(0, 0, 0).vcpy(8, 0).vfqmul(13, 8, imm=3).vfqmul(10, 8, imm=5).vfqadd(13, 13, 10).vse(13, 2, rp)
    p.brk()
    outs = []
    for m in (O.GoldenModel(vlmax_bits=n * 64, spm_rows=(nterms + 2) * rp, ksk_rows=1, moduli=()),
              A.Engine(vlmax_bits=n * 64, spm_rows=(nterms + 2) * rp, ksk_rows=1, moduli=())):
        m.load_isram(p.words(), 0)
        m.dma_mem_h2d(0, x.reshape(-1))
        m.run_vp(0, 0, 0, nterms * rp)
        outs.append(m.dma_mem_d2h(nterms * rp, 2 * n))
        if isinstance(m, A.Engine):
            assert m.stats()["ops_fused"] >= 2 * nterms
    assert (outs[0] == outs[1]).all()


def test_rescale_engine_vs_oracle():
    n, L = 4096, 5
    q, _, psi, rng = T.synth(n, L, 0)
    ct = [[rng.integers(0, qi, n, dtype=np.uint64) for qi in q] for _ in range(2)]
    outs = []
    for machine in (T.oracle_machine(8 * L * n // 128, n, [(m, psi[m]) for m in q]),
                    A.Engine(vlmax_bits=n * 64, spm_rows=8 * L * n // 128, ksk_rows=1, moduli=[(m, psi[m]) for m in q])):
        rs = hks.Rescale(machine, n, q)
        for i in range(L):
            rs.load_input(i, ct[0][i], ct[1][i])
        rs.run()
        outs.append([rs.read_output(i) for i in range(L - 1)])
    for a, b in zip(*outs):
        assert (a[0] == b[0]).all() and (a[1] == b[1]).all()


def replay_case(world, n, L, K, dnum, overlap):
    """Write the case directory for `world` ranks, run the C tool, return {limb: (x, y)} it dumped."""
    prm, psi, ct, ksk = T.make_problem(n, L, K, dnum, "rotate")
    k = pow(3, 3, 2 * n)
    switches, loads, dumps = [], [], []
    for r in range(world):
        lay = hks.Layout(prm, world, r)
        rec = hks.Recorder()
        ks = hks.KeySwitch(rec, lay, type("C", (), {"world": world, "rank": r})(), overlap=overlap, lockstep=True)
        fill(ks, lay, prm, ct, ksk)
        switches.append(ks)
        loads.append(rec.loads)
        dumps.append([(lay.OUT + (c * L + i) * prm.rp, n, f"out_{c}_{i}.u64") for i in lay.owned() if i < L for c in (0, 1)])
    d = tempfile.mkdtemp(prefix="aloha_replay_")
    cfg = {"vlmax_bits": n * 64, "moduli": [(m, psi[m]) for m in prm.moduli], "pool_buffers": 256, "isram_depth": 65536}
    hks.write_replay_case(d, switches, cfg, loads, dumps, k)
    tool, env = os.path.join(os.path.dirname(A.__file__), "aloha_group_replay"), None
    if SIMULATED:                    # the same C program linked against the engine's host code on the simulated device
        import sim_engine
        tool, env = sim_engine.REPLAY, dict(os.environ, ALOHA_NCCL_LIB=sim_engine.NCCL, ALOHA_SIM_DEVICES="8")
    out = subprocess.run([tool, d, str(world)], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr + out.stdout
    got = {i: tuple(np.fromfile(os.path.join(d, f"out_{c}_{i}.u64"), dtype=np.uint64) for c in (0, 1)) for i in range(L)}
    want = T.run_machine(prm, psi, ct, ksk, k, "rotate")
    return got, want, out.stdout


@pytest.mark.parametrize("overlap", [False, "chunks"])
def test_c_replay_tool_one_rank(overlap):
    """the C host program + aloha_group_* with a group of one (NCCL initialised, collectives degenerate)"""
    L = 6
    got, want, log = replay_case(1, 1024, L, 2, 3, overlap)
    for i in range(L):
        assert (got[i][0] == want[0, i][0]).all() and (got[i][1] == want[0, i][1]).all(), i
    assert "kernel launches" in log


def _ngpus():
    if SIMULATED:
        return int(os.environ.get("ALOHA_SIM_DEVICES", "1"))
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("overlap", [False, "chunks", "own"])
def test_c_replay_tool_two_ranks_nccl(overlap):
    if _ngpus() < 2:
        pytest.skip("needs two GPUs")
    L = 6
    got, want, _ = replay_case(2, 1024, L, 2, 3, overlap)
    for i in range(L):
        assert (got[i][0] == want[0, i][0]).all() and (got[i][1] == want[0, i][1]).all(), i


def local_group_case(world, n, L, K, dnum, overlap, batch=2):
    """aloha_group_create_local from Python: `world` engines in one process, the engine's own (non-torch) streams,
    against the one-machine oracle run; twice, the second time on cached plans."""
    prm, psi, ct, ksk = T.make_problem(n, L, K, dnum, "rotate")
    k = pow(3, 9, 2 * n)
    want = T.run_machine(prm, psi, ct, ksk, k, "rotate", batch=batch)
    engines, switches = [], []
    for r in range(world):
        lay = hks.Layout(prm, world, r, batch=batch)
        engines.append(engine_for(lay, prm, psi, device=r))
    grp = A.Group.local(engines)
    for r in range(world):
        lay = hks.Layout(prm, world, r, batch=batch)
        ks = hks.KeySwitch(engines[r], lay, type("C", (), {"world": world, "rank": r})(), overlap=overlap, lockstep=True)
        fill(ks, lay, prm, ct, ksk)
        switches.append(ks)
    for _ in range(2):
        hks.run_local_group(switches, grp, k)
    seen = set()
    for r, ks in enumerate(switches):
        for b in range(batch):
            for i in ks.lay.owned():
                if i < L:
                    gx, gy = ks.read_output(i, b)
                    assert (gx == want[b, i][0]).all() and (gy == want[b, i][1]).all(), (r, b, i)
                    seen.add((b, i))
    assert len(seen) == batch * L
    grp.close()
    for e in engines:
        e.close()


@pytest.mark.parametrize("overlap", [False, "chunks", "own"])
def test_python_group_two_engines_nccl(overlap):
    """chunked all-gather with per-source waits, and the two other overlap modes, on two GPUs of one process"""
    if _ngpus() < 2:
        pytest.skip("needs two GPUs")
    local_group_case(2, 2048, 6, 2, 3, overlap)
