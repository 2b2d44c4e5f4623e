"""The batcher without a GPU: the engine's host sources built against a simulated device (tests/sim_engine.py,
tests/native/sim/) and driven by the GPU suite's own test bodies.

  * every `-m gpu` test that does not need NCCL or the C replay binary is run here, in a subprocess with
    ALOHA_TEST_DEVICE=sim, against the oracle: the plans the batcher builds (store forwarding, copy-on-write,
    fusion, levelling, plan cache, deferred queue, graphs, asynchronous DMA, the C host driver) compute what the
    instruction streams say, and every launch satisfies the checks in sim_kernels.cpp (operands inside device
    memory, no destination overlapping a permuted operand, no job reading what another job of the same launch
    writes, row groups / tensor-map coordinates / tile plans / Shoup companions consistent with the job records),
    and no two operations on different streams touch the same device bytes without an event or a synchronisation
    between them (vector clocks in sim_cuda.cpp: the asynchronous DMA channels against the engine's stream);
  * the simulated kernels do refuse what the real ones cannot take (so the first point is not vacuous);
  * more fuzzing seeds than the GPU budget allows.
None of this says anything about the sm_100a kernels: those are only checked on a B200."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import sim_engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gpu_suite_on_the_simulated_device():
    env = dict(os.environ, ALOHA_TEST_DEVICE="sim")
    sim_engine.build()
    workers = min(4, os.cpu_count() or 1)
    cmd = [sys.executable, "-m", "pytest", os.path.join(ROOT, "tests"), "-q", "-m", "gpu", "-p", "no:cacheprovider"]
    if workers > 1:
        cmd += ["-n", str(workers)]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=1500, cwd=ROOT)
    tail = out.stdout[-3000:] + out.stderr[-1000:]
    assert out.returncode == 0, tail
    summary = [l for l in out.stdout.splitlines() if " passed" in l][-1]
    passed = int(summary.split(" passed")[0].split()[-1])
    assert passed >= 150 and "failed" not in summary, summary


@pytest.mark.parametrize("seed", range(1000, 1040))
def test_more_fuzzing_than_the_gpu_budget_allows(seed):
    import test_gpu_fuzz as F
    with sim_engine.simulated() as A:
        if seed % 3 == 2:
            F.run_case(seed + 200000, strict=True, flags=A.F_STRICT | (A.F_DEFER if seed % 2 else 0))
        else:
            F.run_case(seed, strict=False, flags=[0, A.F_DEFER, A.F_NO_FUSE, A.F_GRAPHS][seed % 4])


def test_simulated_kernels_refuse_what_the_real_ones_cannot_take():
    """hand-made job tables and stream programs straight into the simulated runtime: an automorphism in place, a job
    that reads another job's output, an operand outside device memory, a kernel reading what another stream wrote
    with nothing ordering the two -- each must fail; the legal twins of those must not"""
    code = "import ctypes as C, sys\nfn = getattr(C.CDLL(sys.argv[1]), sys.argv[2]); fn.restype = C.c_int; print(fn())"
    lib = sim_engine.build()
    got = {}
    for name in ("sim_test_legal", "sim_test_vaut_in_place", "sim_test_cross_job_read", "sim_test_foreign_pointer",
                 "sim_test_race", "sim_test_race_ordered_by_event", "sim_test_race_ordered_by_host"):
        # one process per case: a violation is sticky by design
        out = subprocess.run([sys.executable, "-c", code, lib, name], capture_output=True, text=True, timeout=120)
        assert out.returncode == 0, out.stderr[-2000:]
        got[name] = int(out.stdout.split()[-1])
    assert got["sim_test_legal"] == 0 and got["sim_test_race_ordered_by_event"] == 0 and got["sim_test_race_ordered_by_host"] == 0
    assert got["sim_test_vaut_in_place"] != 0 and got["sim_test_cross_job_read"] != 0 and got["sim_test_foreign_pointer"] != 0
    assert got["sim_test_race"] != 0


@pytest.mark.parametrize("overlap", [False, "chunks", "own"])
@pytest.mark.parametrize("world,L,K,dnum", [(8, 40, 8, 5), (8, 47, 1, 47), (4, 40, 8, 5), (3, 7, 2, 3)])
def test_limb_sharded_key_switch_on_simulated_devices(world, L, K, dnum, overlap):
    """group.cpp at the bench's two key-switch shapes over 8 and 4 devices (at 8, one rank owns nothing but special
    primes) and over a world size that divides nothing: every output word against the one-machine oracle run, with the
    simulated runtime checking that the engines' streams and the communication streams are ordered by events wherever
    they touch the same rows.  NCCL is tests/native/sim/sim_nccl.cpp (one process, copies between the devices)."""
    import test_gpu_hks as H
    with sim_engine.simulated() as A:
        H.A = A
        try:
            H.local_group_case(world, 256, L, K, dnum, overlap, batch=2)
        finally:
            import aloha_b200
            H.A = aloha_b200


@pytest.mark.parametrize("world", [4, 8])
def test_c_host_program_on_simulated_devices(world):
    """aloha_group_replay (the C program of INTEGRATION.md, which sees only include/aloha_b200.h) linked against the
    simulated library: config 5 without Python at 4 and 8 ranks"""
    import test_gpu_hks as H
    saved = H.SIMULATED
    H.SIMULATED = True
    try:
        L = 13
        got, want, log = H.replay_case(world, 256, L, 3, 5, "chunks")
        for i in range(L):
            assert (got[i][0] == want[0, i][0]).all() and (got[i][1] == want[0, i][1]).all(), i
        assert f"{world} rank(s)" in log
    finally:
        H.SIMULATED = saved


def rank_per_thread_case(A, world, n, L, K, dnum, overlap, batch, passes=3, flags=0):
    """What bench.py's key-switch leg does with one process per GPU, here with one THREAD per simulated device:
    aloha_group_create from a shared id, every rank walking its OWN op list (no lockstep), nobody waiting for a
    block it does not need.  -> {(batch element, limb): (out_0, out_1)} collected from the ranks that own the limbs."""
    import threading
    import test_gpu_hks as H
    import test_hks as T
    from aloha_b200 import hks
    prm, psi, ct, ksk = T.make_problem(n, L, K, dnum, "rotate")
    k = pow(3, 9, 2 * n)
    want = T.run_machine(prm, psi, ct, ksk, k, "rotate", batch=batch)
    uid = A.Group.unique_id()
    got, errors = {}, []

    def rank_main(r):
        try:
            lay = hks.Layout(prm, world, r, batch=batch)
            eng = A.Engine(vlmax_bits=n * 64, spm_rows=lay.spm_rows, ksk_rows=max(lay.ksk_rows, 1), device=r,
                           moduli=[(m, psi[m]) for m in prm.moduli], pool_buffers=512, isram_depth=65536, flags=flags)
            grp = A.Group.create(eng, uid, r, world)
            ks = hks.KeySwitch(eng, lay, hks.GroupComm(grp), overlap=overlap)
            H.fill(ks, lay, prm, ct, ksk)
            for _ in range(passes):
                ks.run(k)
            eng.sync()
            for b in range(batch):
                for i in lay.owned():
                    if i < L:
                        got[b, i] = ks.read_output(i, b)
            grp.close()
            eng.close()
        except BaseException as e:              # noqa: BLE001 -- reported by the caller
            errors.append((r, repr(e)))
    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not errors, errors
    assert not any(t.is_alive() for t in threads), "a rank is stuck in a collective"
    assert sorted(got) == sorted(want)
    for key, (x, y) in want.items():
        assert (got[key][0] == x).all() and (got[key][1] == y).all(), key


@pytest.mark.parametrize("overlap", [False, "chunks", "own"])
@pytest.mark.parametrize("world,L,K,dnum,batch", [(8, 40, 8, 5, 1), (8, 47, 1, 47, 1), (8, 40, 8, 5, 2), (4, 40, 8, 5, 1), (2, 40, 8, 5, 1), (5, 9, 2, 4, 2)])
def test_one_rank_per_thread_like_one_process_per_gpu(world, L, K, dnum, batch, overlap):
    """bench.py's multi-GPU key-switch path (aloha_group_create + each rank's own op list) at its two shapes over
    2 / 4 / 8 simulated devices, and an awkward world size, under the stream-race check: this is the code the 8-GPU
    run of the scaling bench executes, which no B200 box has run yet."""
    with sim_engine.simulated() as A:
        rank_per_thread_case(A, world, 256, L, K, dnum, overlap, batch)


@pytest.mark.parametrize("world,overlap", [(2, False), (3, "own"), (4, "chunks")])
def test_multiply_chain_one_rank_per_thread(world, overlap):
    """tensor product -> relinearise -> rescale, limb-sharded over simulated devices with aloha_group_* transfers"""
    import threading
    import test_hks_multiply as M
    from aloha_b200 import hks
    n, L, K, dnum = 256, 7, 2, 3
    prm, psi, a, b, ksk = M.problem(n, L, K, dnum)
    want = M.run_multiply(prm, psi, a, b, ksk)
    with sim_engine.simulated() as A:
        uid = A.Group.unique_id()
        got, errors = {}, []

        def rank_main(r):
            try:
                lay = hks.Layout(prm, world, r, 1, "relin")
                eng = A.Engine(vlmax_bits=n * 64, spm_rows=hks.Multiply.spm_rows(prm, world), ksk_rows=max(lay.ksk_rows, 1), device=r,
                               moduli=[(m, psi[m]) for m in prm.moduli], pool_buffers=256, isram_depth=65536)
                grp = A.Group.create(eng, uid, r, world)
                mul = hks.Multiply(eng, prm, world, r, hks.GroupComm(grp), overlap=overlap)
                for i in mul.lay.owned():
                    if i < L:
                        mul.load_input(i, (a[0][i], a[1][i]), (b[0][i], b[1][i]))
                for t in mul.lay.owned():
                    mul.load_ksk(t, np.stack([ksk[t][d][c] for d in range(prm.dnum) for c in (0, 1)]))
                for _ in range(2):
                    mul.run()
                eng.sync()
                for i in mul.lay.owned():
                    if i < L - 1:
                        got[i] = mul.read_output(i)
                grp.close()
                eng.close()
            except BaseException as e:          # noqa: BLE001
                errors.append((r, repr(e)))
        threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
        for t in threads:
            t.start()
        for t in threads:
            t.join(timeout=300)
        assert not errors, errors
        assert sorted(got) == sorted(want)
        for i, (x, y) in want.items():
            assert (got[i][0] == x).all() and (got[i][1] == y).all(), i


def test_random_sharded_shapes_one_rank_per_thread():
    """world size, L, K, dnum, overlap mode, batch and engine flags drawn at random (120 such draws have run clean)"""
    import random
    with sim_engine.simulated() as A:
        flagsets = [0, A.F_DEFER, A.F_GRAPHS, A.F_DEFER | A.F_GRAPHS, A.F_NO_FUSE, A.F_STRICT]
        for seed in range(int(os.environ.get("ALOHA_SWEEP_SHAPES", "12"))):
            rng = random.Random(seed)
            world, L, K = rng.randrange(2, 9), rng.randrange(2, 20), rng.randrange(1, 5)
            dnum = rng.randrange(1, L + 1)
            overlap, batch, flags = rng.choice([False, "chunks", "own"]), rng.randrange(1, 3), rng.choice(flagsets)
            rank_per_thread_case(A, world, 256, L, K, dnum, overlap, batch, passes=2, flags=flags)
