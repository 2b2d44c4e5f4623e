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
