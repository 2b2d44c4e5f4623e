"""bench.py's key-switch leg walked by EIGHT ranks on the CPU (gloo): the oracle stands in for the engine and
torch.distributed for the NCCL group, everything else is bench.py's own code.  What this guards is the program's
collective discipline -- every rank must reach the same barriers and all-reduces whatever it owns (at eight ranks
one of them holds nothing but special primes; an 8-GPU run once deadlocked on exactly that) -- and that the line
it returns says every rank's output limbs matched the oracle."""
import os
import sys
import types

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class FakeEngine:
    """the Engine method set bench.py and hks.py use, on the CPU golden model"""

    def __init__(self, vlmax_bits, spm_rows, ksk_rows, moduli, **_):
        from oracle import oracle as O
        self.m = O.GoldenModel(vlmax_bits=vlmax_bits, spm_rows=spm_rows, ksk_rows=ksk_rows, moduli=moduli)
        self.launches = 0

    def __getattr__(self, name):
        return getattr(self.m, name)

    def run_vp_multi(self, calls):
        self.launches += 1
        self.m.run_vp_multi(calls)

    def set_stream(self, s):
        pass

    def stats(self):
        return {"kernel_launches": self.launches, "plans_built": 0, "plans_reused": 0, "ops_fused": 0}

    def close(self):
        self.m = None


class FakeGroup:
    """aloha_group_* over gloo, staged through host arrays (hks.TorchComm does the moving)"""

    def __init__(self, engine, rank, world):
        from aloha_b200 import hks
        self.engine, self.rank, self.size, self.comm = engine, rank, world, hks.TorchComm()

    @staticmethod
    def unique_id():
        return b"\0" * 128

    @classmethod
    def create(cls, engine, uid, rank, world):
        return cls(engine, rank, world)

    def all_gather_rows(self, row, rows_per_rank, count=1, stride=0, chunked=False):
        self.comm.all_gather(self.engine, row, rows_per_rank, count, stride, chunked)

    def broadcast_rows(self, row, nrows, root):
        self.comm.broadcast(self.engine, row, nrows, root, 1, 0)

    def wait(self, source=-1):
        pass

    def close(self):
        pass


def _rank_main(rank, world, port, n, qout):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), ALOHA_ORACLE_NATIVE="")
    os.environ.pop("ALOHA_ORACLE_NATIVE", None)
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import bench
        bench.N = n
        bench.KS_SHAPES = [s[:4] + (min(s[4], 2),) for s in bench.KS_SHAPES]      # batch 8 -> 2: same code path
        fake_torch = types.SimpleNamespace(
            cuda=types.SimpleNamespace(synchronize=lambda: None),
            tensor=lambda data, device=None: torch.tensor(data))
        A = types.SimpleNamespace(Engine=FakeEngine, Group=FakeGroup)
        calls = {"timed": 0}

        def timed(fn, steps):
            calls["timed"] += 1
            dist.barrier()
            for _ in range(min(steps, 2)):
                fn()
            dist.barrier()
            ms = torch.tensor([1.0 + rank])
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            return float(ms.item())
        out = bench.measure_keyswitch(fake_torch, dist, A, {}, types.SimpleNamespace(cuda_stream=1), timed, world, rank)
        qout.put((rank, calls["timed"], {k: (v["checked_against_oracle"], v["output_limbs_checked_against_oracle"],
                                              v["transfers"]["overlap_mode"]) for k, v in out.items()}))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,port", [(8, 29561), (4, 29562)])
def test_keyswitch_leg_of_bench_at_many_ranks(world, port):
    ctx = mp.get_context("spawn")
    qout = ctx.Queue()
    procs = [ctx.Process(target=_rank_main, args=(r, world, port, 256, qout)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        results = [qout.get(timeout=300) for _ in range(world)]
    finally:
        for p in procs:
            p.join(timeout=30)
            if p.is_alive():
                p.terminate()
    assert all(p.exitcode == 0 for p in procs)
    timed_calls = {r[1] for r in results}
    assert len(timed_calls) == 1, f"ranks made different numbers of barrier-holding timed() calls: {sorted(r[:2] for r in results)}"
    for rank, _, shapes in results:
        assert len(shapes) == 3
        for name, (ok, nchecked, mode) in shapes.items():
            assert ok and nchecked >= 4 and mode == "own", (rank, name, ok, nchecked, mode)


def test_run_guarded_reports_instead_of_raising_or_hanging():
    """bench.run_guarded: an auxiliary measurement that raises is reported in the line (one rank), and one that
    does not come back makes the process leave cleanly after printing the line (checked in a child process)."""
    import subprocess
    sys.path.insert(0, ROOT)
    import bench
    assert bench.run_guarded(lambda: {"fine": 1}, 5, {}, "k", 1) == {"fine": 1}
    out = bench.run_guarded(lambda: 1 / 0, 5, {}, "k", 1)
    assert "ZeroDivisionError" in out["error"]
    code = ("import sys, time; sys.path.insert(0, %r); import bench; "
            "bench.run_guarded(lambda: time.sleep(60), 1, {'metric': 'm', 'value': 1.0}, 'keyswitch', 8); print('not reached')" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "not reached" not in r.stdout
    import json
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["value"] == 1.0 and "did not finish" in line["keyswitch"]["error"]
