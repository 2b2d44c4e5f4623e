"""bench.py's whole GPU arm (`run_gpu`) walked on the CPU at a toy size, with one rank, two and eight (gloo): the
engine is the product's host code on the simulated device (tests/sim_engine.py), NCCL is gloo, torch's CUDA
stream / event / pinned-memory calls are stubbed, and everything else --
argument handling, warm-up policy, the end-to-end and op-chain legs, the automorphism and key-switch legs with
their oracle checks, the assembly of the JSON line, the guarded tail -- is the program's own code.  It cannot say
anything about speed; it says the program reaches its last line on every rank and prints ONE parseable line with
the keys the driver reads."""
import ctypes
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

DRIVER = r'''
import ctypes, json, os, sys, time, types
import numpy as np
sys.path.insert(0, %(root)r)
os.environ["ALOHA_BENCH_NO_SAMPLER"] = "1"
os.environ["ALOHA_SIM_DEVICES"] = "8"
os.environ.pop("ALOHA_ORACLE_NATIVE", None)
import torch, torch.distributed as dist
import aloha_b200 as A
from oracle import oracle as O

sys.path.insert(0, %(root)r + "/tests")
import sim_engine
_sim = sim_engine.simulated()
_sim.__enter__()          # aloha_b200's own Engine class, bound to the engine's host code on the simulated device

class FakeGroup:
    def __init__(self, e, r, w):
        from aloha_b200 import hks
        self.engine, self.rank, self.size, self.comm = e, r, w, hks.TorchComm()
    unique_id = staticmethod(lambda: b"\0" * 128)
    create = classmethod(lambda cls, e, uid, r, w: cls(e, r, w))
    def all_gather_rows(self, row, rpr, count=1, stride=0, chunked=False): self.comm.all_gather(self.engine, row, rpr, count, stride, chunked)
    def broadcast_rows(self, row, nrows, root): self.comm.broadcast(self.engine, row, nrows, root, 1, 0)
    def wait(self, source=-1): pass
    def close(self): pass

A.Group = FakeGroup

# ---- CUDA stubs
class Stream:
    cuda_stream = 1
    def synchronize(self): pass
    def __enter__(self): return self
    def __exit__(self, *a): return False
class Event:
    def __init__(self, enable_timing=False): self.t = 0.0
    def record(self, stream=None): self.t = time.perf_counter()
    def elapsed_time(self, other): return max(1e-3, 1e3 * (other.t - self.t))
torch.cuda.set_device = lambda d: None
torch.cuda.synchronize = lambda *a: None
torch.cuda.Stream = Stream
torch.cuda.set_stream = lambda s: None
torch.cuda.stream = lambda s: s
torch.cuda.Event = Event
torch.Tensor.pin_memory = lambda self: self
_tensor, _empty = torch.tensor, torch.empty
torch.tensor = lambda data, device=None, **k: _tensor(data, **k)
torch.empty = lambda *a, device=None, **k: _empty(*a, **k)
_init = dist.init_process_group
dist.init_process_group = lambda backend, device_id=None, **k: _init("gloo", **k)

import bench
bench.N, bench.LIMBS, bench.ROWS_PER_POLY, bench.ALG_BYTES_PER_NTT = 256, 4, 2, 2 * 256 * 8
bench.KS_SHAPES = [(n, 5, k, d, min(b, 2)) for n, _, k, d, b in bench.KS_SHAPES for d in [min(d, 5)]]
bench.cpu_ntt_rate = lambda primes, psis, nthreads, seconds_target=0: (1.0, "stubbed")
bench.measure_tv_latency = lambda A: {"stubbed": True}
_ge = bench.galois_elements
bench.galois_elements = lambda: [("3^1", 3), ("3^5", pow(3, 5, 2 * bench.N))]
sys.argv = ["bench.py"] + %(argv)r
bench.main()
'''


def run(world, argv, port=29581):
    code = DRIVER % {"root": ROOT, "argv": argv}
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(world))
    procs = []
    for r in range(world):
        e = dict(env, RANK=str(r), LOCAL_RANK=str(r))
        procs.append(subprocess.Popen([sys.executable, "-c", code], env=e, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = []
    for p in procs:
        try:
            o, err = p.communicate(timeout=420)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            pytest.fail("bench.py did not finish on every rank (a collective out of step?)")
        assert p.returncode == 0, err[-3000:]
        outs.append(o)
    return outs


def the_line(stdout):
    lines = [l for l in stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, stdout[-2000:]
    return json.loads(lines[0])


def test_one_rank_prints_the_contract_line():
    line = the_line(run(1, ["--gpus", "1", "--steps", "2", "--warmup", "3", "--polys", "8"])[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in line, key
    assert line["n_gpus"] == 1 and line["config"]["workload"] and line["dtype"] == "u64"
    assert set(line["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step", "host_link_ceiling", "op_chain"}
    assert set(line["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    assert line["automorphism"]["all_checked_against_oracle"] is True
    assert set(line["keyswitch"]) == {"digit_1_limb", "dnum5_k8", "dnum5_k8_batch8"}
    assert all(v["checked_against_oracle"] for v in line["keyswitch"].values())
    assert line["generic_primes"]["value"] > 0 and line["cpu_baseline"]["gpu_output_checked_against_oracle"] is True


def test_two_ranks_reach_the_end_together():
    outs = run(2, ["--gpus", "2", "--steps", "2", "--warmup", "3", "--polys", "8"], port=29582)
    line = the_line(outs[0])
    assert not [l for l in outs[1].splitlines() if l.startswith("{")]          # rank 0 alone prints
    assert line["n_gpus"] == 2 and line["scaling"] == "weak" and len(line["per_rank"]["ms_timed_region"]) == 2
    assert all(v["checked_against_oracle"] and v["transfers"]["overlap_mode"] == "own" for v in line["keyswitch"].values())
    assert "automorphism" not in line and "cpu_baseline" not in line           # one-GPU legs


def test_eight_ranks_reach_the_end_together():
    """the driver's largest scaling point: one rank owns nothing but special primes in the key-switch leg"""
    outs = run(8, ["--gpus", "8", "--steps", "2", "--warmup", "3", "--polys", "8"], port=29583)
    line = the_line(outs[0])
    assert all(not [l for l in o.splitlines() if l.startswith("{")] for o in outs[1:])
    assert line["n_gpus"] == 8 and len(line["per_rank"]["ms_timed_region"]) == 8
    assert all("error" not in v and v["checked_against_oracle"] for v in line["keyswitch"].values())


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, env=dict(os.environ, ALOHA_ORACLE_NATIVE=""))
    assert out.returncode == 0, out.stderr[-2000:]
    line = the_line(out.stdout)
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
