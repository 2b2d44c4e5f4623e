"""The engine's host code on a simulated device (TEST INFRASTRUCTURE -- tests/native/sim/cuda_runtime.h says what
that is).  build() compiles aloha_b200/csrc/{engine,host,group}.cpp with g++ against the simulated runtime and the
contract-checking host kernels into tests/native/libaloha_sim.so; simulated() makes aloha_b200's ctypes face use
that library for the duration of a `with` block, so the very same Python that drives a B200 (A.Engine, HostDriver,
hks.KeySwitch, the GPU tests' own bodies) runs here against the batcher.

What it can show: the batcher's plans compute what the instruction stream says (against the oracle), under every
flag set, and never hand a kernel operands the real kernels cannot take.  What it cannot show: anything about the
sm_100a kernels themselves -- those are only ever checked on a B200 (-m gpu)."""
from __future__ import annotations

import contextlib
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SIM = os.path.join(HERE, "native", "sim")
CSRC = os.path.join(ROOT, "aloha_b200", "csrc")
LIB = os.path.join(HERE, "native", "libaloha_sim.so")
NCCL = os.path.join(HERE, "native", "libaloha_simnccl.so")      # sim_nccl.cpp: in-process collectives for local groups
REPLAY = os.path.join(HERE, "native", "aloha_group_replay_sim") # the C host program linked against the simulated library
SOURCES = [os.path.join(CSRC, f) for f in ("engine.cpp", "host.cpp", "group.cpp")] + \
          [os.path.join(SIM, f) for f in ("sim_cuda.cpp", "sim_kernels.cpp")]
HEADERS = [os.path.join(CSRC, f) for f in ("engine.hpp", "isa.hpp", "kernels.cuh", "aut_plan.hpp", "group_replay_main.cpp")] + \
          [os.path.join(SIM, "sim_nccl.cpp")] + \
          [os.path.join(SIM, f) for f in ("cuda.h", "cuda_runtime.h", "sim.hpp")] + [os.path.join(ROOT, "include", "aloha_b200.h")]


def build() -> str:
    if os.path.exists(LIB) and all(os.path.getmtime(f) <= os.path.getmtime(LIB) for f in SOURCES + HEADERS):
        return LIB
    tmp = f"{LIB}.{os.getpid()}.tmp.so"
    cmd = [os.environ.get("CXX", "g++"), "-std=c++17", "-O2", "-fPIC", "-shared", "-Wall", "-Wno-unused-function",
           "-I" + SIM, *SOURCES, "-o", tmp, "-ldl", "-lpthread"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    here = os.path.dirname(LIB)
    tmp2, tmp3 = f"{NCCL}.{os.getpid()}.tmp.so", f"{REPLAY}.{os.getpid()}.tmp"
    cxx = cmd[0]
    os.replace(tmp, LIB)
    subprocess.run([cxx, "-std=c++17", "-O2", "-fPIC", "-shared", "-Wall", "-I" + SIM, os.path.join(SIM, "sim_nccl.cpp"), "-o", tmp2,
                    "-L" + here, "-l:libaloha_sim.so", "-Wl,-rpath,$ORIGIN"], check=True, capture_output=True, text=True)
    os.replace(tmp2, NCCL)
    subprocess.run([cxx, "-std=c++17", "-O2", "-Wall", os.path.join(CSRC, "group_replay_main.cpp"), "-o", tmp3,
                    "-L" + here, "-l:libaloha_sim.so", "-Wl,-rpath,$ORIGIN"], check=True, capture_output=True, text=True)
    os.replace(tmp3, REPLAY)
    return LIB


@contextlib.contextmanager
def simulated():
    """aloha_b200 bound to the simulated-device library inside the block; the real binding is restored after."""
    import aloha_b200 as A
    path = build()
    saved = A._lib
    os.environ.setdefault("ALOHA_NCCL_LIB", NCCL)
    os.environ.setdefault("ALOHA_SIM_DEVICES", "8")
    A._lib = A.bind_library(path)
    A._lib.sim_reset_violation()        # (a broken contract is sticky inside the library: it must not leak into the next test)          # load_library() returns whatever is bound: no build logic on this path
    try:
        yield A
    finally:
        A._lib = saved
