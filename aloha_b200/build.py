"""Build libaloha_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, os.environ.get("ALOHA_LIB_NAME", "libaloha_b200.so"))
REPLAY = os.path.join(HERE, "aloha_group_replay")
SOURCES = ["ntt_kernels.cu", "ew_kernels.cu", "engine.cpp", "host.cpp", "group.cpp"]
HEADERS = ["group_replay_main.cpp", "kernels.cuh", "modarith.cuh", "engine.hpp", "isa.hpp", "aut_plan.hpp", "../../include/aloha_b200.h"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function"] + os.environ.get("ALOHA_NVCC_DEFS", "").split()


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(HERE, "build", os.path.basename(LIB) + "." + src.rsplit(".", 1)[0] + ".o")
        cmd = [NVCC, *FLAGS, "-x", "cu", "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(obj)
    for src, p in procs:
        out = p.communicate()[0].decode()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose or "warning" in out:
            sys.stderr.write(out)
    subprocess.check_call([NVCC, "-Wno-deprecated-gpu-targets", "-shared", "-o", LIB, *objs, "-lcudart_static", "-lpthread", "-ldl", "-lrt"])
    # the C host program for limb-sharded cases: sees nothing but include/aloha_b200.h and the library
    if os.path.basename(LIB) == "libaloha_b200.so":     # (not for A/B variants built under another name)
        subprocess.check_call([os.environ.get("CXX", "g++"), "-std=c++17", "-O2", "-Wall", os.path.join(CSRC, "group_replay_main.cpp"),
                               "-o", REPLAY, "-L" + HERE, "-l:" + os.path.basename(LIB), "-Wl,-rpath,$ORIGIN"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
