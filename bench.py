#!/usr/bin/env python
"""bench.py -- limb-NTTs/s at N = 2^16 (BASELINE.json's metric) on N B200s of one node.

Workload (BASELINE.json configs[2]): batched negacyclic NTT, N = 65536, 32 RNS limbs
(the first 32 primes below 2^60 with q = 1 mod 2^17, psi = minimal primitive 2N-th root), B = 64
polynomials per limb => 2048 limb-NTTs per step, 1 GiB in + 1 GiB out (larger than the 126 MB L2, so
no flush is needed between steps).  One step = the instruction stream
`VSETQ/IQ; VLE; VNTT; VSE` x 32 limbs (aloha_b200.asm.transform_stream) issued for all 64
polynomials through aloha_run_vp_batch -- the reference-facing C-ABI.

  value : device-resident throughput (inputs already in the engine's SPM in HBM), CUDA-event timed
  e2e   : the same step with HOST buffers: pinned host -> aloha_dma_mem_h2d -> run -> aloha_dma_mem_d2h
  roofline : algorithmic bytes (2*N*8 per limb-NTT) / event time of the transform kernels, against the
             measured HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline : the oracle (C++ golden model, a port of the RTL's algorithm) on the box's host cores,
             bounded sample.  `--impl reference` times only that.

Multi-GPU (torchrun, one rank per GPU): the limbs of each polynomial batch are independent units, so
every rank transforms its own 32-limb x 64-poly shard with no data-path collective ("weak" scaling);
value = limb-NTTs of all ranks / max-over-ranks time.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# the timed CPU reference is compiled -O3 -march=native on the machine that times it
os.environ.setdefault("ALOHA_ORACLE_NATIVE", "1")

N = 65536
LIMBS = 32
POLYS = 64
ROWS_PER_POLY = N // 128
ALG_BYTES_PER_NTT = 2 * N * 8


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw,clocks.mem")

    def __init__(self, indices, enabled=True, period_ms=200):
        """One nvidia-smi process for all the job's GPUs (rank 0 only: NVML polling from every rank
        perturbs the launches it is supposed to observe)."""
        self.indices, self.rows, self.proc, self.enabled = list(indices), [], None, enabled
        self.period_ms = period_ms

    def __enter__(self):
        if not self.enabled:
            return self
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", str(self.period_ms),
                 "-i", ",".join(str(i) for i in self.indices)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        rows = [r for r in self.rows if len(r) >= 7 and r[1].isdigit()]
        sm = [int(r[1]) for r in rows]
        mx = [int(r[2]) for r in rows if r[2].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows for i in range(4) if r[3 + i] == "Active"})
        out = {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
               "reasons": reasons, "samples": len(sm)}
        try:
            pw = [float(r[7]) for r in rows if len(r) > 7]
            out["power_w_max"] = max(pw) if pw else None
            out["mem_mhz_min"] = min(int(r[8]) for r in rows if len(r) > 8 and r[8].isdigit())
        except Exception:
            pass
        if len(self.indices) > 1:
            per = {}
            for g in self.indices:
                mine = [r for r in rows if r[0] == str(g)]
                if not mine:
                    continue
                pw = [float(r[7]) for r in mine if len(r) > 7]
                per[g] = {"sm_mhz_median": int(np.median([int(r[1]) for r in mine])), "sm_mhz_min": min(int(r[1]) for r in mine),
                          "power_w_max": max(pw) if pw else None, "power_w_median": float(np.median(pw)) if pw else None}
            out["per_gpu"] = per
        return out


def bind_to_gpu_numa_node(gpu: int):
    """Pin this rank to the CPU cores nearest its GPU (nvidia-smi topo) BEFORE it allocates page-locked host
    memory, so the staging buffers of the end-to-end path are first-touched on the GPU's own NUMA node instead
    of all ranks' buffers landing on one memory controller.  Returns the affinity string, or None."""
    try:
        out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
        hdr = None
        for line in out.splitlines():
            cols = line.split("\t")
            if hdr is None and "CPU Affinity" in line:
                hdr = [c.strip() for c in cols]
                continue
            if hdr and cols and cols[0].strip() == f"GPU{gpu}":
                idx = hdr.index("CPU Affinity")
                spec = cols[idx].strip() if idx < len(cols) else ""
                cores = set()
                for part in spec.split(","):
                    if "-" in part:
                        a, b = part.split("-")
                        cores.update(range(int(a), int(b) + 1))
                    elif part.strip().isdigit():
                        cores.add(int(part))
                cores &= os.sched_getaffinity(0)
                if cores:
                    os.sched_setaffinity(0, cores)
                    return spec
    except Exception:
        pass
    return None


def run_guarded(fn, seconds, line, key, world):
    """fn() on every rank; if it raises, or does not return within `seconds` on this rank, the rank leaves the job
    cleanly: rank 0 (the one holding `line`) prints the line with line[key] = {"error": ...} first.  A rank that
    raised cannot rejoin the others' collectives, so with several ranks it leaves at once and the rest follow
    through their own watchdogs."""
    def bail(reason):
        if line is not None:
            line[key] = {"error": reason}
            print(json.dumps(line), flush=True)
        sys.stdout.flush()
        os._exit(0)
    timer = threading.Timer(seconds, bail, args=(f"did not finish within {seconds} s on this rank; skipped",))
    timer.daemon = True
    timer.start()
    try:
        return fn()
    except Exception as e:                      # noqa: BLE001 -- reported in the line, not swallowed
        timer.cancel()
        if world > 1:
            bail(f"{type(e).__name__}: {e}")
        return {"error": f"{type(e).__name__}: {e}"}
    finally:
        timer.cancel()


def workload_params():
    from aloha_b200 import params
    primes = params.synthetic_primes(LIMBS, 2 * N)
    psis = [params.min_primitive_root(q, 2 * N) for q in primes]
    return primes, psis


def synth_batch(primes, polys, seed):
    rng = np.random.default_rng(seed)
    qv = np.array(primes, dtype=np.uint64)[None, :, None]
    return rng.integers(0, 1 << 59, (polys, LIMBS, N), dtype=np.uint64) % qv


# ------------------------------------------------------------------------------ reference arm (CPU)
def cpu_ntt_rate(primes, psis, nthreads, seconds_target=12.0):
    """Oracle NTT on a bounded sample of the same workload; returns (limb-NTTs/s, sample text)."""
    from oracle import oracle as O
    tabs = O.NttTables(N, primes, psis)
    count = max(nthreads * 2, 16)
    x = synth_batch(primes, 1, 0xA10A)[0]
    reps = (count + LIMBS - 1) // LIMBS
    a = np.ascontiguousarray(np.concatenate([x] * reps)[:count])
    idx = (np.arange(count) % LIMBS).astype(np.uint32)
    t0 = time.perf_counter()
    tabs.batch(a, idx, nthreads=nthreads)
    dt = time.perf_counter() - t0
    done, total = count, dt
    while total < seconds_target:
        t0 = time.perf_counter()
        tabs.batch(a, idx, nthreads=nthreads)
        total += time.perf_counter() - t0
        done += count
    return done / total, f"{done} limb-NTTs (N=2^16, limbs cycled over the 32 primes) in {total:.1f} s"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    primes, psis = workload_params()
    cores = os.cpu_count() or 1
    from oracle import oracle as O
    tabs = O.NttTables(N, primes, psis)
    count = max(2 * cores, 16)   # one step = a bounded sample: `count` limb-NTTs across all cores
    x = synth_batch(primes, 1, 0xA10A)[0]
    a = np.ascontiguousarray(np.concatenate([x] * ((count + LIMBS - 1) // LIMBS))[:count])
    idx = (np.arange(count) % LIMBS).astype(np.uint32)
    for _ in range(args.warmup):
        tabs.batch(a, idx, nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        tabs.batch(a, idx, nthreads=cores)
    dt = time.perf_counter() - t0
    value = args.steps * count / dt
    sample = f"{count} limb-NTTs per step x {args.steps} steps, {cores} threads"
    print(json.dumps({
        "impl": "reference", "metric": "limb_ntts_per_s_n65536", "value": value, "unit": "limb-NTTs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": dict(workload_config(), reference_arm_sample=sample + " (a bounded sample of the 2048-limb-NTT step: the rate is what is compared)"),
        "cpu_baseline": {"value": value, "unit": "limb-NTTs/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "limb-NTTs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config():
    return {"workload": "batched negacyclic NTT, N=65536, 32 RNS limbs x 64 polynomials per GPU "
                        "(BASELINE.json configs[2])",
            "n": N, "limbs": LIMBS, "polys": POLYS, "limb_ntts_per_step_per_gpu": LIMBS * POLYS,
            "moduli": "the 32 primes below 2^60 with q = 1 mod 2^17, scanning downward (SURVEY 8(d)3): all of the form "
                      "2^60 - d with d < 2^27, which the engine detects and serves with its pseudo-Mersenne product",
            "l2": "inputs (1 GiB) + outputs (1 GiB) per step exceed L2; no flush needed",
            "warmup_policy": "W steps, then continuous load until 0.4 s have passed (sustained clocks; sw_power_cap may be active)",
            "parallelism": "limb/poly-sharded, no data-path collective"}


# ------------------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    import aloha_b200 as A
    from aloha_b200 import asm

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    primes, psis = workload_params()
    per_poly = LIMBS * ROWS_PER_POLY
    rows = POLYS * per_poly
    if args.only:
        stream = torch.cuda.Stream()
        torch.cuda.set_stream(stream)

        def timed0(fn, steps):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(steps):
                fn()
            e1.record(stream)
            torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            return float(ms.item())
        if args.only == "keyswitch":
            out = measure_keyswitch(torch, dist, A, {"device": local}, stream, timed0, world, rank,
                                    shapes=args.shapes.split(",") if args.shapes else None)
        elif args.only == "tv":
            out = measure_tv_latency(A) if rank == 0 else None
        elif args.only == "rotmac_tiled":
            out = measure_rotmac(torch, A, asm, {"device": local}, primes, psis, stream, timed0, args.polys, flags=A.F_AUT_TILED,
                                 quick=args.quick, only_k=args.galois)
        elif args.only == "rotmac_gather":
            out = measure_rotmac(torch, A, asm, {"device": local}, primes, psis, stream, timed0, args.polys, flags=A.F_AUT_GATHER,
                                 quick=args.quick, only_k=args.galois)
        else:
            out = measure_rotmac(torch, A, asm, {"device": local}, primes, psis, stream, timed0, args.polys, quick=args.quick,
                                 only_k=args.galois)
        if rank == 0:
            print(json.dumps(out), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return
    # renaming pool: the op chain of the end-to-end leg keeps ~3 live temporaries per limb of a 16-polynomial chunk
    eng = A.Engine(vlmax_bits=N * 64, spm_rows=2 * rows, ksk_rows=0, device=local, moduli=list(zip(primes, psis)),
                   l2_chunk_bytes=args.chunk_mib << 20, pool_buffers=4096,
                   flags=(A.F_GRAPHS if args.graphs else 0) | (A.F_GENERIC_MODMUL if args.generic else 0))
    # A dedicated non-default torch stream: the engine launches on it (aloha_set_stream) and the CUDA
    # events below are recorded on it.  (Stream handle 0 would mean "the engine's own stream".)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    eng.set_stream(stream.cuda_stream)
    eng.load_isram(asm.transform_stream(N, primes).words(), 0)
    eng.load_isram(asm.transform_stream(N, primes, inverse=True).words(), 1024)
    calls = A.Engine.make_args([(b * per_poly, 0, rows + b * per_poly, 0, 0) for b in range(POLYS)])
    back = A.Engine.make_args([(rows + b * per_poly, 0, b * per_poly, 0, 0) for b in range(POLYS)])

    host_in = torch.empty(POLYS * LIMBS * N, dtype=torch.int64).pin_memory()
    host_out = torch.empty(POLYS * LIMBS * N, dtype=torch.int64).pin_memory()
    x = synth_batch(primes, POLYS, 0xA10A + rank)
    host_in.numpy().view(np.uint64)[:] = x.reshape(-1)
    nbytes = host_in.numel() * 8
    eng.dma_mem_h2d(0, (host_in.data_ptr(), nbytes))
    eng.sync()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        h0 = time.perf_counter()
        for _ in range(steps):
            fn()
        host_ms = 1e3 * (time.perf_counter() - h0)      # what the host needed to enqueue the region (it does not wait in fn)
        e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1), host_ms], device="cuda")
        timed.last_host_ms = [host_ms]
        if world > 1:
            allms = [torch.zeros_like(ms) for _ in range(world)]
            dist.all_gather(allms, ms)
            timed.last_per_rank = [float(t[0].item()) for t in allms]
            timed.last_host_ms = [float(t[1].item()) for t in allms]
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms[0].item())

    step = lambda: eng.run_vp_batch(0, calls)
    # Warm-up: at least W (>= 3) steps, and enough of them to keep every GPU busy for ~0.4 s -- with
    # several GPUs starting from idle at once the SM clocks need tens of milliseconds to ramp up under
    # the node's power management, and a 30 ms timed region would otherwise measure the ramp.  The
    # clock sampler runs from the start of the warm-up to the end of the timed region.
    with ClockSampler(range(world), enabled=(rank == 0 and not os.environ.get("ALOHA_BENCH_NO_SAMPLER")),
                      period_ms=200) as clk:
        n_warm = max(args.warmup, 3)
        for _ in range(n_warm):
            step()
        barrier()
        if not args.quick:                      # --quick (profiling): exactly W warm-ups
            t0 = time.perf_counter()
            for _ in range(8):
                step()
            torch.cuda.synchronize()
            per_step = max((time.perf_counter() - t0) / 8, 1e-4)
            extra = torch.tensor([int(0.4 / per_step)], device="cuda")
            if world > 1:                       # every rank runs the same count, so they stay aligned and
                dist.all_reduce(extra, op=dist.ReduceOp.MAX)   # nobody cools down waiting at the barrier
            for _ in range(int(extra.item())):
                step()
            n_warm += 8 + int(extra.item())
        s0 = eng.stats()
        ms = timed(step, args.steps)
    s1 = eng.stats()
    per_rank_ms = getattr(timed, "last_per_rank", None)
    host_enqueue_ms = list(timed.last_host_ms)          # of the headline's timed region: a rank whose host time approaches
                                                        # its device time is enqueue-bound, not GPU-bound
    launches = s1["kernel_launches"] - s0["kernel_launches"]
    ntts_per_step = LIMBS * POLYS
    value = world * ntts_per_step * args.steps / (ms / 1e3)

    burst = None
    if world == 1 and not args.quick:
        # the same K steps from a cool start (1 s idle, 3 warm-ups): the regime before the power cap
        # engages -- reported next to the sustained `value`, never instead of it
        torch.cuda.synchronize()
        time.sleep(1.0)
        for _ in range(3):
            step()
        ms_b = timed(step, args.steps)
        burst = {"value": LIMBS * POLYS * args.steps / (ms_b / 1e3), "unit": "limb-NTTs/s",
                 "ms_per_step": ms_b / args.steps, "note": "3 warm-up steps after 1 s idle; `value` is after >= 0.4 s of continuous load"}
    if args.quick:
        if rank == 0:
            print(json.dumps({"quick": True, "value": value, "ms_per_step": ms / args.steps, "gpu_launches": launches,
                              "chunk_mib": args.chunk_mib, "polys": POLYS, "clocks": clk.summary()}))
        if world > 1:
            dist.destroy_process_group()
        return

    # inverse transform rate (same machinery, reported alongside)
    for _ in range(3):
        eng.run_vp_batch(1024, back)
    ms_inv = timed(lambda: eng.run_vp_batch(1024, back), args.steps)
    inv_value = world * ntts_per_step * args.steps / (ms_inv / 1e3)

    # the same workload with the any-prime (Shoup) arithmetic every kernel also carries -- what moduli that are
    # not of the form 2^60 - d (the reference's own q0, q1, P) get
    generic = None
    if world == 1 and not args.generic:
        eng.close()
        eng = A.Engine(vlmax_bits=N * 64, spm_rows=2 * rows, ksk_rows=0, device=local, moduli=list(zip(primes, psis)),
                       flags=A.F_GENERIC_MODMUL)
        eng.set_stream(stream.cuda_stream)
        eng.load_isram(asm.transform_stream(N, primes).words(), 0)
        eng.load_isram(asm.transform_stream(N, primes, inverse=True).words(), 1024)
        eng.dma_mem_h2d(0, (host_in.data_ptr(), nbytes))
        for _ in range(5):
            eng.run_vp_batch(0, calls)
        ms_g = timed(lambda: eng.run_vp_batch(0, calls), args.steps)
        for _ in range(3):
            eng.run_vp_batch(1024, back)
        ms_gi = timed(lambda: eng.run_vp_batch(1024, back), args.steps)
        generic = {"value": ntts_per_step * args.steps / (ms_g / 1e3), "unit": "limb-NTTs/s", "ms_per_step": ms_g / args.steps,
                   "intt": ntts_per_step * args.steps / (ms_gi / 1e3),
                   "frac_of_hbm_peak": ntts_per_step * args.steps / (ms_g / 1e3) * ALG_BYTES_PER_NTT / 1e9 / peaks()[0],
                   "note": "same primes, ALOHA_F_GENERIC_MODMUL: Shoup / Harvey arithmetic (10 IMAD per product), the path any 60-bit prime takes"}
        eng.close()
        eng = A.Engine(vlmax_bits=N * 64, spm_rows=2 * rows, ksk_rows=0, device=local, moduli=list(zip(primes, psis)),
                       l2_chunk_bytes=args.chunk_mib << 20, pool_buffers=4096)
        eng.set_stream(stream.cuda_stream)
        eng.load_isram(asm.transform_stream(N, primes).words(), 0)
        eng.load_isram(asm.transform_stream(N, primes, inverse=True).words(), 1024)
        eng.dma_mem_h2d(0, (host_in.data_ptr(), nbytes))
        eng.run_vp_batch(0, calls)
        eng.sync()

    extra = {}
    if not args.no_extra:
        eng_kwargs = {"device": local}
        if world == 1:
            extra["automorphism"] = measure_rotmac(torch, A, asm, eng_kwargs, primes, psis, stream, timed, POLYS)
            extra["tv_latency"] = measure_tv_latency(A)

    # end to end: host buffers in, host buffers out, every step
    # Host buffers -> SPM -> NTT -> host buffers, through the C-ABI only.  The batch is cut into
    # chunks so the upload of chunk c+1, the transform of chunk c and the download of chunk c-1
    # overlap (aloha_dma_mem_*_async = the DMA block running beside the VP).
    n_chunks = int(os.environ.get("ALOHA_E2E_CHUNKS", "16"))
    if POLYS % n_chunks:
        n_chunks = 1
    cp = POLYS // n_chunks
    chunk_bytes = cp * LIMBS * N * 8
    chunk_calls = [A.Engine.make_args([(b * per_poly, 0, rows + b * per_poly, 0, 0)
                                       for b in range(c * cp, (c + 1) * cp)]) for c in range(n_chunks)]

    def e2e_step():
        for c in range(n_chunks):
            eng.dma_mem_h2d_async(c * cp * per_poly, host_in.data_ptr() + c * chunk_bytes, chunk_bytes)
            eng.run_vp_batch(0, chunk_calls[c])
            eng.dma_mem_d2h_async(host_out.data_ptr() + c * chunk_bytes, rows + c * cp * per_poly, chunk_bytes)
        eng.sync()
    host_out.zero_()
    e2e_step()
    e2e_steps = max(1, min(args.steps, 5))
    ms_e2e = timed(e2e_step, e2e_steps)
    e2e_value = world * ntts_per_step * e2e_steps / (ms_e2e / 1e3)

    # What the host link allows: the same two pinned buffers moved up and down concurrently by plain
    # cudaMemcpyAsync (torch copies on two streams), no kernels -- the ceiling of any per-step host round trip.
    dev_in = torch.empty(host_in.numel(), dtype=torch.int64, device="cuda")
    dev_out = torch.empty(host_in.numel(), dtype=torch.int64, device="cuda")
    s_up, s_down = torch.cuda.Stream(), torch.cuda.Stream()

    def copy_step():
        with torch.cuda.stream(s_up):
            dev_in.copy_(host_in, non_blocking=True)
        with torch.cuda.stream(s_down):
            host_out2.copy_(dev_out, non_blocking=True)
        s_up.synchronize()
        s_down.synchronize()
    host_out2 = torch.empty(host_in.numel(), dtype=torch.int64).pin_memory()
    copy_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        copy_step()
    barrier()
    copy_s = torch.tensor([(time.perf_counter() - t0) / 3], device="cuda")
    if world > 1:
        dist.all_reduce(copy_s, op=dist.ReduceOp.MAX)
    copy_ms = 1e3 * float(copy_s.item())
    del dev_in, dev_out, host_out2
    # The reference's actual flow (top_noaxilite_tb.sv:450-496, 596-638): upload the operands ONCE, run a chain of
    # ops on resident data, download the result ONCE.  Chain: NTT -> element-wise product with a resident
    # plaintext -> INTT on every limb (encode_post, mul_plain and their inverse at N = 2^16), 3 kernels' worth of
    # work per host round trip.
    chain_depth = 4
    chain = asm.Program().vsetvl(N)
    for l, q in enumerate(primes):
        chain.vsetq(q).vle(0, asm.BASE_SRC0, l * ROWS_PER_POLY).vle(1, asm.BASE_SRC1, l * ROWS_PER_POLY)
        src = 0
        for _ in range(chain_depth):
            chain.vntt(2, src).vfqmul(4, 2, 1).vintt(6, 4)
            src = 6
        chain.vse(6, asm.BASE_RSLT, l * ROWS_PER_POLY)
    eng.load_isram(chain.brk().words(), 2048)
    cn = 4 if POLYS % 4 == 0 else 1                   # fewer, larger chunks: every launch of the chain fills the GPU
    ccp, cbytes = POLYS // cn, POLYS // cn * LIMBS * N * 8
    chain_calls = [A.Engine.make_args([(b * per_poly, 0, rows + b * per_poly, 0, 0) for b in range(c * ccp, (c + 1) * ccp)])
                   for c in range(cn)]                # src1 = row 0: polynomial 0 doubles as the plaintext

    def chain_step():
        for c in range(cn):
            eng.dma_mem_h2d_async(c * ccp * per_poly, host_in.data_ptr() + c * cbytes, cbytes)
            eng.run_vp_batch(2048, chain_calls[c])
            eng.dma_mem_d2h_async(host_out.data_ptr() + c * cbytes, rows + c * ccp * per_poly, cbytes)
        eng.sync()
    chain_step()
    sc0 = eng.stats()
    ms_chain = timed(chain_step, e2e_steps)
    chain_launches = (eng.stats()["kernel_launches"] - sc0["kernel_launches"]) / e2e_steps

    ok = True
    if rank == 0 and world == 1:
        # cpu_baseline leg, part 1: the oracle checks two limb-polys of what was just timed
        from oracle import oracle as O
        e2e_step()
        got = host_out.numpy().view(np.uint64).reshape(POLYS, LIMBS, N)
        tabs = O.NttTables(N, primes, psis)
        sel = np.array([0, LIMBS - 1])
        ok = bool((got[POLYS - 1, sel] == tabs.batch(x[POLYS - 1, sel].copy(), sel)).all())

    if rank == 0:
        peak, peak_src = peaks()
        per_gpu_rate = ntts_per_step * args.steps / (ms / 1e3)
        achieved = per_gpu_rate * ALG_BYTES_PER_NTT / 1e9
        cores = os.cpu_count() or 1
        cpu_all, sample_all = cpu_ntt_rate(primes, psis, cores, 10.0) if world == 1 else (None, None)
        cpu_one, _ = cpu_ntt_rate(primes, psis, 1, 4.0) if world == 1 else (None, None)
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r1_ntt_traffic.json")))
            traffic = tr["total"] * (POLYS * LIMBS / 2048.0)
        except Exception:
            traffic = None
        line = {
            "metric": "limb_ntts_per_s_n65536", "value": value, "unit": "limb-NTTs/s", "n_gpus": world,
            "steps": args.steps, "warmup": n_warm, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic", "config": workload_config(),
            "clocks": clk.summary(),
            "e2e": {"value": e2e_value, "unit": "limb-NTTs/s", "h2d_bytes_per_step": nbytes,
                    "d2h_bytes_per_step": nbytes, "ms_per_step": ms_e2e / e2e_steps,
                    "pipeline": f"{n_chunks} chunks, upload / transform / download overlapped (pinned host memory)",
                    "host_link_ceiling": {"ms_per_step": copy_ms, "what": "the same 1 GiB up + 1 GiB down per GPU as two concurrent plain "
                                          "cudaMemcpyAsync on the same pinned buffers, no kernels; max over ranks",
                                          "gb_per_s_each_way": nbytes / copy_ms / 1e6,
                                          "e2e_fraction_of_ceiling": copy_ms / (ms_e2e / e2e_steps)},
                    "op_chain": {"value": world * 2 * chain_depth * ntts_per_step * e2e_steps / (ms_chain / 1e3), "unit": "limb-NTTs/s",
                                 "ms_per_step": ms_chain / e2e_steps, "depth": chain_depth, "launches_per_step": chain_launches,
                                 "what": "the reference's flow: operands uploaded once, a chain of ops on resident data (here `depth` rounds of "
                                         "NTT -> product with a resident plaintext -> INTT per limb), result downloaded once",
                                 "limb_ntts_per_step_per_gpu": 2 * chain_depth * ntts_per_step}},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "traffic_note": "dram__bytes_read+write of the column + row kernels for one step, from one ncu --set full capture (profiles/r1_ntt_traffic.json)",
                         "co_bound": {"what": "integer issue: the butterflies alone (registers only, no memory; tools/bf_bench.cu) run at "
                                              "4.09 butterflies/clk/SM with the pseudo-Mersenne product (3.02 with Shoup's), i.e. the "
                                              "arithmetic caps this two-pass transform at 2.27 M limb-NTTs/s; the board's 600 W software "
                                              "power cap holds the sustained figure ~7 % under the burst one",
                                      "arithmetic_only_ceiling_limb_ntts_per_s": 2.27e6,
                                      "arithmetic_only_ceiling_generic_primes_limb_ntts_per_s": 1.67e6,
                                      "see": "DESIGN.md section 5"},
                         "kernel": "ntt_fwd_cols<8,1> + ntt_fwd_rows_tma<8,1> (one limb-NTT = one column pass + one TMA-staged row pass; "
                                   "<.,1> = pseudo-Mersenne arithmetic, which the workload's prime rule selects)",
                         "algorithmic_bytes_per_limb_ntt": ALG_BYTES_PER_NTT},
            "intt": {"value": inv_value, "unit": "limb-NTTs/s", "ms_per_step": ms_inv / args.steps},
            "generic_primes": generic,
            "engine_stats": {k: s1[k] - s0[k] for k in s1},
        }
        if burst:
            line["burst"] = burst
        line["host_enqueue_ms"] = host_enqueue_ms        # per rank, for the headline's timed region (ms_per_step * steps on the device)
        if world > 1:
            line["per_rank"] = {"ms_timed_region": per_rank_ms, "host_enqueue_ms": host_enqueue_ms, "cpu_affinity_of_rank0": numa}
        line.update(extra)
        if world == 1:
            line["cpu_baseline"] = {"value": cpu_all, "unit": "limb-NTTs/s", "cores": cores, "kind": "port",
                                    "sample": sample_all, "single_thread": cpu_one, "gpu_output_checked_against_oracle": ok}
    if not args.no_extra:
        # The key-switch streams (config 5) come last and under a watchdog: they are the only part of this program
        # with a data-path collective, and an auxiliary measurement must not take the headline line down with it.
        eng.close()
        ks = run_guarded(lambda: measure_keyswitch(torch, dist, A, {"device": local}, stream, timed, world, rank),
                         400, line if rank == 0 else None, "keyswitch", world)
        if rank == 0:
            line["keyswitch"] = ks
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if not ok:
        sys.exit(3)


def measure_tv_latency(A):
    """tv-case latency vs CPU (BASELINE.json metric, second half): the three shipped cases replayed
    end to end on the GPU engine through the C host driver, and (cpu_baseline leg) on the oracle
    single-threaded.  Inputs: the committed fixtures under tests/golden (no oracle code on the GPU leg)."""
    gold = os.path.join(ROOT, "tests", "golden")
    manifest = json.load(open(os.path.join(gold, "manifest.json")))
    pool = np.load(os.path.join(gold, "pool.npz"))
    micro = json.load(open(os.path.join(gold, "microcode.json")))
    kernels = [(np.frombuffer(bytes.fromhex("".join(k["words"])), dtype=np.uint8).reshape(-1, 12).copy(), k["pc"])
               for k in micro.values()]
    dram_vp_base = 10485760                                   # top_noaxilite_tb.sv:45
    out = {}
    for case in ("case0_4_4", "case1_8_8", "case2_16_16"):
        n, entry = manifest["n"], manifest["cases"][case]
        prog_lines = entry["program"]
        dram_addr = {i: (int(l.split(",")[1], 16) << 32) | int(l.split(",")[2], 16) for i, l in enumerate(prog_lines)}

        def gpu_once(flags=0, dumps=True):
            eng = A.Engine(flags=flags)
            for words, pc in kernels:
                eng.load_isram(words, pc)
            for row, key in entry["ksk"].items():
                eng.dma_ksk_h2d(int(row), pool[key])
            host = A.HostDriver(eng, "\n".join(prog_lines), n)
            for i, key in entry["loads"].items():
                host.dram_write(dram_vp_base + dram_addr[int(i)], pool[key])
            for i, key in entry["encoder"].items():
                host.set_encoder_output(int(i), pool[key])
            eng.sync()
            run = host.run_op if dumps else host.run_op_nodump

            def one_pass():
                if dumps == "async":          # every dump the testbench writes, read-backs overlapped
                    host.run_all_async()
                else:
                    for i in range(len(host)):
                        run(i)
                    eng.sync()
            t0 = time.perf_counter()
            one_pass()
            dt = time.perf_counter() - t0
            best = 1e9
            for _ in range(5):                # later passes: plans are cached (steady state)
                t1 = time.perf_counter()
                one_pass()
                best = min(best, time.perf_counter() - t1)
            st = eng.stats()
            eng.close()
            return dt, best, st
        first, steady, st = gpu_once()
        _, steady_async, _ = gpu_once(dumps="async")
        _, steady_nodump, _ = gpu_once(dumps=False)
        _, steady_graphs, _ = gpu_once(flags=A.F_GRAPHS, dumps=False)
        out[case] = {"ops": len(prog_lines), "gpu_ms_first_run": 1e3 * first, "gpu_ms_steady": 1e3 * steady,
                     "gpu_ms_steady_async_dumps": 1e3 * steady_async,
                     "gpu_ms_steady_no_dumps": 1e3 * steady_nodump, "gpu_ms_steady_no_dumps_cuda_graphs": 1e3 * steady_graphs,
                     "kernel_launches_per_pass": st["kernel_launches"] / 6.0,
                     "note": "gpu_ms_steady includes every per-op DMA and the blocking 256 KiB dump read-back the testbench does "
                             "after each op; gpu_ms_steady_async_dumps produces the same dumps through aloha_host_run_op_async "
                             "(read-backs on the download channel, one sync at the end; includes the Python-side allocation of "
                             "the dump arrays); the no_dumps figures run the same ops with one sync at the end"}
    # cpu_baseline leg: the same replays on the oracle, one thread
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import golden_util as G
    from oracle import oracle as O
    for case in out:
        ops, dram, enc, ksk = G.case_inputs(case)
        model = O.GoldenModel()
        for words, pc in G.microcode():
            model.load_isram(words, pc)
        for row, data in ksk.items():
            model.dma_ksk_h2d(row, data)
        t0 = time.perf_counter()
        for _ in O.replay(model, ops, dram.copy(), enc, manifest["n"]):
            pass
        out[case]["cpu_oracle_ms_1thread"] = 1e3 * (time.perf_counter() - t0)
    return out


def galois_elements():
    """BASELINE.json configs[3]: rotation Galois elements 3^step mod 2N (the reference host's convention,
    top_noaxilite_tb.sv:531) for steps 1, 2, 8, N/8 (= N/2 + 1: the most skewed lattice there is) and 18, the
    first step whose element is >= N (102729; k mod N = 37193: a pseudo-random permutation, top bit in the signs)."""
    ks = [("3^1", pow(3, 1, 2 * N)), ("3^2", pow(3, 2, 2 * N)), ("3^8", pow(3, 8, 2 * N)), ("3^(N/8)", pow(3, N // 8, 2 * N)),
          ("3^18", pow(3, 18, 2 * N))]
    assert any(k >= N and k % N != 1 for _, k in ks)
    if os.environ.get("ALOHA_BENCH_GALOIS"):        # experiments: "name:k,name:k"
        ks = [(a.split(":")[0], int(a.split(":")[1])) for a in os.environ["ALOHA_BENCH_GALOIS"].split(",")]
    return ks


def measure_rotmac(torch, A, asm, eng_kwargs, primes, psis, stream, timed, polys, flags=0, quick=False, only_k=""):
    """BASELINE.json configs[3]: automorphism and rotate-and-sum at N = 2^16, 32 limbs x `polys` polynomials.
      vaut       : rslt[l] = aut_k(x[l])                 algorithmic bytes 2*N*8 per limb
      rotate_mac : rslt[l] = acc[l] + aut_k(x[l]) * p[l] algorithmic bytes 4*N*8 per limb (one fused kernel)
    for every Galois element of galois_elements(); one limb-polynomial per element is checked against the
    oracle inside the run (cpu_baseline leg).  The engine is built with vlmax = 2N so that k >= N keeps
    its top bit (on a vlmax = N machine the RTL truncates k to log2 N bits)."""
    from oracle import oracle as O
    peak, _ = peaks()
    per_poly = LIMBS * ROWS_PER_POLY
    eng = A.Engine(vlmax_bits=2 * N * 64, spm_rows=4 * polys * per_poly, ksk_rows=0, moduli=(), pool_buffers=2 * LIMBS + 8,
                   flags=flags, **eng_kwargs)
    eng.set_stream(stream.cuda_stream)
    prog_mac, prog_aut = asm.rotate_mac_stream(N, primes), asm.vaut_stream(N, primes)
    eng.load_isram(prog_mac.words(), 0)
    eng.load_isram(prog_aut.words(), 2048)
    rng = np.random.default_rng(7)
    qv = np.array(primes, dtype=np.uint64)
    # 8 distinct polynomials tiled over the batch (host memory: 3 x 128 MiB instead of 3 x 1 GiB)
    base = [rng.integers(0, 1 << 59, (min(polys, 8), LIMBS, N), dtype=np.uint64) % qv[None, :, None] for _ in range(3)]
    x, p, acc = base
    reps = (polys + 7) // 8
    for r in range(reps):
        cnt = min(8, polys - 8 * r)
        eng.dma_mem_h2d(8 * r * per_poly, x[:cnt].reshape(-1))
        eng.dma_mem_h2d(polys * per_poly + 2 * 8 * r * per_poly, np.concatenate([p[:cnt], acc[:cnt]], axis=1).reshape(-1))
    out_row = 3 * polys * per_poly
    steps = 5 if quick else 10
    res = {"vaut": {}, "rotate_mac": {}}
    ok = True
    for name, k in galois_elements():
        if only_k and name != only_k:
            continue
        mac_calls = A.Engine.make_args([(b * per_poly, polys * per_poly + 2 * b * per_poly, out_row + b * per_poly, 0, k)
                                        for b in range(polys)])
        aut_calls = A.Engine.make_args([(b * per_poly, 0, out_row + b * per_poly, 0, k) for b in range(polys)])
        for kind, pc, calls, bytes_per in (("vaut", 2048, aut_calls, 2 * N * 8), ("rotate_mac", 0, mac_calls, 4 * N * 8)):
            for _ in range(3):
                eng.run_vp_batch(pc, calls)
            torch.cuda.synchronize()
            s0 = eng.stats()
            ms = timed(lambda: eng.run_vp_batch(pc, calls), steps)
            s1 = eng.stats()
            per_s = LIMBS * polys * steps / (ms / 1e3)
            # one limb-polynomial of the last polynomial against the oracle
            b, l = polys - 1, (k >> 3) % LIMBS
            got = eng.dma_mem_d2h(out_row + b * per_poly + l * ROWS_PER_POLY, N)
            if kind == "vaut":
                want = O.automorph(x[b % 8, l], k, primes[l])
            else:
                want = O.aut_mac_batch(acc[b % 8, l:l + 1].copy(), x[b % 8, l:l + 1], p[b % 8, l:l + 1], k, qv[l:l + 1],
                                       np.zeros(1, dtype=np.uint32))[0]
            good = bool((got == want).all())
            ok = ok and good
            res[kind][name] = {"k": k, "value": per_s, "ms_per_step": ms / steps,
                               "achieved_gbs_algorithmic": per_s * bytes_per / 1e9,
                               "frac_of_hbm_peak": per_s * bytes_per / 1e9 / peak,
                               "launches_per_step": (s1["kernel_launches"] - s0["kernel_launches"]) / steps,
                               "checked_against_oracle": good}
    eng.close()
    out = {"polys": polys, "limbs": LIMBS, "n": N,
           "kernels": "8-byte gather for both (ALOHA_F_AUT_GATHER)" if flags & A.F_AUT_GATHER else
                      "shared-memory tile permutation for both (ALOHA_F_AUT_TILED)" if flags & A.F_AUT_TILED else
                      "vaut: shared-memory tile permutation (aut_plan.hpp); rotate_mac: fused gather-multiply-add",
           "all_checked_against_oracle": ok}
    for kind, bytes_per, unit in (("vaut", 2 * N * 8, "limb automorphisms/s"), ("rotate_mac", 4 * N * 8, "limb rotate-MACs/s")):
        vals = [v["value"] for v in res[kind].values()]
        worst = min(vals)
        out[kind] = {"value": worst, "unit": unit, "value_is": "the slowest Galois element", "mean": float(np.mean(vals)),
                     "algorithmic_bytes_per_unit": bytes_per,
                     "roofline": {"bound": "hbm", "achieved": worst * bytes_per / 1e9, "peak": peak, "unit": "GB/s",
                                  "frac": worst * bytes_per / 1e9 / peak},
                     "per_galois_element": res[kind]}
    return out


KS_SHAPES = [
    # name, L, K, dnum, batch
    ("digit_1_limb", 47, 1, 47, 1),      # the reference kernel's shape (one limb per digit, K = 1) at 48 limbs
    ("dnum5_k8", 40, 8, 5, 1),           # a hybrid shape real schemes use: 5 digits of 8 limbs, 8 special primes
    ("dnum5_k8_batch8", 40, 8, 5, 8),    # the same, eight key-switches side by side (launches fill the GPUs)
]


def measure_keyswitch(torch, dist, A, eng_kwargs, stream, timed, world, rank, shapes=None):
    """BASELINE.json configs[4]: limb-sharded key-switch streams, N = 2^16, 48 limbs, generated by
    aloha_b200.hks; NCCL inside the C library (aloha_group_*), transfers overlapped with phase 2.
    Every rank checks output limbs of its own against the oracle (>= 4 limbs over the whole job)."""
    from aloha_b200 import hks, params
    from oracle import oracle as O
    peak, _ = peaks()
    out = {}
    grp = None
    for name, L, K, dnum, batch in KS_SHAPES:
        if shapes and name not in shapes:
            continue
        primes = params.synthetic_primes(L + K, 2 * N)
        p, q = primes[:K], primes[K:]
        psi = {m: params.min_primitive_root(m, 2 * N) for m in primes}
        prm = hks.Params(N, q, p, dnum)
        lay = hks.Layout(prm, world, rank, batch)
        eng = A.Engine(vlmax_bits=N * 64, spm_rows=lay.spm_rows, ksk_rows=max(lay.ksk_rows, 1),
                       moduli=[(m, psi[m]) for m in prm.moduli], pool_buffers=min(32768, max(2048, 4 * batch * lay.per_rank * (prm.dnum * (4 if prm.alpha > 1 else 2) + 8))),
                       isram_depth=1 << 17,
                       flags=int(os.environ.get("ALOHA_BENCH_KS_FLAGS", "0")), **eng_kwargs)
        eng.set_stream(stream.cuda_stream)
        if world > 1:
            uid = [A.Group.unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            grp = A.Group.create(eng, uid[0], rank, world)
            comm = hks.GroupComm(grp)
        else:
            comm = hks.LocalComm()
        # phase 2 under the all-gather: "own" = the digits this rank produced run first, without waiting
        overlap = {"none": False, "own": "own", "chunks": "chunks"}[os.environ.get("ALOHA_BENCH_KS_OVERLAP", "own")] if world > 1 else False
        ks = hks.KeySwitch(eng, lay, comm, overlap=overlap)

        def limb_inputs(i, b=0):      # any rank can regenerate any limb's inputs
            rng = np.random.default_rng(1000 + 64 * b + i)
            return rng.integers(0, prm.q[i], N, dtype=np.uint64), rng.integers(0, prm.q[i], N, dtype=np.uint64)

        def modulus_key(t):
            rng = np.random.default_rng(5000 + t)
            return rng.integers(0, prm.moduli[t], 2 * prm.dnum * N, dtype=np.uint64)
        for t in lay.owned():
            if t < L:
                for b in range(batch):
                    ks.load_input(t, limb_inputs(t, b), b)
            ks.load_ksk(t, modulus_key(t))
        k = pow(3, 18, 2 * N) & (N - 1)            # a pseudo-random permutation (vlmax = N keeps log2 N bits of k)
        prog = ks.program(k)
        for _ in range(2):
            ks.execute(prog)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(4):
            ks.execute(prog)
        torch.cuda.synchronize()
        extra = torch.tensor([int(0.3 / max((time.perf_counter() - t0) / 4, 1e-4))], device="cuda")
        if world > 1:                              # same count on every rank (see the main warm-up)
            dist.all_reduce(extra, op=dist.ReduceOp.MAX)
        for _ in range(int(extra.item())):
            ks.execute(prog)
        s0 = eng.stats()
        steps = 20
        ms = timed(lambda: ks.execute(prog), steps)
        s1 = eng.stats()
        # the same stream with the transfers left out (results are meaningless, the kernels are the same):
        # what is left of the transfers after overlap = ms - ms_compute
        ms_compute = None
        if world > 1:
            compute_only = [op for op in prog if op[0] == "run"]
            ks.execute(compute_only)
            ms_compute = timed(lambda: ks.execute(compute_only), steps) / steps
        # per-phase device time of one key-switch (each phase timed alone, transfers excluded).  EVERY rank makes
        # exactly three timed() calls -- timed() holds barriers and an all-reduce -- whatever it owns: a rank
        # holding only special primes has nothing to run in phases 1 and 3, and under the overlap modes the
        # number of phase-2 chunks differs from rank to rank.
        run_ops = [op for op in prog if op[0] == "run"]
        assert len(run_ops) >= 3
        phase_ms = []
        for part in (run_ops[:1], run_ops[1:-1], run_ops[-1:]):
            ks.execute(part)
            phase_ms.append(timed(lambda part=part: ks.execute(part), 10) / 10)
        # parity at full size: output limbs of this rank against the oracle
        ks.execute(prog)
        mine = [i for i in lay.owned() if i < L]
        check = mine[:: max(1, len(mine) // max(1, -(-4 // world)))][: max(1, -(-4 // world))]
        ok = True
        if check:
            lay1 = hks.Layout(prm)
            om = O.GoldenModel(vlmax_bits=N * 64, spm_rows=lay1.spm_rows, ksk_rows=lay1.ksk_rows, moduli=[(m, psi[m]) for m in prm.moduli])
            oks = hks.KeySwitch(om, lay1)
            for i in range(L):
                oks.load_input(i, limb_inputs(i, batch - 1))
            for t in set(check) | set(range(L, L + K)):
                oks.load_ksk(t, modulus_key(t))
            oks.run(k, only=check)
            for i in check:
                want, got = oks.read_output(i), ks.read_output(i, batch - 1)
                ok = ok and bool((want[0] == got[0]).all() and (want[1] == got[1]).all())
            del oks, om
        flag = torch.tensor([1 if ok else 0, len(check)], device="cuda")
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.SUM)
        all_ok = bool(flag[0].item() == world)
        tcount = prm.transform_count()
        alg_bytes = (2 * prm.dnum * (L + K) + 2 * tcount + L) * N * 8 * batch - (batch - 1) * 2 * prm.dnum * (L + K) * N * 8
        per_s = batch * steps / (ms / 1e3)
        res = {"value": per_s, "unit": "key-switches/s", "ms_per_step": ms / steps, "ms_per_keyswitch": ms / steps / batch,
               "limbs": L + K, "special_primes": K, "dnum": prm.dnum, "alpha": prm.alpha, "batch": batch,
               "limb_ntts_per_keyswitch": tcount, "limb_ntts_per_s": tcount * per_s,
               "phase_ms": phase_ms, "launches_per_step_per_rank": (s1["kernel_launches"] - s0["kernel_launches"]) / steps,
               "plans_built_in_timed_region": s1["plans_built"] - s0["plans_built"],
               "ops_fused_per_step": (s1["ops_fused"] - s0["ops_fused"]) / steps,
               "transfers": None if world == 1 else {
                   "backend": "NCCL inside libaloha_b200.so (aloha_group_*), communication stream per GPU",
                   "all_gather_bytes": lay.slots * N * 8 * batch, "broadcast_bytes": 2 * K * N * 8 * batch,
                   "overlap_mode": overlap or "none", "ms_compute_only": ms_compute,
                   "ms_exposed": ms / steps - ms_compute},
               "output_limbs_checked_against_oracle": int(flag[1].item()), "checked_against_oracle": all_ok,
               "roofline": {"bound": "hbm", "unit": "GB/s", "peak": peak,
                            "algorithmic_bytes_per_step": alg_bytes,
                            "algorithmic_bytes_are": "the key (2 dnum (L+K) polynomials, read once per step) + 2 N 8 per limb-(I)NTT + the gathered digits",
                            "achieved": alg_bytes / world / (ms / steps / 1e3) / 1e9,
                            "frac": alg_bytes / world / (ms / steps / 1e3) / 1e9 / peak}}
        out[name] = res
        if grp is not None:
            grp.close()
            grp = None
        eng.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--graphs", action="store_true", help="replay the step as one CUDA graph (ALOHA_F_GRAPHS): takes the host's enqueue cost out")
    ap.add_argument("--generic", action="store_true", help="force the any-prime (Shoup) arithmetic (ALOHA_F_GENERIC_MODMUL)")
    ap.add_argument("--quick", action="store_true", help="device-resident timing only (profiling runs)")
    ap.add_argument("--chunk-mib", type=int, default=0, help="override the engine's L2 chunk size")
    ap.add_argument("--polys", type=int, default=64)
    ap.add_argument("--no-extra", action="store_true", help="skip the rotate-MAC and key-switch workloads")
    ap.add_argument("--shapes", default="", help="with --only keyswitch: comma-separated names from KS_SHAPES")
    ap.add_argument("--galois", default="", help="with --only rotmac: one Galois element by name, e.g. 3^18")
    ap.add_argument("--only", default="", choices=["", "keyswitch", "rotmac", "rotmac_gather", "rotmac_tiled", "tv"], help="profiling: run one extra workload alone")
    args = ap.parse_args()
    globals()["POLYS"] = args.polys
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
