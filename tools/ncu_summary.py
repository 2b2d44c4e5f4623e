#!/usr/bin/env python3
"""Summarise an `ncu --set full` report of the transform kernels into the two JSON files bench.py and
DESIGN.md cite:  profiles/<tag>_ncu_full_summary.json (selected metrics + stall ratios + the hottest
SASS lines by warp samples) and profiles/<tag>_traffic.json (DRAM bytes per launch pair).

usage: tools/ncu_summary.py gpurun_out/r1_final_fwd.ncu-rep profiles/r1_ntt_fwd "<capture command>"
       tools/ncu_summary.py <rep> <prefix> "<command>" --kernels      (any kernels: one entry per kernel name, the
                                                                      longest launch; no NTT traffic file)
"""
import csv
import io
import json
import subprocess
import sys

METRICS = [
    "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
    "smsp__warps_eligible.avg.per_cycle_active",
]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True, check=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, prefix, command = sys.argv[1], sys.argv[2], sys.argv[3]
    raw = page(rep, "raw")
    hdr, units, data = raw[0], raw[1], raw[2:]
    kernels = []
    for row in data:
        k = {m: {"value": row[hdr.index(m)], "unit": units[hdr.index(m)]} for m in METRICS if m in hdr}
        stalls = {h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): float(row[i])
                  for i, h in enumerate(hdr)
                  if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and row[i]}
        k["warp_stall_cycles_per_issued_instruction"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:10])
        kernels.append(k)
    # hottest SASS lines (warp-state samples) per kernel
    src = page(rep, "source")
    blocks, cur = [], None
    for row in src:
        if row and row[0] == "Kernel Name":
            cur = {"kernel": row[1], "rows": [], "hdr": None}
            blocks.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = row
        elif cur is not None and row:
            cur["rows"].append(row)
    per = max(1, len(blocks) // max(1, len(kernels)))      # the source page lists every launch once per view
    for k, blk in zip(kernels, blocks[::per]):
        h = blk["hdr"]
        si, so = h.index("# Samples"), h.index("Source")
        total = sum(int(r[si]) for r in blk["rows"])
        top = sorted(blk["rows"], key=lambda r: -int(r[si]))[:8]
        k["hottest_sass_by_warp_samples"] = {"total_samples": total,
                                             "lines": [{"samples": int(r[si]), "sass": " ".join(r[so].split())} for r in top]}
    if "--kernels" in sys.argv:
        best = {}
        for k in kernels:
            name = k["Kernel Name"]["value"]
            if name not in best or float(k["gpu__time_duration.sum"]["value"]) > float(best[name]["gpu__time_duration.sum"]["value"]):
                best[name] = k
        json.dump({"source": command, "kernels": list(best.values())}, open(prefix + "_ncu_full_summary.json", "w"), indent=1)
        print(prefix + "_ncu_full_summary.json", len(best), "kernels")
        return
    json.dump({"source": command, "kernels": kernels}, open(prefix + "_ncu_full_summary.json", "w"), indent=1)
    gb = lambda k, m: int(round(float(k[m]["value"]) * 1e9)) if k[m]["unit"] == "Gbyte" else int(float(k[m]["value"]))
    traffic = {"source": prefix + "_ncu_full_summary.json", "unit": "bytes per launch pair (one bench step: 2048 limb-NTTs, N=65536)"}
    total = 0
    for k in kernels:
        name = k["Kernel Name"]["value"].replace("void ", "").split("(")[0]
        traffic[name] = {"dram_read": gb(k, "dram__bytes_read.sum"), "dram_write": gb(k, "dram__bytes_write.sum")}
        total += traffic[name]["dram_read"] + traffic[name]["dram_write"]
    traffic["total"] = total
    traffic["algorithmic_bytes"] = 2048 * 2 * 65536 * 8
    traffic["note"] = ("two-pass 4-step transform: each pass reads and writes the polynomial once; the row pass also pulls "
                       "its per-row twiddle blocks (from L2 after the first of the 64 polynomials of a modulus)")
    json.dump(traffic, open(prefix.replace("_fwd", "") + "_traffic.json", "w"), indent=1)
    print(json.dumps(traffic))


if __name__ == "__main__":
    main()
