#!/usr/bin/env python
"""Write profiles/r2_sass_excerpts.md: per hot kernel of the built library, the SASS mnemonics that prove what the
design claims (TMA bulk copies, tensor-map copies, mbarrier traffic, cp.async, the multiply-pipe instruction mix),
plus the disassembly of one pseudo-Mersenne butterfly.  Runs on the CPU build box (cuobjdump only)."""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "aloha_b200", "libaloha_b200.so")
HOT = [("ntt_fwd_cols<8, 1>", r"ntt_fwd_colsILi8ELi1E"), ("ntt_fwd_rows_tma<8, 1>", r"ntt_fwd_rows_tmaILi8ELi1E"),
       ("ntt_inv_rows_tma<8, 1>", r"ntt_inv_rows_tmaILi8ELi1E"), ("ntt_inv_cols<8, 1>", r"ntt_inv_colsILi8ELi1E"),
       ("ntt_fwd_cols<8, 0> (any-prime arithmetic)", r"ntt_fwd_colsILi8ELi0E"),
       ("vaut_tiled_kernel", r"vaut_tiled_kernel"), ("autmac_kernel (fused gather-multiply-add)", r"13autmac_kernel"),
       ("vaut_kernel (gather)", r"11vaut_kernel"), ("bext_kernel", r"bext_kernel"), ("sop_kernel", r"sop_kernel")]
WATCH = ["UBLKCP", "UTMALDG", "SYNCS", "LDGSTS", "IMAD.WIDE", "IMAD", "IADD3", "LOP3", "SHF", "SEL", "ISETP", "LDS", "STS",
         "LDG", "STG", "BAR", "ATOMS", "SHFL", "CCTL"]


def functions(sass):
    cur, out = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            out[cur] = []
        elif cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            out[cur].append(re.sub(r"/\* 0x[0-9a-f]+ \*/", "", line).split(None, 1)[1].strip())
    return out


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    fns = functions(sass)
    md = ["# SASS evidence (round 2) -- `cuobjdump -sass aloha_b200/libaloha_b200.so`, sm_100a", "",
          "Counts are static instructions of the kernel body (loops unrolled by ptxas as built).", "",
          "| kernel | instructions | " + " | ".join(WATCH) + " |", "|---|---|" + "---|" * len(WATCH)]
    for label, pat in HOT:
        hits = [k for k in fns if re.search(pat, k)]
        if not hits:
            continue
        body = fns[hits[0]]
        ops = [re.sub(r"^@!?U?P\d+\s+", "", l).split()[0].rstrip(";") for l in body]
        cnt = collections.Counter()
        for o in ops:
            for w in WATCH:
                if o == w or o.startswith(w + "."):
                    cnt[w] += 1
        md.append(f"| `{label}` | {len(body)} | " + " | ".join(str(cnt[w]) if cnt[w] else "" for w in WATCH) + " |")
    md += ["", "(`IMAD` includes `IMAD.WIDE`; `UBLKCP` = `cp.async.bulk`, `UTMALDG` = `cp.async.bulk.tensor`, `SYNCS` = mbarrier "
           "operations, `LDGSTS` = `cp.async`.)", ""]
    # one butterfly
    src = r'''
#include "%s/aloha_b200/csrc/modarith.cuh"
using namespace alb;
__global__ void one(u64 *io, const ulonglong2 *tw, u32 d2, u64 q3) {
    u64 x = io[threadIdx.x], y = io[threadIdx.x + 256];
    const ulonglong2 t = tw[threadIdx.x & 7];
    u64 P, L;
    mul_pm_parts(y, t.x, t.y, d2, P, L);
    const u64 yn = (x + q3 - P) - L;
    x = x + P + L;
    io[threadIdx.x] = x; io[threadIdx.x + 256] = yn;
}''' % ROOT
    with tempfile.TemporaryDirectory() as d:
        cu, cubin = os.path.join(d, "one.cu"), os.path.join(d, "one.cubin")
        open(cu, "w").write(src)
        subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-cubin", "-o", cubin, cu], check=True)
        one = functions(subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True, check=True).stdout)
    body = list(one.values())[0]
    start = next(i for i, l in enumerate(body) if l.startswith("IMAD.WIDE.U32") and "RZ" in l)
    end = max(i for i, l in enumerate(body) if l.startswith("IADD3.X"))
    md += ["## One pseudo-Mersenne Cooley-Tukey butterfly (`modarith.cuh` `mul_pm_parts` + the two outputs)", "", "```"]
    md += body[start:end + 1] + ["```", "",
          f"{end + 1 - start} instructions: 5 `IMAD.WIDE.U32` (four partial products, two carrying a 64-bit addend and a carry-out "
          "predicate; one fold) + `IMAD.X` on the multiply pipe, 3 `IADD3` + 2 `IADD3.X` + `LOP3` + `SHF` on the ALU pipe.", ""]
    out = os.path.join(ROOT, "profiles", "r2_sass_excerpts.md")
    open(out, "w").write("\n".join(md))
    print(out)


if __name__ == "__main__":
    sys.exit(main())
