"""__graft_entry__.smoke() walked on the CPU: the oracle stands in for the engine, so the checks inside smoke()
compare the oracle with itself and say nothing about the CUDA path -- what this covers is smoke()'s own code
(the calls it makes, the argument order, the ROM and scratchpad sizes it asks for), which otherwise runs for the
first time on the GPU box."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

DRIVER = r'''
import sys
sys.path.insert(0, %(root)r)
import aloha_b200 as A
from oracle import oracle as O

class FakeEngine:
    def __init__(self, vlmax_bits=A.VLMAX_BITS, spm_rows=A.SPM_ROWS, ksk_rows=A.KSK_ROWS, device=0, flags=0,
                 moduli=A.REFERENCE_MODULI, pool_buffers=0, l2_chunk_bytes=0, isram_depth=0):
        self.depth = isram_depth or 4096            # aloha_cfg.isram_depth: 0 = 4096
        self.m = O.GoldenModel(vlmax_bits=vlmax_bits, spm_rows=spm_rows, ksk_rows=max(ksk_rows, 1), moduli=list(moduli))
        self.runs = 0
    def load_isram(self, w, pc):
        assert pc + len(w) <= self.depth, "ROM image beyond isram_depth"
        self.m.load_isram(w, pc)
    def dma_ksk_h2d(self, row, data): self.m.dma_ksk_h2d(row, data)
    def dma_mem_h2d(self, row, data): self.m.dma_mem_h2d(row, data)
    def dma_mem_d2h(self, row, nwords): return self.m.dma_mem_d2h(row, nwords)
    def spm_written(self, row, nwords): return self.m.spm_written(row, nwords)
    def run_vp(self, *a): self.runs += 1; self.m.run_vp(*a)
    def run_vp_multi(self, calls): self.runs += 1; self.m.run_vp_multi(calls)
    def stats(self): return {"kernel_launches": 2 * self.runs}
    def close(self): self.m = None

A.Engine = FakeEngine
import __graft_entry__ as g
g.smoke()
'''


def test_smoke_reaches_its_last_line():
    p = subprocess.run([sys.executable, "-c", DRIVER % {"root": ROOT}], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-3000:]
    assert "smoke ok" in p.stdout and "key-switch bit-exact" in p.stdout
