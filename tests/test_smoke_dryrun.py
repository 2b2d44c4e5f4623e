"""__graft_entry__.smoke() on the simulated device (tests/sim_engine.py): smoke()'s own code -- the calls it makes,
the ROM and scratchpad sizes it asks for -- and the plans the batcher builds for it are checked here against the
oracle; the sm_100a kernels it launches on the GPU box are not."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

DRIVER = r'''
import sys
sys.path.insert(0, %(root)r)
sys.path.insert(0, %(root)r + "/tests")
import sim_engine
import __graft_entry__ as g
with sim_engine.simulated():
    g.smoke()
'''


def test_smoke_reaches_its_last_line():
    p = subprocess.run([sys.executable, "-c", DRIVER % {"root": ROOT}], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-3000:]
    assert "smoke ok" in p.stdout and "key-switch bit-exact" in p.stdout
