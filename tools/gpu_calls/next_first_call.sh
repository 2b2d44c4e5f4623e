#!/bin/bash
# First GPU call after round 2: everything below was changed on the host side after the round's last GPU call and has
# only run on the simulated device.  One GPU; nothing here runs under ncu.
#   gpurun --timeout 1500 -- bash tools/gpu_calls/next_first_call.sh
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/n1_pytest.log 2>&1; echo "pytest rc=$?"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/n1_smoke.log 2>&1; echo "smoke rc=$?"
timeout 300 python bench.py --only tv > gpurun_out/n1_tv.json 2> gpurun_out/n1_tv.err            # dumped ops: host time was the limit
timeout 600 python bench.py --only keyswitch > gpurun_out/n1_keyswitch.json 2> gpurun_out/n1_keyswitch.err   # retire streams: 43 -> ~29 launches
timeout 300 python bench.py --only rotmac --polys 64 > gpurun_out/n1_rotmac.json 2> gpurun_out/n1_rotmac.err  # one launch per step
timeout 600 python bench.py > gpurun_out/n1_bench.json 2> gpurun_out/n1_bench.err                 # host_enqueue_ms next to ms_per_step
echo finished
