set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/c1_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c1_pytest.log
timeout 300 python bench.py --only rotmac --polys 64 > gpurun_out/c1_rotmac_tiled.json 2> gpurun_out/c1_rotmac_tiled.err
timeout 300 python bench.py --only rotmac_gather --polys 64 > gpurun_out/c1_rotmac_gather.json 2> gpurun_out/c1_rotmac_gather.err
timeout 200 python bench.py --quick --steps 10 --warmup 5 --no-extra > gpurun_out/c1_ntt_quick.json 2> gpurun_out/c1_ntt_quick.err
timeout 120 python bench.py --only rotmac --galois '3^18' --polys 16 --quick > gpurun_out/c1_p1.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tiled -c 12 -o gpurun_out/r2_aut_tiled python bench.py --only rotmac --galois '3^18' --polys 16 --quick > gpurun_out/c1_ncu1.log 2>&1
timeout 120 python bench.py --only rotmac_gather --galois '3^18' --polys 16 --quick > gpurun_out/c1_p2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'vaut_kernel|autmac_kernel' -c 12 -o gpurun_out/r2_aut_gather python bench.py --only rotmac_gather --galois '3^18' --polys 16 --quick > gpurun_out/c1_ncu2.log 2>&1
echo finished
