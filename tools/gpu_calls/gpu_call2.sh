set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/c2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c2_pytest.log
timeout 300 python bench.py --only rotmac --polys 64 > gpurun_out/c2_rotmac_tiled.json 2> gpurun_out/c2_rotmac_tiled.err
timeout 600 python bench.py --only keyswitch > gpurun_out/c2_keyswitch.json 2> gpurun_out/c2_keyswitch.err
timeout 120 python bench.py --only rotmac --galois '3^18' --polys 16 --quick > gpurun_out/c2_p1.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tiled -s 2 -c 8 -o gpurun_out/r2_aut_tiled_b python bench.py --only rotmac --galois '3^18' --polys 16 --quick > gpurun_out/c2_ncu1.log 2>&1
timeout 300 python bench.py --only keyswitch --shapes dnum5_k8 > gpurun_out/c2_p2.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_keyswitch_dnum5.csv python bench.py --only keyswitch --shapes dnum5_k8 > gpurun_out/c2_ncu2.log 2>&1
echo finished
