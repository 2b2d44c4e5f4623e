set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/c3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c3_pytest.log
timeout 300 python bench.py --only rotmac --polys 64 > gpurun_out/c3_rotmac.json 2> gpurun_out/c3_rotmac.err
ALOHA_LIB_NAME=libaloha_b200_t12.so timeout 300 python bench.py --only rotmac_tiled --polys 64 > gpurun_out/c3_rotmac_t12.json 2> gpurun_out/c3_rotmac_t12.err
timeout 600 python bench.py --only keyswitch > gpurun_out/c3_keyswitch.json 2> gpurun_out/c3_keyswitch.err
timeout 300 python bench.py --only tv > gpurun_out/c3_tv.json 2> gpurun_out/c3_tv.err
timeout 300 python bench.py --only keyswitch --shapes dnum5_k8 > gpurun_out/c3_p2.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_keyswitch_dnum5.csv python bench.py --only keyswitch --shapes dnum5_k8 > gpurun_out/c3_ncu2.log 2>&1
echo finished
