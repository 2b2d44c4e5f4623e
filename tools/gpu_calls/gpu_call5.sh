set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/c5_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c5_pytest.log
timeout 300 python bench.py --only rotmac_tiled --polys 64 > gpurun_out/c5_rotmac_tpc4.json 2> gpurun_out/c5_rotmac_tpc4.err
ALOHA_LIB_NAME=libaloha_b200_tpc2.so timeout 300 python bench.py --only rotmac_tiled --polys 64 > gpurun_out/c5_rotmac_tpc2.json 2> gpurun_out/c5_rotmac_tpc2.err
ALOHA_LIB_NAME=libaloha_b200_tpc8.so timeout 300 python bench.py --only rotmac_tiled --polys 64 > gpurun_out/c5_rotmac_tpc8.json 2> gpurun_out/c5_rotmac_tpc8.err
timeout 300 python bench.py --only tv > gpurun_out/c5_tv.json 2> gpurun_out/c5_tv.err
timeout 120 python bench.py --only rotmac --galois '3^18' --polys 16 --quick > gpurun_out/c5_p1.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vaut_tiled -s 2 -c 2 -o gpurun_out/r2_aut_tiled_c python bench.py --only rotmac --galois '3^18' --polys 16 --quick > gpurun_out/c5_ncu1.log 2>&1
echo finished
