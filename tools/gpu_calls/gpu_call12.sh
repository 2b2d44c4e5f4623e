set -x
mkdir -p gpurun_out
timeout 300 python bench.py --only rotmac_tiled --polys 64 > gpurun_out/c12_rotmac_q4.json 2> gpurun_out/c12_rotmac_q4.err
for v in q1 q2 q8 q16; do
ALOHA_LIB_NAME=libaloha_b200_$v.so timeout 300 python bench.py --only rotmac_tiled --polys 64 > gpurun_out/c12_rotmac_$v.json 2> gpurun_out/c12_rotmac_$v.err
done
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "vaut_every or rotate_mac" > gpurun_out/c12_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c12_pytest.log
timeout 120 python bench.py --only rotmac --galois '3^18' --polys 16 --quick > gpurun_out/c12_p1.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vaut_tiled -s 2 -c 2 -o gpurun_out/r2_aut_tiled_d python bench.py --only rotmac --galois '3^18' --polys 16 --quick > gpurun_out/c12_ncu1.log 2>&1
echo finished
