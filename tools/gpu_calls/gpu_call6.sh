set -x
mkdir -p gpurun_out
timeout 300 python bench.py --only rotmac_tiled --polys 64 > gpurun_out/c6_rotmac_va.json 2> gpurun_out/c6_rotmac_va.err
for v in b c d e; do
ALOHA_LIB_NAME=libaloha_b200_v$v.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "vaut_every or rotate_mac" > gpurun_out/c6_pytest_v$v.log 2>&1; echo "rc=$?" >> gpurun_out/c6_pytest_v$v.log
ALOHA_LIB_NAME=libaloha_b200_v$v.so timeout 300 python bench.py --only rotmac_tiled --polys 64 > gpurun_out/c6_rotmac_v$v.json 2> gpurun_out/c6_rotmac_v$v.err
done
timeout 300 python bench.py --only tv > gpurun_out/c6_tv.json 2> gpurun_out/c6_tv.err
echo finished
