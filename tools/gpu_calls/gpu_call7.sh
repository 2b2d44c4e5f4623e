set -x
mkdir -p gpurun_out
timeout 300 python bench.py --only rotmac --polys 64 > gpurun_out/c7_rotmac_p32.json 2> gpurun_out/c7_rotmac_p32.err
for v in p0 p8 p96; do
ALOHA_LIB_NAME=libaloha_b200_$v.so timeout 300 python bench.py --only rotmac --polys 64 > gpurun_out/c7_rotmac_$v.json 2> gpurun_out/c7_rotmac_$v.err
done
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "vaut_every or rotate_mac or asynchronous" > gpurun_out/c7_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c7_pytest.log
timeout 300 python bench.py --only tv > gpurun_out/c7_tv.json 2> gpurun_out/c7_tv.err
echo finished
