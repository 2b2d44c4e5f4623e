set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/c4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c4_pytest.log
ALOHA_AUT_DIRECT=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "vaut or rotate or tv_replay_bit" > gpurun_out/c4_pytest_direct.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c4_pytest_direct.log
timeout 300 python bench.py --only rotmac --polys 64 > gpurun_out/c4_rotmac.json 2> gpurun_out/c4_rotmac.err
ALOHA_AUT_DIRECT=1 timeout 300 python bench.py --only rotmac --polys 64 > gpurun_out/c4_rotmac_direct.json 2> gpurun_out/c4_rotmac_direct.err
timeout 600 python bench.py --only keyswitch > gpurun_out/c4_keyswitch.json 2> gpurun_out/c4_keyswitch.err
timeout 300 python bench.py --only tv > gpurun_out/c4_tv.json 2> gpurun_out/c4_tv.err
ALOHA_AUT_DIRECT=1 timeout 120 python bench.py --only rotmac --galois '3^18' --polys 16 --quick > gpurun_out/c4_p1.log 2>&1 && \
ALOHA_AUT_DIRECT=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:vaut_direct -s 2 -c 2 -o gpurun_out/r2_aut_direct python bench.py --only rotmac --galois '3^18' --polys 16 --quick > gpurun_out/c4_ncu1.log 2>&1
echo finished
