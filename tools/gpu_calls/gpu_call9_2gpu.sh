set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/c9_topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_hks.py -m gpu -x -q > gpurun_out/c9_pytest_hks.log 2>&1; echo "rc=$?" >> gpurun_out/c9_pytest_hks.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29611 bench.py --gpus 2 --only keyswitch > gpurun_out/c9_ks_2gpu.json 2> gpurun_out/c9_ks_2gpu.err
ALOHA_BENCH_KS_NO_OVERLAP=1 timeout 900 $TR --master-port 29612 bench.py --gpus 2 --only keyswitch > gpurun_out/c9_ks_2gpu_nooverlap.json 2> gpurun_out/c9_ks_2gpu_nooverlap.err
timeout 1200 $TR --master-port 29613 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/c9_bench_2gpu.json 2> gpurun_out/c9_bench_2gpu.err
echo finished
