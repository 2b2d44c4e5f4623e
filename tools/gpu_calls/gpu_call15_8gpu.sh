set -x
G=8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/c15_topo.txt 2>&1
timeout 1500 $TR --master-port 29810 bench.py --gpus $G > gpurun_out/c15_bench_${G}gpu.json 2> gpurun_out/c15_bench_${G}gpu.err
ALOHA_BENCH_KS_OVERLAP=none timeout 600 $TR --master-port 29811 bench.py --gpus $G --only keyswitch > gpurun_out/c15_ks_${G}gpu_none.json 2> gpurun_out/c15_ks_${G}gpu_none.err
ALOHA_BENCH_KS_OVERLAP=own ALOHA_BENCH_KS_FLAGS=4 timeout 600 $TR --master-port 29812 bench.py --gpus $G --only keyswitch > gpurun_out/c15_ks_${G}gpu_own_graphs.json 2> gpurun_out/c15_ks_${G}gpu_own_graphs.err
timeout 900 $TR --master-port 29813 bench.py --gpus $G --graphs --quick --steps 20 --warmup 5 --no-extra > gpurun_out/c15_quick_graphs_${G}gpu.json 2> gpurun_out/c15_quick_graphs_${G}gpu.err
echo finished
