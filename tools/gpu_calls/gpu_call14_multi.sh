# usage: bash tools/gpu_call14_multi.sh <ngpus> [full]
set -x
G=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
p=29700
for mode in none own chunks; do
p=$((p+1))
ALOHA_BENCH_KS_OVERLAP=$mode timeout 600 $TR --master-port $p bench.py --gpus $G --only keyswitch > gpurun_out/c14_ks_${G}gpu_$mode.json 2> gpurun_out/c14_ks_${G}gpu_$mode.err
done
if [ "$2" = "full" ]; then
timeout 1500 $TR --master-port 29710 bench.py --gpus $G > gpurun_out/c14_bench_${G}gpu.json 2> gpurun_out/c14_bench_${G}gpu.err
fi
echo finished
