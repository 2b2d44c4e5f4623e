set -x
mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/c13_bench_1gpu.json 2> gpurun_out/c13_bench_1gpu.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/c13_bench_ref.json 2> gpurun_out/c13_bench_ref.err
echo finished
