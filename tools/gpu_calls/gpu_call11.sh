mkdir -p gpurun_out
tools/copy_bench > gpurun_out/r2_copy_bench.txt 2>&1
cat gpurun_out/r2_copy_bench.txt
