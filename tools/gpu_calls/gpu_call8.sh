set -x
mkdir -p gpurun_out
timeout 300 python bench.py --only rotmac --polys 64 > gpurun_out/c8_rotmac_f3.json 2> gpurun_out/c8_rotmac_f3.err
for v in f7 f8 f9; do
ALOHA_LIB_NAME=libaloha_b200_$v.so timeout 300 python bench.py --only rotmac --polys 64 > gpurun_out/c8_rotmac_$v.json 2> gpurun_out/c8_rotmac_$v.err
done
echo finished
