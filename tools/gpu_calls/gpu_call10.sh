set -x
mkdir -p gpurun_out
ALOHA_BENCH_GALOIS="identity:65537,reversal:131071,three:3,rand:102729" timeout 300 python bench.py --only rotmac_tiled --polys 64 > gpurun_out/c10_rotmac_tiled_special.json 2> gpurun_out/c10_rotmac_tiled_special.err
ALOHA_BENCH_GALOIS="identity:65537,reversal:131071,three:3,rand:102729" timeout 300 python bench.py --only rotmac_gather --polys 64 > gpurun_out/c10_rotmac_gather_special.json 2> gpurun_out/c10_rotmac_gather_special.err
echo finished
