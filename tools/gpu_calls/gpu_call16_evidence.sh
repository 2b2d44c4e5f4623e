set -x
mkdir -p gpurun_out
Q="python bench.py --quick --steps 3 --warmup 3 --no-extra"
timeout 300 $Q > gpurun_out/c16_q.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file gpurun_out/r2_launches_ntt.csv $Q > gpurun_out/c16_ncu_a.log 2>&1
timeout 300 $Q > gpurun_out/c16_q.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:ntt_fwd -s 6 -c 2 -o gpurun_out/r2_ntt_fwd $Q > gpurun_out/c16_ncu_b.log 2>&1
K="python bench.py --only keyswitch --shapes digit_1_limb,dnum5_k8"
timeout 400 $K > gpurun_out/c16_k.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches_keyswitch.csv $K > gpurun_out/c16_ncu_c.log 2>&1
K2="python bench.py --only keyswitch --shapes dnum5_k8"
timeout 400 $K2 > gpurun_out/c16_k2.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:'sop_kernel|bext_kernel|ew_kernel|muladd' -s 20 -c 10 -o gpurun_out/r2_ks_kernels $K2 > gpurun_out/c16_ncu_d.log 2>&1
R="python bench.py --only rotmac --galois 3^18 --polys 16 --quick"
timeout 200 $R > gpurun_out/c16_r.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:'vaut_tiled|autmac_kernel' -s 2 -c 9 -o gpurun_out/r2_aut_final $R > gpurun_out/c16_ncu_e.log 2>&1
RG="python bench.py --only rotmac_gather --galois 3^18 --polys 16 --quick"
timeout 200 $RG > gpurun_out/c16_rg.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:'vaut_kernel' -s 2 -c 1 -o gpurun_out/r2_aut_gather_final $RG > gpurun_out/c16_ncu_f.log 2>&1
echo finished
