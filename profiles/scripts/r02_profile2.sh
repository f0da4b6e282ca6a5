#!/bin/bash
# Round-2 final GPU evidence (one B200): bench record with the configs table, ncu launch lists of the config 5 / 4 / 3 steps and
# `ncu --set full` captures of the INT8 kernels (exchange contraction, fused half-transform, grid GEMMs) and the slicing kernel.
set -u
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_cfg5_n1_r02.json 2> gpurun_out/bench_cfg5_n1_r02.err; echo "bench rc=$?"
for c in 5 4 3; do
  B="python bench.py --config $c --steps 1 --warmup 1 --davidson 0 --no-cpu-baseline --configs-table 0"
  $B > gpurun_out/plain_cfg${c}_r02.json 2> gpurun_out/plain_cfg${c}_r02.err && \
  XTD_PROFILE_PHASE=8 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 4000 --csv --log-file gpurun_out/launches_cfg${c}_r02.csv $B > gpurun_out/ncu_launch_cfg${c}_r02.log 2>&1; echo "ncu launches cfg$c rc=$?"
done
B="python bench.py --config 5 --steps 1 --warmup 1 --davidson 0 --no-cpu-baseline --configs-table 0"
# phase ids (include/xtd_sigma.h XTD_T_*): 4 k2 (oz_gemm_kernel + reduce), 3 k1 (bound scale + oz_k1_kernel), 1 xc_gemm, 10 xc_slice
XTD_PROFILE_PHASE=4 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:oz_gemm -c 1 -f -o gpurun_out/k2_int8_full_r02 $B > gpurun_out/ncu_k2_int8.log 2>&1; echo "ncu k2 rc=$?"
XTD_PROFILE_PHASE=3 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:oz_k1 -c 1 -f -o gpurun_out/k1_int8_full_r02 $B > gpurun_out/ncu_k1_int8.log 2>&1; echo "ncu k1 rc=$?"
XTD_PROFILE_PHASE=1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:oz_gemm -c 2 -f -o gpurun_out/xc_int8_full_r02 $B > gpurun_out/ncu_xc_int8.log 2>&1; echo "ncu xc rc=$?"
XTD_PROFILE_PHASE=10 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -s 2 -c 2 -f -o gpurun_out/xc_slice_full_r02 $B > gpurun_out/ncu_xc_slice.log 2>&1; echo "ncu xc slice rc=$?"
ls -la gpurun_out | tail -8
