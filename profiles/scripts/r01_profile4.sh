#!/bin/bash
# ncu full captures of the grid GEMMs of config 4 in the split-gradient form (forward <1,1,1>, <1,0,1>; backward <0,0,1>)
set -u
mkdir -p gpurun_out
B="python bench.py --config 4 --steps 1 --warmup 1 --davidson 0 --no-cpu-baseline"
XTD_PROFILE_PHASE=1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -c 8 -f -o gpurun_out/xc_gemm_cfg4_full_r01 $B > gpurun_out/ncu_xc_gemm_cfg4.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/xc_gemm_cfg4_full_r01.ncu-rep
