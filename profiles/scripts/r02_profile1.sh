#!/bin/bash
# Round-2 GPU evidence run (one B200): parity tests, the bench line of config 5 with the emulated exchange / grid GEMMs,
# the ncu launch list of the same command and `ncu --set full` captures of the INT8 kernel (K2 phase) and the slicing kernels.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_r2_02.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_r2_02.log
python bench.py --no-cpu-baseline > gpurun_out/bench_r2_cfg5.json 2> gpurun_out/bench_r2_cfg5.err; echo "bench rc=$?"
B="python bench.py --steps 1 --warmup 1 --davidson 0 --no-cpu-baseline"
$B > gpurun_out/plain_r2.json 2> gpurun_out/plain_r2.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_cfg5_r02.csv $B > gpurun_out/ncu_launch_r02.log 2>&1; echo "ncu launches rc=$?"
# phase ids (include/xtd_sigma.h XTD_T_*): 4 k2 (oz_gemm_kernel + reduce), 9 k2_slice
XTD_PROFILE_PHASE=4 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:oz_gemm -c 1 -f -o gpurun_out/k2_int8_full_r02 $B > gpurun_out/ncu_k2_int8.log 2>&1; echo "ncu k2 rc=$?"
XTD_PROFILE_PHASE=9 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -c 2 -f -o gpurun_out/k2_slice_full_r02 $B > gpurun_out/ncu_k2_slice.log 2>&1; echo "ncu slice rc=$?"
ls -la gpurun_out | tail -12
