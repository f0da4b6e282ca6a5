#!/bin/bash
# Round-1 ncu evidence (one B200), after the plain bench of the same command exited 0:
#  - launch list of every kernel inside the sigma calls (cudaProfilerStart/Stop around xtd_sigma, XTD_PROFILE_PHASE=8)
#  - one `ncu --set full` capture per dominant kernel (profiler range around the phase)
set -u
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --davidson 0 --no-cpu-baseline"
$B > gpurun_out/prof_plain.json 2> gpurun_out/prof_plain.err && \
XTD_PROFILE_PHASE=8 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 1000 --csv --log-file gpurun_out/launches_r01.csv $B > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
# phase ids: 1 xc_gemm, 2 xc_stream, 3 k1, 4 k2 (include/xtd_sigma.h XTD_T_*)
for ph in 4:k2:1 3:k1:1 2:xc_stream:1 1:xc_gemm:2; do
  IFS=: read id name cnt <<< "$ph"
  XTD_PROFILE_PHASE=$id timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -c $cnt -f -o gpurun_out/${name}_full_r01 $B > gpurun_out/ncu_${name}.log 2>&1; echo "ncu $name rc=$?"
done
ls -la gpurun_out | head -40
