#!/bin/bash
# ncu evidence for config 4 (X-TDA, value+gradient kernels in the split-gradient form): launch list of one sigma call and
# one full capture of the streaming kernel, after the plain run of the same command exited 0.
set -u
mkdir -p gpurun_out
B="python bench.py --config 4 --steps 1 --warmup 1 --davidson 0 --no-cpu-baseline"
$B > gpurun_out/prof_cfg4_plain.json 2> gpurun_out/prof_cfg4_plain.err && \
XTD_PROFILE_PHASE=8 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 1000 --csv --log-file gpurun_out/launches_cfg4_r01.csv $B > gpurun_out/ncu_launch_cfg4.log 2>&1; echo "ncu launches rc=$?"
XTD_PROFILE_PHASE=2 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -c 1 -f -o gpurun_out/xc_split_full_r01 $B > gpurun_out/ncu_xc_split.log 2>&1; echo "ncu xc_split rc=$?"
ls -la gpurun_out | tail -8
