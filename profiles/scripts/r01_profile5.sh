#!/bin/bash
# launch list of one config-3 (XSF-TDA) sigma call
set -u
mkdir -p gpurun_out
B="python bench.py --config 3 --steps 1 --warmup 1 --davidson 0 --no-cpu-baseline"
XTD_PROFILE_PHASE=8 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 2000 --csv --log-file gpurun_out/launches_cfg3_r01.csv $B > gpurun_out/ncu_launch_cfg3.log 2>&1; echo "ncu launches rc=$?"
