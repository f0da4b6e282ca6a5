#!/bin/bash
# Round-2 evidence for the kernels changed late in the round (one B200): integer digit extraction in the fused half-transform
# (oz_k1_kernel), the persistent INT8 contraction kernel (oz_gemm_p_kernel: grid GEMMs, short contractions), the small-batch
# split-gradient streaming kernel (xc_weight_split_op_kernel: Z-vector operator).  Plain run first, then the same command under ncu.
set -u
mkdir -p gpurun_out
for c in 5 4; do
  B="python bench.py --config $c --steps 1 --warmup 1 --davidson 0 --no-cpu-baseline --configs-table 0"
  $B > gpurun_out/plain_cfg${c}_r02b.json 2> gpurun_out/plain_cfg${c}_r02b.err && \
  XTD_PROFILE_PHASE=8 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 4000 --csv --log-file gpurun_out/launches_cfg${c}_r02b.csv $B > gpurun_out/ncu_launch_cfg${c}_r02b.log 2>&1; echo "ncu launches cfg$c rc=$?"
done
B="python bench.py --config 5 --steps 1 --warmup 1 --davidson 0 --no-cpu-baseline --configs-table 0"
XTD_PROFILE_PHASE=3 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:oz_k1 -c 1 -f -o gpurun_out/k1_int8_full_r02c $B > gpurun_out/ncu_k1_int8_c.log 2>&1; echo "ncu k1 rc=$?"
XTD_PROFILE_PHASE=1 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:oz_gemm_p -c 1 -f -o gpurun_out/xc_int8p_full_r02c $B > gpurun_out/ncu_xc_int8p_c.log 2>&1; echo "ncu xc persistent rc=$?"
# Z-vector operator at config-4 inputs (one vector per call): plain, then the small-batch streaming kernel under ncu
Z="python -c \"import json, bench; print(json.dumps(bench.zvector_record(4)))\""
eval $Z > gpurun_out/zvector_cfg4_r02.json 2> gpurun_out/zvector_cfg4_r02.err && \
XTD_PROFILE_PHASE=2 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:xc_weight_split_op -c 1 -f -o gpurun_out/xc_split_op_full_r02c python -c "import json, bench; print(json.dumps(bench.zvector_record(4)))" > gpurun_out/ncu_split_op_c.log 2>&1; echo "ncu split op rc=$?"
ls -la gpurun_out | tail -8
