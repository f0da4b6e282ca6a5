#!/bin/bash
# Round-1 GPU evidence run (one B200): parity tests, the bench line, the ncu launch list of the same command
# and one `ncu --set full` capture per dominant kernel.  Outputs land in gpurun_out/ and are summarised
# into profiles/ by profiles/scripts/summarise_ncu.py.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" 
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
B="python bench.py --steps 1 --warmup 1 --davidson 0 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_r01.csv $B > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
# phase ids: 1 xc_gemm, 2 xc_stream, 3 k1, 4 k2 (include/xtd_sigma.h XTD_T_*)
for ph in 4:k2:1 3:k1:1 2:xc_stream:1 1:xc_gemm:3; do
  IFS=: read id name cnt <<< "$ph"
  XTD_PROFILE_PHASE=$id timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -c $cnt -f -o gpurun_out/${name}_full_r01 $B > gpurun_out/ncu_${name}.log 2>&1; echo "ncu $name rc=$?"
done
ls -la gpurun_out
