set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sigma.py tests/test_gpu_drivers.py tests/test_gpu_golden.py -x -q > gpurun_out/pytest_gpu13.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu13.log
for cfg in 1 2; do
for g in 0 x; do
if [ $g = 0 ]; then export XTD_GRAPH=0; else unset XTD_GRAPH; fi
timeout 300 python bench.py --config $cfg --no-cpu-baseline --steps 20 --warmup 5 2>gpurun_out/bench13_cfg$cfg.err | tee gpurun_out/bench13_cfg${cfg}_g$g.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print($cfg, '$g', d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], d.get('davidson'))"
done; done
