set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_sigma.py tests/test_gpu_golden.py -x -q > gpurun_out/pytest_gpu7.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu7.log
for nn in 1 0; do
if [ $nn = 1 ]; then export XTD_NO_NARROW=1; else unset XTD_NO_NARROW; fi
python bench.py --config 3 --davidson 0 --no-cpu-baseline 2>gpurun_out/bench7_cfg3.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['phase_ms_per_step'])"
done
