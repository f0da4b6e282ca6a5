set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_sigma.py tests/test_gpu_golden.py -x -q > gpurun_out/pytest_gpu11.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu11.log
python bench.py --config 3 --davidson 0 --no-cpu-baseline 2>gpurun_out/bench11_cfg3.err | tee gpurun_out/bench11_cfg3.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(3, d['ms_per_step'], d['phase_ms_per_step'], d['roofline']['frac'])"
