set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_gemm.py tests/test_gpu_sigma.py -x -q > gpurun_out/pytest_gpu12.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu12.log
for cfg in 4 5; do
python bench.py --config $cfg --davidson 0 --no-cpu-baseline 2>gpurun_out/bench12_cfg$cfg.err | tee gpurun_out/bench12_cfg$cfg.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print($cfg, d['ms_per_step'], d['phase_ms_per_step'], d['roofline']['frac'])"
done
