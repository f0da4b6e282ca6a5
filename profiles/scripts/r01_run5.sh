set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_sigma.py -x -q -k "split_gradient or xtda or sf" > gpurun_out/pytest_gpu5.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu5.log
for sp in 1; do
XTD_XC_SPLIT=$sp python bench.py --config 4 --davidson 0 --no-cpu-baseline > gpurun_out/bench5_cfg4_split$sp.json 2> gpurun_out/bench5_cfg4_split$sp.err; echo "cfg4 split=$sp rc=$?"
python - <<P
import json
d=json.loads(open('gpurun_out/bench5_cfg4_split$sp.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['phase_ms_per_step'])
P
done
