set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu9.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu9.log
python bench.py > gpurun_out/bench9_cfg5.json 2> gpurun_out/bench9_cfg5.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/bench9_cfg5.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['roofline_xc']['frac'], d.get('davidson'), d.get('cpu_baseline',{}).get('value'))"
python -c "import __graft_entry__ as g; g.smoke()"
