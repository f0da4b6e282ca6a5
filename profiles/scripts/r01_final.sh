set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_final.log
python bench.py > gpurun_out/final_cfg5.json 2> gpurun_out/final_cfg5.err; echo "bench cfg5 rc=$?"
for cfg in 4 3 2 1; do
python bench.py --config $cfg --no-cpu-baseline > gpurun_out/final_cfg$cfg.json 2> gpurun_out/final_cfg$cfg.err; echo "bench cfg$cfg rc=$?"
done
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; echo "ref rc=$?"
python - <<P
import json
for c in (5,4,3,2,1):
    d=json.loads(open(f'gpurun_out/final_cfg{c}.json').read().strip().splitlines()[-1])
    print(c, round(d['value'],3), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],3), 'roof', round(d['roofline']['frac'] or 0,3), 'xc', d['roofline_xc'] and round(d['roofline_xc']['frac'],3), 'dav', d.get('davidson') and (round(d['davidson']['time_to_roots_s'],2), d['davidson']['converged'], d['davidson']['sigma_vectors']), 'launches', d['gpu_launches'])
d=json.loads(open('gpurun_out/final_ref.json').read().strip().splitlines()[-1]); print('ref', d['value'], d['cpu_baseline']['cores'])
P
