set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu2.log
python bench.py --davidson 0 --no-cpu-baseline > gpurun_out/bench2_n1.json 2> gpurun_out/bench2_n1.err; echo "bench rc=$?"
python profiles/scripts/dav_sweep.py 0.3 > gpurun_out/dav_sweep_03.jsonl 2> gpurun_out/dav_sweep.err; echo "sweep rc=$?"
for c in 3 4; do python bench.py --config $c --davidson 0 --no-cpu-baseline > gpurun_out/bench2_cfg$c.json 2> gpurun_out/bench2_cfg$c.err; echo "cfg$c rc=$?"; done
