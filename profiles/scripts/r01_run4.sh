set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu4.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu4.log
for c in 5 3 4 2 1; do python bench.py --config $c --davidson 0 --no-cpu-baseline > gpurun_out/bench4_cfg$c.json 2> gpurun_out/bench4_cfg$c.err; echo "cfg$c rc=$?"; done
