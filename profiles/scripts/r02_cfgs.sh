#!/bin/bash
# Round-2: bench lines of configurations 1-4 (one B200), default engine policy.
set -u
mkdir -p gpurun_out
for c in 4 3 2 1; do
  timeout 900 python bench.py --config $c --no-cpu-baseline > gpurun_out/r2_cfg$c.json 2> gpurun_out/r2_cfg$c.err; echo "cfg $c rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_cfg$c.json"))
    print($c, round(d["value"],3), round(d["ms_per_step"],3), {k:round(v,2) for k,v in d["phase_ms_per_step"].items()}, d.get("davidson",{}).get("time_to_roots_s"), d["gpu_launches"])
except Exception as e:
    print("cfg $c failed", e)
PY
done
