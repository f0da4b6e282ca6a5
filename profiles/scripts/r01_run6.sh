python bench.py --config 4 --davidson 0 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['phase_ms_per_step']['xc_stream'], d['roofline_xc']['frac'])"
