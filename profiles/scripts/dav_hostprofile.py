#!/usr/bin/env python
"""Where the Davidson host time goes (cProfile around the solve; config 5 at a given scale)."""
import cProfile, pstats, sys, os, io, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from xtddft_b200.davidson import davidson_for_engine
from xtddft_b200.synth_device import make_device_problem
from xtddft_b200.workloads import default_workspace_bytes, engine_for_device_problem

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
dp = make_device_problem(5, scale)
eng = engine_for_device_problem(dp, max_nvec=16, workspace_bytes=default_workspace_bytes(dp))
z = torch.randn((10, eng.ext_dim), dtype=torch.float64, device="cuda")
eng.sigma(z); torch.cuda.synchronize()
tm = {}
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
conv, e, x, info = davidson_for_engine(eng, 10, "sf_down", timing=tm)
torch.cuda.synchronize()
pr.disable()
tot = time.perf_counter() - t0
print(json.dumps(dict(scale=scale, dim=eng.ext_dim, total_s=tot, sigma_s=tm.get("sigma_s"), cycles=int(info[0]), nsigma=int(info[1]))))
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(35)
print(s.getvalue())
