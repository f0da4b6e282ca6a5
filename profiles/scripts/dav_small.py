#!/usr/bin/env python
"""Davidson time-to-roots of the small configurations (1, 2): three solves in one process (the first pays one-off costs such as
CUDA module loading and graph capture), cProfile of the last."""
import cProfile, pstats, sys, os, io, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from xtddft_b200.davidson import davidson_for_engine
from xtddft_b200.synth_device import make_device_problem
from xtddft_b200.workloads import default_workspace_bytes, engine_for_device_problem

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dp = make_device_problem(cfg, 1.0)
eng = engine_for_device_problem(dp, max_nvec=16, workspace_bytes=default_workspace_bytes(dp))
z = torch.randn((dp.nroots, eng.ext_dim), dtype=torch.float64, device="cuda")
for _ in range(3):
    eng.sigma(z)
torch.cuda.synchronize()
for rep in range(3):
    tm = {}
    pr = cProfile.Profile()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if rep == 2:
        pr.enable()
    conv, e, x, info = davidson_for_engine(eng, dp.nroots, dp.method, timing=tm if rep == 1 else None)
    torch.cuda.synchronize()
    if rep == 2:
        pr.disable()
    tot = time.perf_counter() - t0
    print(json.dumps(dict(cfg=cfg, rep=rep, dim=eng.ext_dim, total_ms=tot * 1e3, sigma_ms=(tm.get("sigma_s") or 0) * 1e3, cycles=int(info[0]),
                          nsigma=int(info[1]), e0=float(e[0]))))
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22)
print(s.getvalue())
