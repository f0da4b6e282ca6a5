#!/usr/bin/env python
"""Davidson iteration counts of the synthetic config-5 problem for generator variants (GPU; exploratory)."""
import itertools, json, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from xtddft_b200.davidson import davidson_for_engine
from xtddft_b200.synth_device import make_device_problem
from xtddft_b200.workloads import default_workspace_bytes, engine_for_device_problem

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.3
variants = [
    dict(spectrum="uniform"),
    dict(spectrum="molecular"),
    dict(spectrum="molecular", open_shift=0.1),
    dict(spectrum="molecular", coupling=0.2),
    dict(spectrum="molecular", coupling=0.2, open_shift=0.1, fock_noise=0.005),
    dict(spectrum="molecular", coupling=0.1, open_shift=0.1, fock_noise=0.005, xc_scale=0.4),
]
for v in variants:
    dp = make_device_problem(5, scale, **v)
    eng = engine_for_device_problem(dp, max_nvec=16, workspace_bytes=min(default_workspace_bytes(dp), 8 << 30))
    t0 = time.perf_counter()
    try:
        conv, e, x, info = davidson_for_engine(eng, 10, "sf_down", max_cycle=400)
        torch.cuda.synchronize()
        hd = np.sort(eng.hdiag())[:4]
        print(json.dumps(dict(variant=v, scale=scale, dim=eng.ext_dim, cycles=int(info[0]), sigma=int(info[1]), conv=bool(np.all(conv)),
                              e=[round(float(t), 5) for t in e[:5]], hdiag_min=[round(float(t), 5) for t in hd],
                              seconds=round(time.perf_counter() - t0, 2))), flush=True)
    except Exception as ex:
        print(json.dumps(dict(variant=v, error=str(ex))), flush=True)
    eng.close()
    torch.cuda.empty_cache()
