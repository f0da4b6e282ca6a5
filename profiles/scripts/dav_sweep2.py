#!/usr/bin/env python
"""Davidson iteration counts of the synthetic config-3 (XSF-TDA) and config-4 (X-TDA) problems at reduced scale, for the
generator's ROHF-form Fock difference (GPU; exploratory)."""
import json, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from xtddft_b200.davidson import davidson_for_engine
from xtddft_b200.synth_device import make_device_problem
from xtddft_b200.workloads import default_workspace_bytes, engine_for_device_problem

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.3
for cfg in (4, 3):
    for v in (dict(), dict(hf_exchange=0.05), dict(xc_scale=0.5), dict(coupling=0.1)):
        dp = make_device_problem(cfg, scale, **v)
        eng = engine_for_device_problem(dp, max_nvec=40, workspace_bytes=min(default_workspace_bytes(dp), 8 << 30))
        t0 = time.perf_counter()
        try:
            conv, e, x, info = davidson_for_engine(eng, dp.nroots, dp.method)
            torch.cuda.synchronize()
            hd = np.sort(eng.hdiag())[:4]
            print(json.dumps(dict(cfg=cfg, variant=v, scale=scale, dim=eng.ext_dim, cycles=int(info[0]), sigma=int(info[1]), conv=bool(np.all(conv)),
                                  e=[round(float(t), 5) for t in e[:5]], hdiag_min=[round(float(t), 5) for t in hd],
                                  seconds=round(time.perf_counter() - t0, 2))), flush=True)
        except Exception as ex:
            print(json.dumps(dict(cfg=cfg, variant=v, error=str(ex))), flush=True)
        eng.close()
        torch.cuda.empty_cache()
