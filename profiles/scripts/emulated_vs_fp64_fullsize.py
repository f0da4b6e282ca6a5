"""Max relative difference of sigma between the engine's emulated path (INT8 tensor cores, S digit planes) and its FP64 DMMA path on
the same full-size BASELINE inputs (configs 5, 4, 3), 3 random unit vectors.  One JSON line per (config, S)."""
import json, sys
import torch
sys.path.insert(0, ".")
from xtddft_b200.synth_device import make_device_problem
from xtddft_b200.workloads import default_workspace_bytes, engine_for_device_problem

for cfg in (5, 4, 3):
    ref, z = None, None
    for slices in (0, 5, 6, 7):
        torch.cuda.empty_cache()
        dp = make_device_problem(cfg, 1.0)
        eng = engine_for_device_problem(dp, max_nvec=4, workspace_bytes=min(default_workspace_bytes(dp, 1), 16 << 30), exchange_slices=slices)
        if z is None:
            g = torch.Generator(device="cuda"); g.manual_seed(21)
            z = torch.randn((3, eng.ext_dim), generator=g, device="cuda", dtype=torch.float64)
            z /= z.norm(dim=1, keepdim=True)
        out = eng.sigma(z).cpu()
        eng.close(); del eng
        if slices == 0:
            ref = out
        else:
            err = (out - ref).abs().max().item() / max(1.0, ref.abs().max().item())
            print(json.dumps(dict(config=cfg, workload=dp.name, digit_planes=slices, max_rel_sigma_difference_vs_fp64=err)), flush=True)
