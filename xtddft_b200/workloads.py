"""BASELINE workloads -> plan + engine (shared by bench.py, the smoke test and the GPU tests)."""
from __future__ import annotations

from . import plan as planmod
from .engine import SigmaEngine
from .synth_device import DeviceProblem


def plan_for(p, method: str):
    if method == "xtda":
        return planmod.build_xtda_plan(p)
    if method == "sf_down":
        return planmod.build_sf_plan(p, isf=-1, method=0, sa=0, layout=planmod.LAYOUT_PYSCF, hdiag_kind="sf")
    if method == "sf_up":
        return planmod.build_sf_plan(p, isf=1, method=0)
    if method == "xsf":
        return planmod.build_sf_plan(p, isf=-1, method=0, sa=3, layout=planmod.LAYOUT_BLOCK, remove=True, hdiag_kind="xsf")
    if method == "zvector":                 # the Z-vector operator on the same inputs as the X-TDA workloads
        return planmod.build_zvector_plan(p)
    raise ValueError(method)


def default_workspace_bytes(dp: DeviceProblem, world: int = 1) -> int:
    """Workspace for the per-call buffers: what is left of HBM after the resident MO-basis tensor blocks and AO values
    of this rank's shard, capped at 24 GiB (more does not change the chunking noticeably)."""
    import torch
    free, _ = torch.cuda.mem_get_info()
    p = dp.p
    pad = lambda n: (n + 15) // 16 * 16
    nv_i, no_i = p.nvir_b + 2, p.nocc_a + 2
    resident = (dp.naux // world + 1) * (nv_i * pad(nv_i) + no_i * pad(no_i)) * 8 * (2 if dp.method in ("xtda", "zvector") else 1)
    if dp.fxc_kind != "none":
        nve = 1 if dp.fxc_kind == "alda0" else dp.nvar
        nch = 2 if dp.method in ("xtda", "zvector") else 1
        resident += (dp.ng // world + 1) * (pad(nv_i) + pad(no_i)) * nve * nch * 8          # MO values on the grid
    return int(min(24 << 30, max(1 << 30, (free - resident) * 0.55)))


def engine_for_device_problem(dp: DeviceProblem, *, max_nvec: int, workspace_bytes: int, rank: int = 0, world: int = 1,
                              reducer=None, exchange_slices=None) -> SigmaEngine:
    plan = plan_for(dp.p, dp.method)
    eng = SigmaEngine(plan, dp.p.nao, dp.p.mo_coeff, workspace_bytes=workspace_bytes, reducer=reducer, exchange_slices=exchange_slices)
    dp.make_grid(eng, rank, world)
    eng.grid_commit()                   # MO values on the grid; the AO array is released before the tensor streams in
    eng.torch.cuda.empty_cache()
    if eng.tensors_used:
        dp.stream_cderi(eng, 0, rank, world)
    eng.finalize(max_nvec)
    return eng
