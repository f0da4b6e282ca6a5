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
    raise ValueError(method)


def oracle_vind_for(p, method: str):
    """Oracle builder matching `plan_for` (imported lazily: tests / smoke / cpu baseline only)."""
    from oracle import sigma as osig
    if method == "xtda":
        return osig.xtda_gen_vind(p)
    if method == "sf_down":
        return osig.sf_gen_vind(p, -1, 0)
    if method == "sf_up":
        return osig.sf_gen_vind(p, 1, 0)
    if method == "xsf":
        return osig.xsf_gen_vind(p, sa=3, method=0, remove=True)
    raise ValueError(method)


def engine_for_device_problem(dp: DeviceProblem, *, max_nvec: int, workspace_bytes: int, rank: int = 0, world: int = 1,
                              reducer=None) -> SigmaEngine:
    plan = plan_for(dp.p, dp.method)
    eng = SigmaEngine(plan, dp.p.nao, dp.p.mo_coeff, workspace_bytes=workspace_bytes, reducer=reducer)
    if eng.tensors_used:
        dp.stream_cderi(eng, 0, rank, world)
    dp.make_grid(eng, rank, world)
    eng.finalize(max_nvec)
    return eng
