"""Shared machinery of the driver classes: engine construction and the device Davidson call."""
from __future__ import annotations

import time
from typing import Optional

import numpy as np

from . import davidson as dav
from .adapters import problem_from_mf
from .dist import SigmaReducer
from .engine import SigmaEngine


class TimeCounter:
    """The reference's timing bag (XTDA_GPU.py:18-21, printed at XTDA_GPU.py:481-499 / XSF_TDA_GPU.py:1285-1302), seconds:
      Ap     engine set-up (upload, MO transforms of the grid basis and of the 3-centre tensor)     Ap_f  Fock extraction
      Ap_k   kernel tables / MO values on the grid                                                dAp   Delta-A plan compilation
      Adv    all sigma builds                A_vxc  grid phases (GEMMs, weighting, slicing)        A_gk  DF Coulomb / exchange phases
      dAdv   local Fock / Delta-A coupling terms (the Delta-A exchange images are block weights inside the same GEMMs as A)
      dv     Davidson wall time              allreduce  NCCL all-reduce of the partial sigma blocks (multi-GPU)
    The per-phase entries come from CUDA events inside libxtdsigma (xtd_get_stats) and are collected for problems with
    >= 50 000 unknowns (below that a sigma call is a replayed CUDA graph and only the total is timed)."""
    Ap = Ap_f = Ap_k = dAp = Adv = A_vxc = A_gk = dAdv = dv = allreduce = 0.0


def make_engine(plan, p, max_nvec: int = 40, workspace_bytes: Optional[int] = None, distributed: bool = True) -> SigmaEngine:
    """`plan`: a compiled Plan, or a callable `p -> Plan`.  With a callable, a ROKS problem that carries no ROHF-form Fock matrices
    (`p.fock_hf is None`) gets their spin difference K[D_open] from the device while the tensor streams in (SigmaEngine.from_problem)."""
    import torch
    builder = plan if callable(plan) else None
    if builder is not None:
        plan = None if (p.restricted and p.fock_hf is None) else builder(p)
    reducer = None
    rank, world = 0, 1
    if distributed and torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
        reducer = SigmaReducer()
        rank, world = reducer.rank, reducer.world
    if workspace_bytes is None:
        free, _ = torch.cuda.mem_get_info()
        workspace_bytes = int(min(8 << 30, max(256 << 20, free // 4)))
    return SigmaEngine.from_problem(plan, p, max_nvec=max_nvec, workspace_bytes=workspace_bytes, reducer=reducer, rank=rank, world=world,
                                    plan_builder=builder)


def timed_engine(tc: Optional[TimeCounter], plan, p, **kw) -> SigmaEngine:
    """make_engine with the set-up seconds recorded in tc.Ap (XTDA_GPU.py:196-215 counts the same stage)."""
    import torch
    t0 = time.perf_counter()
    eng = make_engine(plan, p, **kw)
    torch.cuda.synchronize()
    if tc is not None:
        tc.Ap = time.perf_counter() - t0
    return eng


def solve(eng: SigmaEngine, nstates: int, settings: str, x0=None, tc: Optional[TimeCounter] = None, verbose: int = 0, **over):
    cfg = dict(dav.SOLVER[settings])
    cfg.update({k: v for k, v in over.items() if k != "python_solver"})
    hdiag = eng.hdiag()
    nroots = min(nstates, hdiag.size)
    if x0 is None:
        x0 = dav.init_guess(hdiag, nroots, cfg["window"])
    if eng.ext_dim < dav.NATIVE_MAX_DIM and not verbose and not over.get("python_solver"):
        # launch-bound molecules: the solver loop runs inside libxtdsigma (xtd_davidson)
        t0 = time.perf_counter()
        conv, e, x, cyc = dav.davidson_native(eng, np.asarray(x0), hdiag, nroots, tol=cfg["tol"], tol_residual=cfg["tol_residual"],
                                              lindep=cfg["lindep"], max_cycle=cfg["max_cycle"], level_shift=cfg["level_shift"],
                                              pick_positive=cfg["pick_positive"])
        if tc is not None:
            tc.dv = time.perf_counter() - t0
        return conv, e, np.array(x).T, cyc, hdiag
    aop = eng.sigma
    detailed = tc is not None and eng.ext_dim >= 50000
    if detailed:
        acc = dict(Adv=0.0, A_vxc=0.0, A_gk=0.0, dAdv=0.0, allreduce=0.0)

        def aop(z, out=None):
            t1 = time.perf_counter()
            r = eng.sigma(z, out)
            ms = eng.stats()["ms"]                       # synchronises; per-phase device times of this call
            acc["Adv"] += time.perf_counter() - t1
            acc["A_vxc"] += 1e-3 * (ms["xc_gemm"] + ms["xc_stream"] + ms["xc_slice"])
            acc["A_gk"] += 1e-3 * (ms["k1"] + ms["k2"] + ms["k2_slice"] + ms["j"])
            acc["dAdv"] += 1e-3 * ms["local"]
            acc["allreduce"] += 1e-3 * ms["allreduce"]
            return r
    t0 = time.perf_counter()
    conv, e, x, cyc = dav.davidson1(aop, np.asarray(x0), hdiag, tol=cfg["tol"], tol_residual=cfg["tol_residual"], lindep=cfg["lindep"],
                                    max_cycle=cfg["max_cycle"], nroots=nroots, level_shift=cfg["level_shift"],
                                    pick=dav.pick_positive if cfg["pick_positive"] else None, verbose=verbose)
    if tc is not None:
        tc.dv = time.perf_counter() - t0
        if detailed:
            for k, v in acc.items():
                setattr(tc, k, v)
    return conv, e, np.array(x).T, cyc, hdiag
