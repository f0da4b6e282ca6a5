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
    """Same role as the reference's timing bag (XTDA_GPU.py:18-21): attributes filled by the driver."""
    pass


def make_engine(plan, p, max_nvec: int = 40, workspace_bytes: Optional[int] = None, distributed: bool = True) -> SigmaEngine:
    import torch
    reducer = None
    rank, world = 0, 1
    if distributed and torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
        reducer = SigmaReducer()
        rank, world = reducer.rank, reducer.world
    if workspace_bytes is None:
        free, _ = torch.cuda.mem_get_info()
        workspace_bytes = int(min(8 << 30, max(256 << 20, free // 4)))
    return SigmaEngine.from_problem(plan, p, max_nvec=max_nvec, workspace_bytes=workspace_bytes, reducer=reducer, rank=rank, world=world)


def solve(eng: SigmaEngine, nstates: int, settings: str, x0=None, tc: Optional[TimeCounter] = None, verbose: int = 0, **over):
    cfg = dict(dav.SOLVER[settings])
    cfg.update(over)
    hdiag = eng.hdiag()
    nroots = min(nstates, hdiag.size)
    if x0 is None:
        x0 = dav.init_guess(hdiag, nroots, cfg["window"])
    t0 = time.perf_counter()
    conv, e, x, cyc = dav.davidson1(eng.sigma, np.asarray(x0), hdiag, tol=cfg["tol"], tol_residual=cfg["tol_residual"], lindep=cfg["lindep"],
                                    max_cycle=cfg["max_cycle"], nroots=nroots, level_shift=cfg["level_shift"],
                                    pick=dav.pick_positive if cfg["pick_positive"] else None, verbose=verbose)
    if tc is not None:
        tc.dv = time.perf_counter() - t0
    return conv, e, np.array(x).T, cyc, hdiag
