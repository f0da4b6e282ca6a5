"""SF-TDA drivers (spin-flip up / down, non-spin-adapted) on the B200 sigma engine.

Same factory and classes as xtddft/SF_TDA.py:17-23, 408-447, 562-574, 588-622, 837-849:
`SF_TDA(mf, isf=-1, davidson=True, method=0)` -> `SF_TDA_down` / `SF_TDA_up`, `kernel(nstates) -> (e_eV, v)`.
method: 0 ALDA0, 1 multicollinear, 2 collinear.  (The shipped `gen_tda_operation_sf` ignores method=2 and applies the
ALDA0 kernel, SF_TDA.py:218-221; this build follows the documented meaning -- collinear = no grid term.)
"""
from __future__ import annotations

import numpy as np

from . import plan as planmod
from . import utils
from .adapters import problem_from_mf
from .drivers_common import TimeCounter, solve, timed_engine

ha2eV = utils.ha2eV


def SF_TDA(mf, isf=-1, davidson=True, method=0):
    if isf == -1:
        return SF_TDA_down(mf, method, davidson)
    if isf == 1:
        return SF_TDA_up(mf, method, davidson)
    raise ValueError("isf must be -1 (down) or +1 (up)")


class _SFBase:
    isf = -1

    def __init__(self, mf, method, davidson=True):
        self.mf = mf
        self.method = method
        self.davidson = davidson
        self.problem = problem_from_mf(mf, kernel={0: "alda0", 1: "mcol", 2: "none"}[method])
        p = self.problem
        self.nc, self.no, self.nv = p.nc, p.no, p.nv
        self.nao = p.nao
        self.tc = TimeCounter()
        self._engine = None

    def _get_engine(self):
        if self._engine is None:
            builder = lambda p: planmod.build_sf_plan(p, isf=self.isf, method=self.method, sa=0, layout=planmod.LAYOUT_PYSCF, hdiag_kind="sf")
            self._engine = timed_engine(self.tc, builder, self.problem, max_nvec=40)
            self.plan = self._engine.plan
        return self._engine

    def gen_tda_operation_sf(self):
        eng = self._get_engine()
        return eng.as_vind(), eng.hdiag()

    def get_Amat(self):
        raise NotImplementedError("get_Amat is the dense O(dim^2) path, outside the sigma hot path")

    def kernel(self, nstates=1):
        self.nstates = nstates
        if not self.davidson:
            raise NotImplementedError("davidson=False selects the dense get_Amat path, outside the sigma hot path")
        eng = self._get_engine()
        self.converged, self.e, v, self.Davidcyc, _ = solve(eng, nstates, "sf_down" if self.isf == -1 else "sf_up", tc=self.tc)
        if self.isf == -1:
            v = utils.deal_v_davidson(v, self.nc, self.no, self.nv)
        self.v = v
        return self.e[:nstates] * ha2eV, self.v[:, :nstates]


class SF_TDA_down(_SFBase):
    isf = -1

    def deltaS2(self):
        return utils.delta_s2_sf_roks(self.v, self.nc, self.no, self.nv)


class SF_TDA_up(_SFBase):
    isf = 1
