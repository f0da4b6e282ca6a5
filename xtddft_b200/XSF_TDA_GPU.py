"""The reference's GPU spin-flip class (xtddft/XSF_TDA_GPU.py:130-222, 869-934, 1256-1303) on the B200 engine:
`XSF_TDA_GPU(mf, X=3, collinear='mcol', nstates=7, extype=1, gpu_davidson=False, collinear_samples=20, remove=None,
foo=1.0, d_lda=0.3, fglobal=None)`, `kernel() -> (e_eV, v)`; PySCF vector order at `vind`, block order in `self.v`.
`fglobal` given explicitly is honoured (the shipped class leaves self.fglobal unset in that case)."""
from __future__ import annotations

import numpy as np

from . import plan as planmod
from . import utils
from .adapters import one_electron_ints, problem_from_mf
from .drivers_common import TimeCounter, solve, timed_engine


class XSF_TDA_GPU:
    def __init__(self, mf, X=3, collinear="mcol", nstates=7, extype=1, gpu_davidson=False, collinear_samples=20, remove=None,
                 foo=1.0, d_lda=0.3, fglobal=None):
        if extype not in (0, 1):
            raise ValueError(f"Invalid extype = {extype}. extype must be 0 (spin flip up) or 1 (spin flip down).")
        self.method = {"alda0": 0, "mcol": 1, "col": 2, "ncol": 1}[collinear]
        self.mf = mf
        self.problem = problem_from_mf(mf, kernel={0: "alda0", 1: "mcol", 2: "none"}[self.method], collinear_samples=collinear_samples)
        p = self.problem
        self.level_shift, self.conv_tol, self.lindep, self.max_cycle = 0, 1e-5, 1e-12, 100
        self.collinear, self.collinear_samples, self.extype, self.gpu_davidson = collinear, collinear_samples, extype, gpu_davidson
        self.X = X if p.restricted else 0
        self.nc, self.no, self.nv = p.nc, p.no, p.nv
        nov = self.nc * self.nv if extype == 0 else (self.nc + self.no) * (self.no + self.nv)
        self.nstates = min(nstates, nov)
        self.re = p.restricted if remove is None else remove
        if extype == 0:
            self.re = False
        self.omega, self.alpha, self.hyb = p.omega, p.alpha, p.hyb
        self.fglobal = planmod.xsf_default_fglobal(p, self.method, d_lda, fit=True) if fglobal is None else fglobal
        self.foo = foo
        if self.re:
            self.vects = utils.get_vect(self.no)
        self.tc = TimeCounter()
        self._engine = None

    def get_vect(self):
        return utils.get_vect(self.no)

    def _get_engine(self):
        if self._engine is None:
            if self.extype == 0:
                builder = lambda p: planmod.build_sf_plan(p, isf=1, method=self.method, hdiag_kind="gpu")
            else:
                builder = lambda p: planmod.build_sf_plan(p, isf=-1, method=self.method, sa=self.X, layout=planmod.LAYOUT_PYSCF,
                                                          remove=self.re, foo=self.foo, fglobal=self.fglobal, hdiag_kind="gpu")
            self._engine = timed_engine(self.tc, builder, self.problem, max_nvec=40)
            self.plan = self._engine.plan
        return self._engine

    def gen_vind(self):
        eng = self._get_engine()
        return eng.as_vind(), eng.hdiag()

    def init_guess(self):
        from .davidson import init_guess
        p = self.problem
        ea, eb = p.mo_energy
        if self.extype == 0:
            gaps = (ea[p.nocc_a:] - eb[:p.nocc_b, None]).ravel()
            return init_guess(gaps, self.nstates, 1e-5)
        gaps = (eb[p.nocc_b:] - ea[:p.nocc_a, None]).ravel()
        x0 = init_guess(gaps, self.nstates, 1e-5)
        return x0[:, :-1] if self.re else x0          # XSF_TDA_GPU.py:261-262

    def deal_v_davidson(self):
        if self.extype == 0:
            return self.v
        return utils.deal_v_davidson(self.v, self.nc, self.no, self.nv, removed=self.re)

    def kernel(self):
        eng = self._get_engine()
        over = dict(tol_residual=self.conv_tol, lindep=self.lindep, max_cycle=self.max_cycle)
        self.converged, self.e, self.v, self.Davidcyc, _ = solve(eng, self.nstates, "gpu_class", x0=self.init_guess(), tc=self.tc, **over)
        self.v = self.deal_v_davidson()
        self.os = self.osc_str()      # state-to-state oscillator matrix; None without dipole integrals
        return self.e * utils.ha2eV, self.v

    # ---- property pass (XSF_TDA_GPU.py:936-1116) on the device ------------------------------------------
    def _tdm(self):
        from .properties import PropertyPass
        dip = one_electron_ints(self.mf, self.problem, "int1e_r")
        if dip is None or self.extype == 0:
            return None
        pp = PropertyPass(self.problem)
        rows = np.ascontiguousarray(np.asarray(self.v).T)
        if self.problem.restricted:
            return pp.tdm_r(rows, dip, self.X, planmod.LAYOUT_BLOCK, bool(self.re))
        return pp.tdm_u(rows, dip, planmod.LAYOUT_BLOCK)

    def calculate_TDM_R(self):
        assert self.problem.restricted, "Must be ROHF/ROKS reference !!!"
        return self.osc_str()

    def calculate_TDM_U(self):
        assert not self.problem.restricted, "Must be UHF/UKS reference !!!"
        return self.osc_str()

    def osc_str(self):
        from .properties import PropertyPass
        tdm = self._tdm()
        self.tdm = tdm
        return None if tdm is None else PropertyPass.osc_matrix(self.e, tdm)
