"""XSF-TDA / USF-TDA driver (spin-adapted spin-flip-down TDA) on the B200 sigma engine.

Same constructor, `kernel` signature, defaults and result attributes as `XSF_TDA` in xtddft/XSF_TDA.py:146-213,
1455-1481, 1501-1554: block-order vectors cv|co|ov|oo, optional removal of the S_f = S_i OO component
(`remove`, default True for ROKS), spin-adaptation level SA 0..3, `foo`, `d_lda` / `fglobal`.
"""
from __future__ import annotations

import numpy as np

from . import plan as planmod
from . import utils
from .adapters import one_electron_ints, problem_from_mf
from .drivers_common import TimeCounter, solve, timed_engine

au2ev = utils.au2ev_xsf


class XSF_TDA:
    def __init__(self, mf, SA=None, davidson=True, method=0, collinear_samples=60, calculate_sp=False):
        self.mf = mf
        self.problem = problem_from_mf(mf, kernel={0: "alda0", 1: "mcol", 2: "none"}[method], collinear_samples=collinear_samples)
        p = self.problem
        self.type_u = not p.restricted
        self.SA = (0 if self.type_u else 3) if SA is None else SA
        self.davidson, self.method, self.collinear_samples = davidson, method, collinear_samples
        self.nc, self.no, self.nv = p.nc, p.no, p.nv
        self.nao = p.nao
        self.ground_s = p.no / 2.0
        self.omega, self.alpha, self.hyb = p.omega, p.alpha, p.hyb
        if calculate_sp:
            raise NotImplementedError("get_sp (spin-polarisation diagnostics) is outside the sigma hot path")
        self.tc = TimeCounter()
        self._engine = None
        self._engine_key = None

    def get_vect(self):
        return utils.get_vect(self.no)

    def _get_engine(self, foo, fglobal):
        key = (self.re, self.SA, foo, fglobal)
        if self._engine is None or self._engine_key != key:
            if self._engine is not None:
                self._engine.close()
            builder = lambda p: planmod.build_sf_plan(p, isf=-1, method=self.method, sa=self.SA, layout=planmod.LAYOUT_BLOCK, remove=self.re,
                                                      foo=foo, fglobal=fglobal, hdiag_kind="xsf")
            self._engine = timed_engine(self.tc, builder, self.problem, max_nvec=40)
            self.plan = self._engine.plan
            self._engine_key = key
        return self._engine

    def gen_tda_operation_sf(self, foo, fglobal):
        eng = self._get_engine(foo, fglobal)
        return eng.as_vind(), eng.hdiag()

    def get_Amat(self, *a, **k):
        raise NotImplementedError("get_Amat is the dense O(dim^2) path, outside the sigma hot path")

    def kernel(self, nstates=1, remove=None, frozen=None, foo=1.0, d_lda=0.3, fglobal=None, fit=True):
        self.re = (not self.type_u) if remove is None else remove
        nov = (self.nc + self.no) * (self.no + self.nv)
        self.nstates = min(nstates, nov)
        if fglobal is None:
            fglobal = planmod.xsf_default_fglobal(self.problem, self.method, d_lda, fit)
        if frozen is not None or not self.davidson:
            raise NotImplementedError("frozen / dense get_Amat paths are outside the sigma hot path")
        if self.re:
            self.vects = self.get_vect()
        eng = self._get_engine(foo, fglobal)
        self.converged, self.e, self.v, self.Davidcyc, _ = solve(eng, self.nstates, "xsf", tc=self.tc)
        return self.e * au2ev, self.v

    def deltaS2(self):
        return utils.delta_s2_sf_roks(self.v, self.nc, self.no, self.nv, self.vects if self.re else None)

    # ---- property pass (XSF_TDA.py:429-592, 613-649, 729-790) on the device ---------------------------------
    def _property_pass(self):
        from .properties import PropertyPass
        if getattr(self, "_pp", None) is None:
            self._pp = PropertyPass(self.problem)
        return self._pp

    def calculate_TDM(self, verbose=True):
        """Excited-state to excited-state transition dipoles and oscillator strengths; returns (tdm[3,n,n], osc[n,n]) and
        prints the reference's table."""
        dip = one_electron_ints(self.mf, self.problem, "int1e_r")
        if dip is None:
            raise ValueError("calculate_TDM needs the dipole integrals (a PySCF molecule, or problem.meta['one_electron'])")
        pp = self._property_pass()
        rows = np.ascontiguousarray(np.asarray(self.v).T)
        if self.type_u:
            tdm = pp.tdm_u(rows, dip, planmod.LAYOUT_BLOCK)
        else:
            tdm = pp.tdm_r(rows, dip, self.SA, planmod.LAYOUT_BLOCK, bool(self.re))
        osc = pp.osc_matrix(self.e, tdm)
        if verbose:
            print("Excited state to Excited state transition dipole moments(Au)")
            print("State State    X     Y     Z     OSC.")
            for i in range(len(self.e)):
                for j in range(len(self.e)):
                    print(f"{i+1:2d} {j+1:2d} {tdm[0, i, j]:>8.4f} {tdm[1, i, j]:>8.4f} {tdm[2, i, j]:>8.4f}  {osc[i, j]:>8.4f} ")
        return tdm, osc

    def calculate_TDM_R(self, verbose=True):
        assert not self.type_u
        return self.calculate_TDM(verbose)

    def calculate_TDM_U(self, verbose=True):
        assert self.type_u, "Must be UHF/UKS reference !!!"
        return self.calculate_TDM(verbose)

    def deltaS2_U(self, nstate=None):
        """P_ab of XSF_TDA.py:613-649 for state `nstate` (all states if None); D<S^2> = P_ab - no + 1."""
        ovlp = one_electron_ints(self.mf, self.problem, "int1e_ovlp")
        if ovlp is None:
            ovlp = np.eye(self.nao)
        rows = np.ascontiguousarray(np.asarray(self.v).T)
        pab = self._property_pass().delta_s2_u(rows, ovlp, planmod.LAYOUT_BLOCK) + self.no - 1
        return pab if nstate is None else pab[nstate]

    def analyse(self):
        """D<S^2> labels of XSF_TDA.py:771-787 (symmetry labels need PySCF's symm module and are reported as 'A')."""
        if self.SA == 0 and not self.type_u:
            ds = list(self.deltaS2())
        elif self.type_u:
            ds = list(self.deltaS2_U() - self.no + 1)
        else:
            ds = []
        return ds, ["A"] * self.nstates
