"""X-TDA driver (spin-conserving spin-adapted TDA for ROKS / UKS references) on the B200 sigma engine.

Same constructor, `kernel()` / `Davidson()` entry points, options and result attributes as the reference classes
`XTDA` in xtddft/XTDA.py:21-54,694-829 (CPU) and xtddft/XTDA_GPU.py:24-53,368-500 (GPU); the sigma build
`vind` (XTDA.py:615-690) runs in libxtdsigma.so, and so does the property pass after the solve (oscillator and
rotatory strengths, XTDA.py:838-890 -> xtddft_b200/properties.py).  The dense debug paths (`full_diag`, `X_TDA` tensor
basis) are outside the hot path and raise NotImplementedError.
"""
from __future__ import annotations

import numpy as np

from . import plan as planmod
from . import utils
from .adapters import is_chiral, one_electron_ints, problem_from_mf
from .drivers_common import TimeCounter, solve, timed_engine


class XTDA:
    def __init__(self, mol, mf, nstates=10, basis="orbital", so2st=True, use_Davidson=True):
        self.mol, self.mf = mol, mf
        self.nstates = nstates
        self.basis = basis
        self.so2st = so2st
        self.use_Davidson = use_Davidson
        # pyscf.tdscf.rhf.TDBase defaults the reference copies (XTDA.py:29-31)
        self.conv_tol, self.lindep, self.max_cycle = 1e-5, 1e-12, 100
        self.deg_eia_thresh, self.positive_eig_threshold = 1e-3, 1e-3
        self.problem = problem_from_mf(mf, kernel="uks")
        self.X = bool(self.problem.restricted)
        self.tc = TimeCounter()
        self._settings = "xtda"
        self._engine = None

    # ---- operator interface (XTDA.py:694-698) ----------------------------------------------------------
    def _get_engine(self):
        if self._engine is None:
            self._engine = timed_engine(self.tc, planmod.build_xtda_plan, self.problem, max_nvec=40)
            self.plan = self._engine.plan
        return self._engine

    def gen_vind(self, mf=None):
        assert mf is None or mf is self.mf
        eng = self._get_engine()
        return eng.as_vind(), eng.hdiag()

    def get_init_guess(self, mf=None, nstates=None, wfnsym=None, return_symmetry=False):
        """Koopmans unit vectors on the lowest orbital-energy gaps (+1e-3 window), XTDA.py:700-734."""
        from .davidson import init_guess
        p = self.problem
        ea, eb = p.mo_energy
        e_a = ea[p.nocc_a:] - ea[:p.nocc_a, None]
        e_b = eb[p.nocc_b:] - eb[:p.nocc_b, None]
        gaps = np.append(e_a.ravel(), e_b.ravel())
        return init_guess(gaps, self.nstates if nstates is None else nstates, self.deg_eia_thresh)

    init_guess = get_init_guess

    def kernel(self, x0=None, nstates=None):
        if self.basis == "tensor":
            raise NotImplementedError("tensor-basis X_TDA is a dense O(N^4) reference path, outside the sigma hot path")
        if self.basis != "orbital":
            raise ValueError("basis must be tensor or orbital")
        if not self.use_Davidson:
            raise NotImplementedError("full_diag is a dense debug path, outside the sigma hot path")
        return self.Davidson(x0=x0, nstates=nstates), self.v

    def Davidson(self, x0=None, nstates=None):
        if nstates is not None:
            self.nstates = nstates
        p = self.problem
        eng = self._get_engine()
        if x0 is None:
            x0 = self.get_init_guess(self.mf, self.nstates)
        over = dict(tol_residual=self.conv_tol, lindep=self.lindep, max_cycle=self.max_cycle)
        self.converged, self.e, x1, self.Davidcyc, _ = solve(eng, self.nstates, self._settings, x0=x0, tc=self.tc, **over)
        nc, no, nv = p.nc, p.no, p.nv
        nocca, noccb, nvira = nc + no, nc, nv
        self.nc, self.no, self.nv = nc, no, nv
        self.order = utils.order_pyscf2my(nc, no, nv)
        self.v = x1[self.order, :]
        self.xy_a = self.v.T[:, :nocca * nvira]
        self.xy_b = self.v.T[:, nocca * nvira:]
        self.xycv_a = self.v.T[:, :noccb * nvira]
        self.xyov_a = self.v.T[:, noccb * nvira:nocca * nvira]
        self.xyco_b = self.v.T[:, nocca * nvira:nocca * nvira + noccb * no]
        self.xycv_b = self.v.T[:, nocca * nvira + noccb * no:]
        self.dS2 = self.deltaS2()
        self._x_rows = np.ascontiguousarray(x1.T)          # PySCF-order amplitude rows for the property pass
        self.os = self.osc_str()                           # None when no dipole integrals are available
        self.rs = self.rot_str() if is_chiral(self.mf, p) else np.zeros(self.nstates)     # XTDA.py:818-821
        if self.so2st:
            self.v = utils.so2st(self.v, nc, no, nv)
        return self.e

    def deltaS2(self):
        """XTDA.py:831-836: |X_cv(aa) - X_cv(bb)|^2."""
        d = self.xycv_a - self.xycv_b
        return np.einsum("ij,ij->i", d, d)

    def _property_pass(self):
        from .properties import PropertyPass
        if getattr(self, "_pp", None) is None:
            self._pp = PropertyPass(self.problem)
        return self._pp

    def osc_str(self):
        """Length-form oscillator strengths f = 2/3 w |<0|r|n>|^2 (XTDA.py:838-858), on the device."""
        dip = one_electron_ints(self.mf, self.problem, "int1e_r")
        if dip is None:
            return None
        return self._property_pass().xtda_osc_str(self.e[:self.nstates], self._x_rows, dip)

    def rot_str(self):
        """Rotatory strengths in cgs units (XTDA.py:860-890), on the device."""
        ipo = one_electron_ints(self.mf, self.problem, "int1e_ipovlp")
        rxp = one_electron_ints(self.mf, self.problem, "int1e_cg_irxp")
        if ipo is None or rxp is None:
            return None
        return self._property_pass().xtda_rot_str(self.e[:self.nstates], self._x_rows, ipo, rxp)

    def full_diag(self):
        raise NotImplementedError("full_diag is a dense O(dim^2) debug path, outside the sigma hot path")
