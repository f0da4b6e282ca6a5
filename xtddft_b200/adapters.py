"""What the driver classes accept as `mf`, and how it becomes a `ProblemData`.

  * a `ProblemData` (or a `SynthMF` wrapping one): synthetic / pre-extracted SCF quantities -- the path that is
    exercised and tested here (PySCF, libcint and libxc are not installable in the build or bench containers);
  * a PySCF ROKS/UKS (or ROHF/UHF) object: `from_pyscf` extracts exactly what the reference's `gen_vind()` closures
    capture (xtddft/XTDA.py:558-613, xtddft/SF_TDA.py:39-88,162-221, xtddft/XSF_TDA.py:1029-1121).  This adapter follows the
    PySCF 2.11/2.12 API from the reference's own call sites; it could NOT be executed here and is therefore
    "parity unpinned" until run next to a PySCF install.
"""
from __future__ import annotations

import numpy as np

from .problem import ProblemData, XC_GGA, XC_LDA, XC_MGGA, XC_NONE


class SynthMF:
    """Duck-typed stand-in for a converged SCF object over a ProblemData (used in tests, smoke and the bench)."""

    def __init__(self, problem: ProblemData, xc: str = "synthetic"):
        self.problem = problem
        self.xc = xc
        p = problem
        self.converged = True
        self.level_shift = p.level_shift
        self.max_memory = 4000
        if p.restricted:
            self.mo_coeff, self.mo_energy = p.mo_coeff[0], p.mo_energy[0]
            occ = np.zeros(p.nmo)
            occ[:p.nc] = 2
            occ[p.nc:p.nc + p.no] = 1
            self.mo_occ = occ
        else:
            self.mo_coeff, self.mo_energy = p.mo_coeff, p.mo_energy
            occ = np.zeros((2, p.nmo))
            occ[0, :p.nocc_a] = 1
            occ[1, :p.nocc_b] = 1
            self.mo_occ = occ
        self.mol = _SynthMol(p)

    def spin_square(self):
        s = self.problem.no / 2.0
        return s * (s + 1), 2 * s + 1


class _SynthMol:
    def __init__(self, p):
        self.spin = p.no
        self._nao = p.nao

    def nao_nr(self):
        return self._nao


def one_electron_ints(mf, problem: ProblemData, name: str):
    """One-electron AO integrals the property pass needs (`int1e_r`, `int1e_ipovlp`, `int1e_cg_irxp`, `int1e_ovlp`):
    from `problem.meta["one_electron"][name]` when the caller supplied them (synthetic inputs), else from libcint through
    the PySCF molecule with the reference's own calls (XTDA.py:847,869,872; XSF_TDA_GPU.py:956,1001; XSF_TDA.py:628).
    Returns None when neither source exists (the strengths are then left as None)."""
    given = problem.meta.get("one_electron") if isinstance(problem.meta, dict) else None
    if given is not None and name in given:
        return np.asarray(given[name], dtype=np.float64)
    mol = getattr(mf, "mol", None)
    if mol is None or not hasattr(mol, "intor"):
        return None
    if name == "int1e_r":                                  # pragma: no cover  (needs PySCF)
        return np.asarray(mol.intor_symmetric("int1e_r", comp=3))
    if name == "int1e_ovlp":                               # pragma: no cover
        return np.asarray(mf.get_ovlp())
    return np.asarray(mol.intor(name, comp=3, hermi=2))    # pragma: no cover


def is_chiral(mf, problem: ProblemData) -> bool:
    """XTDA.py:818-821 computes rotatory strengths only for chiral molecules (`gto.mole.chiral_mol`)."""
    flag = problem.meta.get("chiral") if isinstance(problem.meta, dict) else None
    if flag is not None:
        return bool(flag)
    try:                                                   # pragma: no cover  (needs PySCF)
        from pyscf import gto
        return bool(gto.mole.chiral_mol(mf.mol))
    except Exception:
        return False


def problem_from_mf(mf, **kw) -> ProblemData:
    if isinstance(mf, ProblemData):
        return mf
    if hasattr(mf, "problem") and isinstance(mf.problem, ProblemData):
        return mf.problem
    return from_pyscf(mf, **kw)


def from_pyscf(mf, kernel: str = "uks", collinear_samples: int = 60, auxbasis=None) -> ProblemData:   # pragma: no cover
    """Extract a ProblemData from a PySCF mean-field object.  kernel: 'uks' (X-TDA), 'alda0', 'mcol', 'none'.

    Needs `pyscf` (and `mcfun` for kernel='mcol').  Untested in this repository's environments (no PySCF)."""
    try:
        from pyscf import df, lib, scf
    except ImportError as e:
        raise ImportError("from_pyscf needs PySCF; pass a ProblemData / SynthMF instead") from e
    mol = mf.mol
    nao = mol.nao_nr()
    restricted = np.asarray(mf.mo_coeff).ndim == 2
    if restricted:
        c = np.asarray(mf.mo_coeff)
        occ = np.asarray(mf.mo_occ)
        nc, no = int((occ >= 2).sum()), int(((occ >= 1) & (occ < 2)).sum())
        nv = int((occ == 0).sum())
        mo_coeff = np.stack([c, c])
        mo_energy = np.stack([mf.mo_energy, mf.mo_energy])
    else:
        mo_coeff = np.asarray(mf.mo_coeff)
        occ = np.asarray(mf.mo_occ)
        na, nb = int((occ[0] > 0).sum()), int((occ[1] > 0).sum())
        nc, no, nv = nb, na - nb, mo_coeff.shape[2] - na
        mo_energy = np.asarray(mf.mo_energy)
    dm = mf.make_rdm1()
    vhf = mf.get_veff(mol, dm)
    if getattr(mf, "with_solvent", None) is not None:
        vhf = vhf + vhf.v_solvent
    h1e = mf.get_hcore()
    fock_ks = np.stack([mo_coeff[s].T @ (h1e + vhf[s]) @ mo_coeff[s] for s in (0, 1)])
    fock_hf = None
    if restricted:
        hf = scf.ROHF(mol)
        if getattr(mf, "with_x2c", None) is not None:
            hf = hf.x2c()
        veff = hf.get_veff(mol, dm)
        h1 = hf.get_hcore()
        fock_hf = np.stack([mo_coeff[s].T @ (h1 + veff[s]) @ mo_coeff[s] for s in (0, 1)])
    is_dft = hasattr(mf, "xc") and hasattr(mf, "_numint")
    omega = alpha = 0.0
    hyb = 1.0
    if is_dft:
        ni = mf._numint
        omega, alpha, hyb = ni.rsh_and_hybrid_coeff(mf.xc, mol.spin)
    # density-fitting tensor (the engine is DF-only; an mf without .with_df gets a fresh DF object)
    with_df = getattr(mf, "with_df", None)
    if with_df is None:
        with_df = df.DF(mol, auxbasis=auxbasis)
    if with_df._cderi is None:
        with_df.build()
    cderi = lib.unpack_tril(np.asarray(with_df._cderi)) if not isinstance(with_df._cderi, str) else \
        lib.unpack_tril(np.vstack([blk for blk in with_df.loop()]))
    cderi_lr = None
    if omega != 0 and alpha != hyb:
        with mol.with_range_coulomb(omega):
            lr = df.DF(mol, auxbasis=with_df.auxbasis).build()
            cderi_lr = lib.unpack_tril(np.asarray(lr._cderi))
    ao = weights = fxc_uks = fxc_alda0 = fxc_mcol = None
    xctype = XC_NONE
    if is_dft and kernel != "none":
        ni = mf._numint
        xt = ni._xc_type(mf.xc)
        if xt not in ("LDA", "GGA", "MGGA"):
            raise NotImplementedError(f"kernel type {xt} is not supported by the B200 path")
        # meta-GGA: value + gradient AO components and kernel tables with a tau component (no Laplacian, as in the reference)
        xctype = {"LDA": XC_LDA, "GGA": XC_GGA, "MGGA": XC_MGGA}[xt]
        coords, weights = mf.grids.coords, np.asarray(mf.grids.weights)
        aov = ni.eval_ao(mol, coords, deriv=0 if xt == "LDA" else 1)
        ao = aov[None] if xt == "LDA" else aov[:4]
        mo_occ2 = np.zeros((2, mo_coeff.shape[2]))
        mo_occ2[0, :nc + no] = 1
        mo_occ2[1, :nc] = 1
        if kernel == "uks":
            fxc_uks = ni.cache_xc_kernel(mol, mf.grids, mf.xc, mo_coeff, mo_occ2, 1)[2]
        elif kernel == "alda0":
            # SF_TDA.cache_xc_kernel_sf: (w v_a - w v_b) / (rho_a - rho_b + 1e-9), GGA evaluated with zeroed gradients
            rho = []
            for s in (0, 1):
                cs = mo_coeff[s][:, mo_occ2[s] > 0]
                rho.append(((ao[0] @ cs) ** 2).sum(1))
            if xt == "LDA":
                rr = (rho[0], rho[1])
            else:
                z = np.zeros((3, rho[0].size))
                rr = (np.vstack([rho[0], z]), np.vstack([rho[1], z]))
            vxc = ni.eval_xc_eff(mf.xc, rr, deriv=1, xctype=xt)[1]
            fxc_alda0 = (vxc[0, 0] * weights - vxc[1, 0] * weights) / (rho[0] - rho[1] + 1e-9)
        elif kernel == "mcol":
            from pyscf.dft import numint2c
            ni2 = numint2c.NumInt2C()
            ni2.collinear = "mcol"
            ni2.collinear_samples = collinear_samples
            raise NotImplementedError("multicollinear kernel extraction needs mcfun's eval_xc_eff_sf; supply fxc_mcol explicitly")
    return ProblemData(nao=nao, nc=nc, no=no, nv=nv, restricted=restricted, mo_coeff=mo_coeff, mo_energy=mo_energy, fock_ks=fock_ks,
                       fock_hf=fock_hf, cderi=cderi, cderi_lr=cderi_lr, hyb=hyb, alpha=alpha, omega=omega, xctype=xctype, ao=ao,
                       weights=weights, fxc_uks=fxc_uks, fxc_alda0=fxc_alda0, fxc_mcol=fxc_mcol,
                       level_shift=getattr(mf, "level_shift", 0.0) or 0.0, meta=dict(source="pyscf"))
