"""What the driver classes accept as `mf`, and how it becomes a `ProblemData`.

  * a `ProblemData` (or a `SynthMF` wrapping one): synthetic / pre-extracted SCF quantities -- the path that is
    exercised and tested here (PySCF, libcint and libxc are not installable in the build or bench containers);
  * a PySCF ROKS/UKS (or ROHF/UHF) object: `from_pyscf` extracts exactly what the reference's `gen_vind()` closures
    capture (xtddft/XTDA.py:558-613, xtddft/SF_TDA.py:39-88,162-221, xtddft/XSF_TDA.py:1029-1121), duck-typed on the
    attributes the reference itself touches.  PySCF is not installable here, so the adapter is executed in the tests against
    the stand-in mean-field objects that tests/golden/make_golden.py feeds to the REFERENCE's own classes: driver class ->
    adapter -> engine must reproduce the golden sigma vectors the reference produced from the same object.  Real-molecule
    energies (Be, HF, CH2O+) stay unpinned until a PySCF install is reachable (tests/test_gpu_adapter.py has the skip-if test).
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


class NotDensityFittedError(ValueError):
    """The mean-field object carries no density-fitting tensor: the reference would contract exact 4-centre integrals for it
    (`mf.get_jk`, XTDA.py:520-539), this build is density-fitted only."""


def _tril_source(with_df):
    """(source, naux) of PySCF's packed 3-centre tensor: the in-memory array `_cderi` itself, or `with_df.loop` for a tensor that
    lives on disk -- both in lower-triangular packed rows [naux, nao(nao+1)/2]; never unpacked on the host."""
    cd = getattr(with_df, "_cderi", None)
    if cd is None and hasattr(with_df, "build"):
        with_df.build()
        cd = with_df._cderi
    if isinstance(cd, np.ndarray):
        return np.asarray(cd, dtype=np.float64), int(cd.shape[0])
    naux = int(with_df.get_naoaux())
    return (lambda: with_df.loop()), naux


def _kernel_and_grid(mf, mol, kernel, mo_coeff, mo_occ2, restricted, collinear_samples, fxc_mcol):
    """AO values on the grid and the cached kernel, through the same numint calls as the reference: `ni.block_loop`
    (XTDA.py:514 via nr_uks_fxc, SF_TDA.py:63-64), `ni.cache_xc_kernel(..., spin=1)` (XTDA.py:504), the ALDA0 formula of
    SF_TDA.cache_xc_kernel_sf (SF_TDA.py:39-88), the multicollinear sampling of cache_xc_kernel_sf_mc (SF_TDA.py:942-974)."""
    ni = mf._numint
    xt = ni._xc_type(mf.xc)
    if xt not in ("LDA", "GGA", "MGGA"):
        raise NotImplementedError(f"kernel type {xt} is not supported by the B200 path")
    xctype = {"LDA": XC_LDA, "GGA": XC_GGA, "MGGA": XC_MGGA}[xt]
    nao = mol.nao_nr()
    ao_deriv = 0 if xt == "LDA" else 1        # meta-GGA: value + gradient only (MGGA_DENSITY_LAPL off, as in the reference)
    make_rho = None
    if kernel in ("alda0", "mcol"):
        dm0 = mf.make_rdm1()
        if restricted:
            try:                               # SF_TDA.py:58-60
                dm0.mo_coeff = (mf.mo_coeff, mf.mo_coeff)
                dm0.mo_occ = mo_occ2
            except AttributeError:
                pass
        make_rho = ni._gen_rho_evaluator(mol, dm0, hermi=0, with_lapl=False)[0]
    aos, ws, f_alda0, rhoa_all, rhob_all = [], [], [], [], []
    for ao, mask, weight, coords in ni.block_loop(mol, mf.grids, nao, ao_deriv, getattr(mf, "max_memory", 2000)):
        ao = np.asarray(ao)
        aos.append(ao[None] if ao.ndim == 2 else ao[:4])
        ws.append(np.asarray(weight))
        if make_rho is not None:
            rhoa, rhob = make_rho(0, ao, mask, xt), make_rho(1, ao, mask, xt)
            if kernel == "mcol":
                rhoa_all.append(np.asarray(rhoa)); rhob_all.append(np.asarray(rhob))
                continue
            if xt == "LDA":
                rho = (rhoa, rhob)
            else:                              # GGA / meta-GGA evaluated with the gradients zeroed (SF_TDA.py:74-80)
                rha, rhb = np.zeros_like(rhoa), np.zeros_like(rhob)
                rha[0], rhb[0] = rhoa[0], rhob[0]
                rho = (rha, rhb)
                rhoa, rhob = rhoa[0], rhob[0]
            vxc = ni.eval_xc_eff(mf.xc, rho, deriv=1, xctype=xt)[1]
            f_alda0.append((vxc[0, 0] * weight - vxc[1, 0] * weight) / (np.asarray(rhoa) - np.asarray(rhob) + 1e-9))
    ao = np.ascontiguousarray(np.concatenate(aos, axis=1), dtype=np.float64)
    weights = np.ascontiguousarray(np.concatenate(ws), dtype=np.float64)
    fxc_uks = fxc_alda0 = None
    if kernel == "uks":
        fxc_uks = np.asarray(ni.cache_xc_kernel(mol, mf.grids, mf.xc, mo_coeff, mo_occ2, 1)[2], dtype=np.float64)
    elif kernel == "alda0":
        fxc_alda0 = np.ascontiguousarray(np.concatenate(f_alda0), dtype=np.float64)
    elif kernel == "mcol" and fxc_mcol is None:
        try:
            import mcfun
            from pyscf.dft import xc_deriv
        except ImportError as e:
            raise ImportError("the multicollinear kernel is sampled by the third-party `mcfun` package (SF_TDA.py:924-938), which is "
                              "not importable here: install it, or pass the cached kernel as fxc_mcol=[nvar, nvar, ngrids]") from e
        rho_ab = np.asarray((np.hstack(rhoa_all), np.hstack(rhob_all)))
        rho_tmz = np.zeros_like(rho_ab) + 1e-11            # SF_TDA.py:967-970
        rho_tmz[0] += rho_ab[0] + rho_ab[1]
        rho_tmz[1] += rho_ab[0] - rho_ab[1]

        def fn_eval_xc(rho, deriv):
            evfk = list(ni.eval_xc_eff(mf.xc, rho, deriv=deriv, xctype=xt))
            for order in range(1, deriv + 1):
                if evfk[order] is not None:
                    evfk[order] = xc_deriv.ud2ts(evfk[order])
            return evfk
        fxc_mcol = mcfun.eval_xc_eff_sf(fn_eval_xc, rho_tmz, 2, collinear_samples=collinear_samples)
        fxc_mcol = fxc_mcol[2] if isinstance(fxc_mcol, (tuple, list)) else fxc_mcol
    if fxc_mcol is not None:
        fxc_mcol = np.ascontiguousarray(fxc_mcol, dtype=np.float64)
    return xctype, ao, weights, fxc_uks, fxc_alda0, fxc_mcol


def from_pyscf(mf, kernel: str = "uks", collinear_samples: int = 60, auxbasis=None, fxc_mcol=None, rohf_fock: str = "pyscf") -> ProblemData:
    """Extract a ProblemData from a converged PySCF-style ROKS / UKS (ROHF / UHF) object -- exactly what the reference's
    `gen_vind()` closures capture (XTDA.py:558-613, SF_TDA.py:162-221, XSF_TDA.py:1029-1121).  kernel: 'uks' (X-TDA),
    'alda0', 'mcol', 'none'.

    Duck-typed: only attributes and methods of `mf` are used (`mol.nao_nr`, `mo_coeff/mo_occ/mo_energy`, `make_rdm1`, `get_veff`,
    `get_hcore`, `xc`, `_numint`, `grids`, `with_df`), so it runs on PySCF objects and on the stand-in mean-field objects of
    tests/golden/make_golden.py alike (tests/test_adapter_cpu.py, tests/test_gpu_adapter.py execute it).  `scf.ROHF(mol)` --
    the pure-HF Fock provider of the spin-adaptation terms (XTDA.py:607-613, XSF_TDA.py:1103-1111) -- is looked up in
    `pyscf.scf` when that is importable.

    The engine is density-fitted only.  An `mf` without `with_df` raises NotDensityFittedError unless `auxbasis` is given
    explicitly (the caller then accepts the DF approximation, which differs from the reference's exact integrals by the fitting
    error).  The 3-centre tensor stays in PySCF's packed storage and is streamed to the device block by block."""
    mol = mf.mol
    nao = mol.nao_nr()
    restricted = np.asarray(mf.mo_coeff).ndim == 2
    if restricted:
        c = np.asarray(mf.mo_coeff, dtype=np.float64)
        occ = np.asarray(mf.mo_occ)
        nc, no = int((occ >= 2).sum()), int(((occ >= 1) & (occ < 2)).sum())
        nv = int((occ == 0).sum())
        mo_coeff = np.stack([c, c])
        mo_energy = np.stack([mf.mo_energy, mf.mo_energy]).astype(np.float64)
    else:
        mo_coeff = np.asarray(mf.mo_coeff, dtype=np.float64)
        occ = np.asarray(mf.mo_occ)
        na, nb = int((occ[0] > 0).sum()), int((occ[1] > 0).sum())
        nc, no, nv = nb, na - nb, mo_coeff.shape[2] - na
        mo_energy = np.asarray(mf.mo_energy, dtype=np.float64)
    nmo = mo_coeff.shape[2]
    mo_occ2 = np.zeros((2, nmo))
    mo_occ2[0, :nc + no] = 1
    mo_occ2[1, :nc] = 1
    dm = mf.make_rdm1()
    vhf = mf.get_veff(mol, dm)
    if getattr(mf, "with_solvent", None) is not None:
        vhf = vhf + vhf.v_solvent
    h1e = np.asarray(mf.get_hcore())
    vhf = np.asarray(vhf)
    fock_ks = np.stack([mo_coeff[s].T @ (h1e + vhf[s]) @ mo_coeff[s] for s in (0, 1)])
    fock_hf = None
    if restricted and rohf_fock != "device":
        import importlib
        try:
            scf = importlib.import_module("pyscf.scf")
        except ImportError as e:
            raise ImportError("the ROHF-form Fock matrix of the spin-adaptation terms comes from pyscf.scf.ROHF(mol).get_veff "
                              "(XTDA.py:607-613); PySCF is not importable -- pass a ProblemData with fock_hf instead") from e
        hf = scf.ROHF(mol)
        if getattr(mf, "with_x2c", None) is not None:
            hf = hf.x2c()
        veff = np.asarray(hf.get_veff(mol, dm))
        h1 = np.asarray(hf.get_hcore())
        fock_hf = np.stack([mo_coeff[s].T @ (h1 + veff[s]) @ mo_coeff[s] for s in (0, 1)])
    is_dft = hasattr(mf, "xc") and hasattr(mf, "_numint")
    omega = alpha = 0.0
    hyb = 1.0
    if is_dft:
        omega, alpha, hyb = mf._numint.rsh_and_hybrid_coeff(mf.xc, mol.spin)
    # ---- density fitting ---------------------------------------------------------------------------------------
    with_df = getattr(mf, "with_df", None)
    if with_df is None:
        if auxbasis is None:
            raise NotDensityFittedError(
                "this mean-field object is not density-fitted: the reference would contract exact 4-centre integrals (mf.get_jk), the "
                "B200 path is DF-only.  Use mf.density_fit(), or pass auxbasis=... to accept the fitting error explicitly")
        from pyscf import df
        with_df = df.DF(mol, auxbasis=auxbasis)
    cderi_packed, naux = _tril_source(with_df)
    cderi_lr_packed = None
    if omega != 0 and alpha != hyb:
        from pyscf import df
        with mol.with_range_coulomb(omega):
            lr = df.DF(mol, auxbasis=getattr(with_df, "auxbasis", auxbasis)).build()
        cderi_lr_packed, naux_lr = _tril_source(lr)
        assert naux_lr == naux
    # ---- grid --------------------------------------------------------------------------------------------------
    ao = weights = fxc_uks = fxc_alda0 = None
    xctype = XC_NONE
    if fxc_mcol is None:
        fxc_mcol = getattr(mf, "fxc_sf_mc", None)      # a multicollinear kernel the caller cached on the object
    if is_dft and kernel != "none":
        xctype, ao, weights, fxc_uks, fxc_alda0, fxc_mcol = _kernel_and_grid(mf, mol, kernel, mo_coeff, mo_occ2, restricted,
                                                                             collinear_samples, fxc_mcol)
    else:
        fxc_mcol = None
    return ProblemData(nao=nao, nc=nc, no=no, nv=nv, restricted=restricted, mo_coeff=mo_coeff, mo_energy=mo_energy, fock_ks=fock_ks,
                       fock_hf=fock_hf, cderi_packed=cderi_packed, cderi_lr_packed=cderi_lr_packed, naux_packed=naux, hyb=hyb, alpha=alpha,
                       omega=omega, xctype=xctype, ao=ao, weights=weights, fxc_uks=fxc_uks, fxc_alda0=fxc_alda0, fxc_mcol=fxc_mcol,
                       level_shift=getattr(mf, "level_shift", 0.0) or 0.0, meta=dict(source="mean-field object"))
