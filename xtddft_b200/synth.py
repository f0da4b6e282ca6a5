"""Seeded synthetic inputs for the sigma path (generator T0).

PySCF / libcint / libxc are not available where this is developed or benchmarked, so the SCF
quantities the reference's `gen_vind()` closures capture are generated here with the documented
distributions: orthonormal MO coefficients, gapped orbital energies, symmetric Fock matrices, a
positive-semidefinite density-fitted ERI tensor (mu nu|la si) = sum_P L_P,mu nu L_P,la si, Gaussian AO
values (and analytic gradients) on a random grid, and negative-definite-ish f_xc kernels.  Coupling
strengths are scaled so that the response part of A is a perturbation of the orbital-energy gaps and
Davidson converges in a physical number of iterations.
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np

from .problem import ProblemData, XC_GGA, XC_LDA, XC_MGGA, XC_NONE

# BASELINE.json configs -> synthetic shapes.  "method": xtda | sf_down | xsf
CONFIGS = {
    1: dict(name="cfg1-XTDA-OH-svp", method="xtda", nao=19, nc=4, no=1, nv=14, naux=95, ng=24000,
            xctype=XC_GGA, hyb=0.20, nroots=5),
    2: dict(name="cfg2-SF-C2H4-tzvp", method="sf_down", nao=86, nc=7, no=2, nv=77, naux=230, ng=70000,
            xctype=XC_GGA, hyb=0.50, nroots=10),
    3: dict(name="cfg3-XSF-CrIII-tzvp", method="xsf", nao=1725, nc=173, no=3, nv=1549, naux=4300,
            ng=1000000, xctype=XC_GGA, hyb=0.25, nroots=10),
    4: dict(name="cfg4-XTDA-TTM-tzvp", method="xtda", nao=958, nc=136, no=1, nv=821, naux=2275,
            ng=430000, xctype=XC_GGA, hyb=0.20, nroots=20),
    5: dict(name="cfg5-SF-MTTM2-tzvp", method="sf_down", nao=2052, nc=275, no=2, nv=1775, naux=4840,
            ng=1000000, xctype=XC_GGA, hyb=0.50, nroots=10),
}


def _orthonormal(rng: np.random.Generator, n: int) -> np.ndarray:
    q, r = np.linalg.qr(rng.standard_normal((n, n)))
    return q * np.sign(np.diag(r))


def _sym_noise(rng: np.random.Generator, n: int, sigma: float) -> np.ndarray:
    a = rng.standard_normal((n, n)) * sigma
    return 0.5 * (a + a.T)


def orbital_energies(rng, nc, no, nv):
    eo = np.sort(rng.uniform(-1.0, -0.3, nc))
    eop = np.sort(rng.uniform(-0.25, -0.05, no))
    ev = np.sort(rng.uniform(0.05, 2.0, nv))
    return np.concatenate([eo, eop, ev])


def gaussian_ao(rng, ng: int, nao: int, deriv: bool, dtype=np.float64):
    """s-type Gaussians exp(-a|r-A|^2) at random centres, values (+ analytic gradient) on random points.

    Returns ao[nvar, ng, nao] and weights[ng] (uniform(0,1) * V / ng).
    """
    box = max(4.0, 1.2 * nao ** (1.0 / 3.0))
    centres = rng.uniform(-0.5 * box, 0.5 * box, (nao, 3))
    expo = np.exp(rng.uniform(math.log(0.1), math.log(5.0), nao))
    coords = rng.uniform(-0.5 * box - 1.0, 0.5 * box + 1.0, (ng, 3))
    d = coords[:, None, :] - centres[None, :, :]              # [ng, nao, 3]
    r2 = np.einsum("gax,gax->ga", d, d)
    norm = (2.0 * expo / math.pi) ** 0.75
    val = norm * np.exp(-expo * r2)
    nvar = 4 if deriv else 1
    ao = np.empty((nvar, ng, nao), dtype=dtype)
    ao[0] = val
    if deriv:
        for k in range(3):
            ao[1 + k] = -2.0 * expo * d[:, :, k] * val
    vol = (box + 2.0) ** 3
    weights = rng.uniform(0.0, 1.0, ng) * vol / ng
    return ao, weights


def _xc_norm_estimate(ao0: np.ndarray, co: np.ndarray, cv: np.ndarray, d: np.ndarray, iters: int = 12) -> float:
    """Largest eigenvalue of X[(ia),(jb)] = sum_g d_g (phi_i phi_a)(g) (phi_j phi_b)(g), d >= 0, by power iteration."""
    po, pv = ao0 @ co, ao0 @ cv
    x = np.ones((co.shape[1], cv.shape[1])) / math.sqrt(co.shape[1] * cv.shape[1])
    lam = 1.0
    for _ in range(iters):
        t = np.einsum("go,ov,gv->g", po, x, pv, optimize=True) * d
        y = np.einsum("g,go,gv->ov", t, po, pv, optimize=True)
        lam = float(np.linalg.norm(y))
        if lam == 0.0:
            return 1.0
        x = y / lam
    return lam


def make_problem(nao: int, nc: int, no: int, nv: int, naux: int, ng: int, *, xctype: str = XC_GGA,
                 hyb: float = 0.2, restricted: bool = True, seed: int = 0, fxc_kinds=("uks", "alda0", "mcol"),
                 coupling: float = 0.5, xc_strength: float = 0.25, omega: float = 0.0, alpha: float = 0.0,
                 level_shift: float = 0.0) -> ProblemData:
    """Random well-conditioned problem (T0).  nao must equal nc+no+nv (square MO coefficient matrix)."""
    nmo = nc + no + nv
    assert nao == nmo, "synthetic generator uses a square MO coefficient matrix"
    rng = np.random.default_rng(seed)
    ca = _orthonormal(rng, nao)
    if restricted:
        cb = ca
        ea = orbital_energies(rng, nc, no, nv)
        eb = ea.copy()
    else:
        cb = _orthonormal(rng, nao)
        ea = orbital_energies(rng, nc, no, nv)
        eb = orbital_energies(rng, nc, no, nv)
    mo_coeff = np.stack([ca, cb])
    mo_energy = np.stack([ea, eb])
    # KS Fock: orbital energies on the diagonal + small symmetric couplings everywhere (ROKS has non-zero
    # closed-open / open-virtual blocks; the sigma builders read full oo and vv blocks).
    fa = np.diag(ea) + _sym_noise(rng, nmo, 0.02)
    fb = np.diag(eb) + _sym_noise(rng, nmo, 0.02)
    if restricted:
        # open-shell splitting: beta sees the open orbitals as virtual, lift them
        shift = np.zeros(nmo)
        shift[nc:nc + no] = 0.35
        fb = fb + np.diag(shift)
        # a ROKS object carries ONE orbital-energy array (eigenvalues of the effective Fock operator); the
        # reference duplicates it for both spins (SF_TDA.py:32, XTDA.py:566)
        roks_e = 0.5 * (fa.diagonal() + fb.diagonal())
        mo_energy = np.stack([roks_e, roks_e])
    fock_ks = np.stack([fa, fb])
    fock_hf = None
    if restricted:
        fock_hf = np.stack([fa + _sym_noise(rng, nmo, 0.05), fb + _sym_noise(rng, nmo, 0.05)])

    cderi = None
    if naux > 0:
        nocc, nvir = nc + no, no + nv
        c = math.sqrt(coupling / 4.0 / max(1.0, math.sqrt(nocc * nvir / naux)))
        l = rng.standard_normal((naux, nao, nao)) * (c / math.sqrt(naux))
        cderi = 0.5 * (l + l.transpose(0, 2, 1))
    cderi_lr = None
    if omega != 0.0 and naux > 0:
        l = rng.standard_normal((naux, nao, nao)) * (0.5 * c / math.sqrt(naux))
        cderi_lr = 0.5 * (l + l.transpose(0, 2, 1))

    ao = weights = fxc_uks = fxc_alda0 = fxc_mcol = None
    if xctype != XC_NONE and ng > 0:
        ao, weights = gaussian_ao(rng, ng, nao, deriv=(xctype in (XC_GGA, XC_MGGA)))
        nvar = 5 if xctype == XC_MGGA else ao.shape[0]      # kernel components (meta-GGA: rho, grad rho, tau)
        # scale the kernels so that the grid term of A has norm ~ xc_strength (a perturbation of the gaps):
        # power-iteration estimate for a unit kernel on the largest occ x vir block
        est = _xc_norm_estimate(ao[0], ca[:, :nc + no], cb[:, nc:], weights)
        fscale = xc_strength / max(est, 1e-300)
        if "uks" in fxc_kinds:
            f = rng.standard_normal((2 * nvar, 2 * nvar, ng)) * 0.25
            f = 0.5 * (f + f.transpose(1, 0, 2))
            idx = np.arange(2 * nvar)
            f[idx, idx, :] = -np.abs(rng.standard_normal((2 * nvar, ng)))
            gradscale = np.ones(2 * nvar)
            if nvar == 4:
                gradscale[[1, 2, 3, 5, 6, 7]] = 0.15
            if nvar == 5:
                gradscale[[1, 2, 3, 6, 7, 8]] = 0.15
                gradscale[[4, 9]] = 0.3
            f = f * gradscale[:, None, None] * gradscale[None, :, None]
            fxc_uks = (f * fscale).reshape(2, nvar, 2, nvar, ng)
        if "alda0" in fxc_kinds:
            fxc_alda0 = -np.abs(rng.standard_normal(ng)) * fscale * weights
        if "mcol" in fxc_kinds:
            f = rng.standard_normal((nvar, nvar, ng)) * 0.25
            f = 0.5 * (f + f.transpose(1, 0, 2))
            idx = np.arange(nvar)
            f[idx, idx, :] = -np.abs(rng.standard_normal((nvar, ng)))
            if nvar >= 4:
                gs = np.array([1.0, 0.15, 0.15, 0.15, 0.3])[:nvar]
                f = f * gs[:, None, None] * gs[None, :, None]
            fxc_mcol = f * (0.5 * fscale)

    p = ProblemData(nao=nao, nc=nc, no=no, nv=nv, restricted=restricted, mo_coeff=mo_coeff,
                    mo_energy=mo_energy, fock_ks=fock_ks, fock_hf=fock_hf, cderi=cderi, cderi_lr=cderi_lr,
                    hyb=hyb, alpha=alpha, omega=omega, xctype=xctype if ng > 0 else XC_NONE,
                    ao=ao, weights=weights, fxc_uks=fxc_uks, fxc_alda0=fxc_alda0, fxc_mcol=fxc_mcol,
                    level_shift=level_shift, meta=dict(seed=seed, generator="T0"))
    p.validate()
    return p


def make_config(idx: int, scale: float = 1.0, seed: Optional[int] = None, **over) -> ProblemData:
    """Synthetic problem with the shape of BASELINE config `idx` (optionally shrunk by `scale`)."""
    c = dict(CONFIGS[idx])
    c.update(over)
    nc = max(1, int(round(c["nc"] * scale)))
    no = c["no"]
    nv = max(2, int(round(c["nv"] * scale)))
    nao = nc + no + nv
    naux = max(4, int(round(c["naux"] * scale)))
    ng = max(64, int(round(c["ng"] * scale)))
    kinds = {"xtda": ("uks",), "sf_down": ("alda0",), "sf_up": ("alda0",), "xsf": ("alda0",)}[c["method"]]
    p = make_problem(nao, nc, no, nv, naux, ng, xctype=c["xctype"], hyb=c["hyb"], restricted=True,
                     seed=1000 + idx if seed is None else seed, fxc_kinds=kinds)
    p.meta.update(config=idx, name=c["name"], method=c["method"], nroots=c["nroots"], scale=scale)
    return p
