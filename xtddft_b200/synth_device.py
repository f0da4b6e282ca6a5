"""Device-side synthetic inputs at BASELINE sizes: the 3-centre tensor (up to ~160 GB unpacked) and
the AO values on the grid are generated block by block on the GPU with per-block seeds and streamed straight
into the engine, so they never exist on the host and any aux / grid sharding sees the same global data.

Torch is used for buffer ownership and the counter-based RNG only.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import numpy as np

from .dist import split_range
from .problem import ProblemData, XC_GGA, XC_LDA, XC_NONE
from .synth import CONFIGS, _orthonormal, _sym_noise, orbital_energies

AUX_BLOCK = 16        # aux functions per RNG block
GRID_BLOCK = 8192     # grid points per RNG block


@dataclass
class DeviceProblem:
    p: ProblemData            # small host data (orbitals, Fock, scalars); tensor and grid are external
    naux: int
    ng: int
    nvar: int                 # AO components held on the grid (1: values only, 4: values + gradient)
    fxc_kind: str             # none | uks | alda0 | mcol
    seed: int
    l_scale: float
    ao_amp: float
    f_scale: float
    method: str
    nroots: int
    name: str

    # ---- streamed pieces ----------------------------------------------------------------------------
    def cderi_block(self, torch, device, b: int):
        """aux functions [b*AUX_BLOCK, (b+1)*AUX_BLOCK) of the global tensor, symmetric in (mu, nu)."""
        n = self.p.nao
        n_in = min(AUX_BLOCK, self.naux - b * AUX_BLOCK)
        g = torch.Generator(device=device)
        g.manual_seed(self.seed * 1000003 + 17 + b)
        l = torch.randn((AUX_BLOCK, n, n), generator=g, device=device, dtype=torch.float64)[:n_in]
        l = (l + l.transpose(1, 2)) * (0.5 * self.l_scale)
        return l

    def stream_cderi(self, eng, tensor: int = 0, rank: int = 0, world: int = 1):
        torch = eng.torch
        p0, p1 = split_range(self.naux, rank, world)
        eng.df_begin(tensor, p1 - p0)
        for b in range(p0 // AUX_BLOCK, (p1 + AUX_BLOCK - 1) // AUX_BLOCK):
            blk = self.cderi_block(torch, eng.device, b)
            lo = max(p0 - b * AUX_BLOCK, 0)
            hi = min(p1 - b * AUX_BLOCK, blk.shape[0])
            if hi > lo:
                eng.df_add(tensor, blk[lo:hi].contiguous() if (lo, hi) != (0, blk.shape[0]) else blk)
            del blk

    def grid_block(self, torch, device, b: int, ld: int):
        n = self.p.nao
        n_in = min(GRID_BLOCK, self.ng - b * GRID_BLOCK)
        g = torch.Generator(device=device)
        g.manual_seed(self.seed * 7000003 + 31 + b)
        ao = torch.randn((self.nvar, GRID_BLOCK, n), generator=g, device=device, dtype=torch.float64)[:, :n_in] * self.ao_amp
        if self.nvar == 4:
            ao[1:] *= 0.5
        w = torch.rand((GRID_BLOCK,), generator=g, device=device, dtype=torch.float64)[:n_in] * (1000.0 / self.ng)
        r = torch.randn((GRID_BLOCK,), generator=g, device=device, dtype=torch.float64)[:n_in]
        return ao, w, r

    def make_grid(self, eng, rank: int = 0, world: int = 1):
        """Allocate this rank's ao[nvar, ng_loc, ld], weights and kernel on the device and hand them to the engine."""
        torch = eng.torch
        if self.fxc_kind == "none":
            return
        g0, g1 = split_range(self.ng, rank, world)
        if g1 == g0:
            return
        n = self.p.nao
        ld = (n + 15) // 16 * 16
        ao = torch.zeros((self.nvar, g1 - g0, ld), dtype=torch.float64, device=eng.device)
        w = torch.empty((g1 - g0,), dtype=torch.float64, device=eng.device)
        rr = torch.empty((g1 - g0,), dtype=torch.float64, device=eng.device)
        for b in range(g0 // GRID_BLOCK, (g1 + GRID_BLOCK - 1) // GRID_BLOCK):
            a_b, w_b, r_b = self.grid_block(torch, eng.device, b, ld)
            lo = max(g0 - b * GRID_BLOCK, 0)
            hi = min(g1 - b * GRID_BLOCK, a_b.shape[1])
            d0 = b * GRID_BLOCK + lo - g0
            ao[:, d0:d0 + hi - lo, :n] = a_b[:, lo:hi]
            w[d0:d0 + hi - lo] = w_b[lo:hi]
            rr[d0:d0 + hi - lo] = r_b[lo:hi]
            del a_b
        eng.set_grid(ao, w)
        if self.fxc_kind == "alda0":
            f = -(rr.abs()) * (1.5 * self.f_scale) * w
        elif self.fxc_kind == "mcol":
            nv = self.nvar
            f = torch.zeros((nv, nv, g1 - g0), dtype=torch.float64, device=eng.device)
            for c in range(nv):
                f[c, c] = -(rr.abs()) * (0.75 * self.f_scale) * (1.0 if c == 0 else 0.09)
        elif self.fxc_kind == "uks":
            nv = self.nvar
            f = torch.zeros((2, nv, 2, nv, g1 - g0), dtype=torch.float64, device=eng.device)
            for s in range(2):
                for c in range(nv):
                    f[s, c, s, c] = -(rr.abs()) * self.f_scale * (1.0 if c == 0 else 0.09)
            f[0, 0, 1, 0] = f[1, 0, 0, 0] = -(rr.abs()) * (0.25 * self.f_scale)
        else:
            raise ValueError(self.fxc_kind)
        eng.set_fxc(self.fxc_kind, f.contiguous())


def molecular_orbital_energies(rng, nc: int, no: int, nv: int) -> np.ndarray:
    """Orbital-energy ladder shaped like a large organic open-shell molecule in a triple-zeta basis: ~30 % core-like
    closed shells far below the valence band, a valence band up to the HOMO region, singly occupied orbitals in the
    gap, a sparse set of low-lying virtuals and the bulk of the virtuals spread over several Hartree."""
    ncore = int(0.3 * nc)
    eo = np.sort(np.concatenate([rng.uniform(-20.0, -1.5, ncore), rng.uniform(-1.2, -0.28, nc - ncore)]))
    eop = np.sort(rng.uniform(-0.22, -0.12, no))
    nlow = max(2, int(0.05 * nv))
    ev = np.sort(np.concatenate([rng.uniform(0.02, 0.3, nlow), rng.uniform(0.3, 6.0, nv - nlow)]))
    return np.concatenate([eo, eop, ev])


def make_device_problem(cfg: int, scale: float = 1.0, seed: Optional[int] = None, grid_components: Optional[int] = None,
                        spectrum: str = "molecular", open_shift: float = 0.1, coupling: float = 0.2, fock_noise: float = 0.005,
                        xc_scale: float = 1.0, hf_exchange: Optional[float] = None, **over) -> DeviceProblem:
    """Host-side small data + scales for BASELINE config `cfg` (optionally shrunk by `scale`).

    spectrum: "molecular" (default, see molecular_orbital_energies) or "uniform" (the SURVEY 8d T0 ladder U(-1,-0.3) /
    U(0.05,2); with open_shift=0.35, coupling=0.5, fock_noise=0.02 it reproduces the round-1 first-bench inputs).  The sigma
    build costs the same for any of them; what changes is how many Davidson cycles the synthetic spectrum needs: the
    uniform ladder packs the ten lowest roots of config 5 into a few mHartree and needed 168 cycles / 1382 sigma vectors
    (profiles/bench_cfg5_n1_r01_uniform_ladder.json), far from what a molecule shows."""
    c = dict(CONFIGS[cfg])
    c.update(over)
    seed = 1000 + cfg if seed is None else seed
    nc = max(1, int(round(c["nc"] * scale)))
    no = c["no"]
    nv = max(2, int(round(c["nv"] * scale)))
    nao = nc + no + nv
    naux = max(4, int(round(c["naux"] * scale)))
    ng = max(256, int(round(c["ng"] * scale)))
    rng = np.random.default_rng(seed)
    ca = _orthonormal(rng, nao)
    ea = orbital_energies(rng, nc, no, nv) if spectrum == "uniform" else molecular_orbital_energies(rng, nc, no, nv)
    nmo = nao
    fa = np.diag(ea) + _sym_noise(rng, nmo, fock_noise)
    fb = np.diag(ea) + _sym_noise(rng, nmo, fock_noise)
    shift = np.zeros(nmo)
    shift[nc:nc + no] = open_shift
    fb = fb + np.diag(shift)
    roks_e = 0.5 * (fa.diagonal() + fb.diagonal())
    # ROHF-form Fock matrices (XTDA.py:588-613, XSF_TDA.py:1096-1121).  Their spin difference F^b - F^a is the exchange
    # field of the open shells: positive semidefinite, a few tenths of a Hartree, dominated by a handful of directions --
    # not a dense random matrix (whose norm would grow like sqrt(nmo) and swamp the orbital gaps).
    hf_a = fa + _sym_noise(rng, nmo, fock_noise)
    nk = no + 3
    q, _ = np.linalg.qr(rng.standard_normal((nmo, nk)))
    lam = rng.uniform(0.03, 0.15, nk) if hf_exchange is None else np.full(nk, hf_exchange)
    fock_hf = np.stack([hf_a, hf_a + np.diag(shift) + (q * lam) @ q.T])
    method = c["method"]
    fxc_kind = {"xtda": "uks", "sf_down": "alda0", "sf_up": "alda0", "xsf": "alda0"}[method]
    # ALDA0 needs only the density component of the AO values (SF_TDA.py:114-116)
    nvar = grid_components if grid_components is not None else (1 if fxc_kind == "alda0" else 4)
    p = ProblemData(nao=nao, nc=nc, no=no, nv=nv, restricted=True, mo_coeff=np.stack([ca, ca]), mo_energy=np.stack([roks_e, roks_e]),
                    fock_ks=np.stack([fa, fb]), fock_hf=fock_hf, hyb=c["hyb"], xctype=XC_GGA if nvar == 4 else XC_LDA,
                    df_external=True, grid_external=True, meta=dict(config=cfg, name=c["name"], method=method, scale=scale,
                              generator=f"seeded device RNG; {spectrum} orbital ladder, open_shift={open_shift}, coupling={coupling}, "
                                        f"fock_noise={fock_noise}"))
    nocc, nvir = nc + no, no + nv
    cc = math.sqrt(coupling / 4.0 / max(1.0, math.sqrt(nocc * nvir / naux)))
    ao_amp = 0.2
    f_scale = xc_scale * 0.1 / (500.0 * ao_amp ** 4 * (1.0 + math.sqrt(nocc * nvir / ng)) ** 2)
    return DeviceProblem(p=p, naux=naux, ng=ng, nvar=nvar, fxc_kind=fxc_kind, seed=seed, l_scale=cc / math.sqrt(naux), ao_amp=ao_amp,
                         f_scale=f_scale, method=method, nroots=c["nroots"], name=c["name"] + (f"@{scale:g}" if scale != 1.0 else ""))


def host_sample(dp: DeviceProblem, naux_s: int, ng_s: int) -> ProblemData:
    """A host ProblemData with the same small data and a SAMPLE of the tensor / grid (same distributions), for the
    bounded CPU baseline: sigma is linear in the aux functions and grid points, so timings extrapolate linearly."""
    import copy
    p = copy.copy(dp.p)
    rng = np.random.default_rng(dp.seed + 99)
    n = p.nao
    l = rng.standard_normal((naux_s, n, n)) * dp.l_scale
    p.cderi = 0.5 * (l + l.transpose(0, 2, 1))
    p.df_external = False
    if dp.fxc_kind != "none":
        ao = rng.standard_normal((dp.nvar, ng_s, n)) * dp.ao_amp
        w = rng.uniform(0, 1, ng_s) * (1000.0 / dp.ng)
        r = np.abs(rng.standard_normal(ng_s))
        p.ao, p.weights = ao, w
        p.grid_external = False
        if dp.fxc_kind == "alda0":
            p.fxc_alda0 = -r * 1.5 * dp.f_scale * w
        elif dp.fxc_kind == "uks":
            f = np.zeros((2, dp.nvar, 2, dp.nvar, ng_s))
            for s in range(2):
                for c in range(dp.nvar):
                    f[s, c, s, c] = -r * dp.f_scale * (1.0 if c == 0 else 0.09)
            f[0, 0, 1, 0] = f[1, 0, 0, 0] = -r * 0.25 * dp.f_scale
            p.fxc_uks = f
        elif dp.fxc_kind == "mcol":
            f = np.zeros((dp.nvar, dp.nvar, ng_s))
            for c in range(dp.nvar):
                f[c, c] = -r * 0.75 * dp.f_scale * (1.0 if c == 0 else 0.09)
            p.fxc_mcol = f
    else:
        p.xctype = XC_NONE
    return p
