"""Input container for the sigma-vector path.

`ProblemData` holds exactly what the reference's `gen_vind()` closures capture from the SCF object
(reference: xtddft/XTDA.py:558-613, xtddft/SF_TDA.py:162-221, xtddft/XSF_TDA.py:1029-1121):
MO coefficients / energies, the KS Fock matrix in the MO basis, the ROHF-form Fock matrix of the KS
density, the density-fitting 3-centre tensor, AO values on the integration grid with the cached
exchange-correlation kernel, and the hybrid-exchange scalars.  It is plain NumPy (host side); the
CUDA engine uploads it once per solve.

Conventions (reference: XTDA.py:564-586, SF_TDA.py:26-37):
  * MO order is closed (nc) | open (no) | virtual (nv); nmo = nc+no+nv.
  * alpha occupied = closed+open, alpha virtual = virtual,
    beta  occupied = closed,      beta  virtual = open+virtual.
  * A ROKS reference is mapped to "UKS form" with C_alpha = C_beta.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np

XC_NONE = "HF"     # no grid term (Hartree-Fock / collinear spin-flip)
XC_LDA = "LDA"     # nvar = 1
XC_GGA = "GGA"     # nvar = 4
XC_MGGA = "MGGA"   # AO components 4 (value + gradient, no Laplacian), kernel components 5 (rho, grad rho, tau)

# kinds of cached kernel (xtd_set_fxc)
FXC_NONE = 0
FXC_UKS = 1      # spin-conserving f_xc[2,nvar,2,nvar,Ng]   (XTDA.py:504, numint.cache_xc_kernel)
FXC_ALDA0 = 2    # spin-flip ALDA0 scalar kernel f[Ng]       (SF_TDA.py:39-88), already weighted
FXC_MCOL = 3     # multicollinear spin-flip kernel [nvar,nvar,Ng] (SF_TDA.py:942-974), unweighted


@dataclass
class ProblemData:
    nao: int
    nc: int
    no: int
    nv: int
    restricted: bool                     # True: ROKS reference, False: UKS reference
    mo_coeff: np.ndarray                 # [2, nao, nmo]
    mo_energy: np.ndarray                # [2, nmo]
    fock_ks: np.ndarray                  # [2, nmo, nmo]  C^T (h + veff[s]) C
    fock_hf: Optional[np.ndarray] = None  # [2, nmo, nmo]  ROHF-form Fock of the KS density (ROKS only)
    cderi: Optional[np.ndarray] = None    # [naux, nao, nao] symmetric in the last two indices
    cderi_lr: Optional[np.ndarray] = None  # long-range (erf-attenuated) tensor for range-separated hybrids
    # PySCF's native storage of the same tensors: lower-triangular packed rows [naux, nao(nao+1)/2] (`with_df._cderi`) or a
    # zero-argument callable returning an iterator over such row blocks (`with_df.loop()` for an on-disk tensor).  Streamed to
    # the engine block by block (xtd_df_add(packed=1)); never unpacked on the host.
    cderi_packed: object = None
    cderi_lr_packed: object = None
    naux_packed: int = 0
    hyb: float = 0.0
    alpha: float = 0.0
    omega: float = 0.0
    xctype: str = XC_NONE
    ao: Optional[np.ndarray] = None      # [nvar, ng, nao]  (value, d/dx, d/dy, d/dz)
    weights: Optional[np.ndarray] = None  # [ng]
    fxc_uks: Optional[np.ndarray] = None  # [2, nvar, 2, nvar, ng]  unweighted
    fxc_alda0: Optional[np.ndarray] = None  # [ng]  weighted (SF_TDA.py:82-84)
    fxc_mcol: Optional[np.ndarray] = None  # [nvar, nvar, ng]  unweighted
    level_shift: float = 0.0
    df_external: bool = False             # the 3-centre tensor is streamed to the engine by the caller (device-resident data)
    grid_external: bool = False           # AO values / kernel are given to the engine by the caller
    meta: dict = field(default_factory=dict)

    # ---- derived sizes -------------------------------------------------------------------
    @property
    def nmo(self) -> int:
        return self.nc + self.no + self.nv

    @property
    def nocc_a(self) -> int:
        return self.nc + self.no

    @property
    def nocc_b(self) -> int:
        return self.nc

    @property
    def nvir_a(self) -> int:
        return self.nv

    @property
    def nvir_b(self) -> int:
        return self.no + self.nv

    @property
    def naux(self) -> int:
        if self.cderi is None:
            return int(self.naux_packed)
        return int(self.cderi.shape[0])

    @property
    def ng(self) -> int:
        return 0 if self.ao is None else int(self.ao.shape[1])

    @property
    def nvar(self) -> int:
        return 0 if self.ao is None else int(self.ao.shape[0])

    @property
    def nkern(self) -> int:
        """components of the response density / kernel tables: 1, 4 or 5 (meta-GGA adds tau)"""
        return 5 if self.xctype == XC_MGGA else self.nvar

    @property
    def spin_s(self) -> float:
        """S of the reference state (= no/2)."""
        return 0.5 * self.no

    @property
    def has_df(self) -> bool:
        return self.cderi is not None or self.cderi_packed is not None or self.df_external

    @property
    def has_df_lr(self) -> bool:
        return self.cderi_lr is not None or self.cderi_lr_packed is not None or (self.df_external and self.omega != 0.0)

    @property
    def hybrid(self) -> bool:
        return self.hyb != 0.0 or (self.omega != 0.0 and self.alpha != 0.0)

    def validate(self) -> None:
        nmo = self.nmo
        assert self.mo_coeff.shape == (2, self.nao, nmo), self.mo_coeff.shape
        assert self.mo_coeff.dtype == np.float64
        assert self.mo_energy.shape == (2, nmo)
        assert self.fock_ks.shape == (2, nmo, nmo)
        if self.fock_hf is not None:
            assert self.fock_hf.shape == (2, nmo, nmo)
        if self.cderi is not None:
            assert self.cderi.ndim == 3 and self.cderi.shape[1:] == (self.nao, self.nao)
        if self.xctype != XC_NONE and not self.grid_external:
            assert self.ao is not None and self.weights is not None
            assert self.ao.shape[2] == self.nao and self.ao.shape[1] == self.weights.shape[0]
            assert self.ao.shape[0] == (1 if self.xctype == XC_LDA else 4)
        if self.restricted:
            assert np.array_equal(self.mo_coeff[0], self.mo_coeff[1])
            assert np.array_equal(self.mo_energy[0], self.mo_energy[1])

    # ---- aux / grid sharding: sigma is linear in P and in g -------------------
    def shard(self, rank: int, world: int) -> "ProblemData":
        """This rank's slice: contiguous aux block [p0,p1) and grid batch [g0,g1); small data replicated."""
        from .dist import split_range
        import copy
        out = copy.copy(self)
        if self.cderi is not None:
            p0, p1 = split_range(self.naux, rank, world)
            out.cderi = self.cderi[p0:p1]
            if self.cderi_lr is not None:
                out.cderi_lr = self.cderi_lr[p0:p1]
        if self.ao is not None:
            g0, g1 = split_range(self.ng, rank, world)
            out.ao = self.ao[:, g0:g1]
            out.weights = self.weights[g0:g1]
            if self.fxc_uks is not None:
                out.fxc_uks = self.fxc_uks[..., g0:g1]
            if self.fxc_alda0 is not None:
                out.fxc_alda0 = self.fxc_alda0[g0:g1]
            if self.fxc_mcol is not None:
                out.fxc_mcol = self.fxc_mcol[..., g0:g1]
        out.meta = dict(self.meta, rank=rank, world=world)
        return out
