"""Wire format handed to the reference's spin-orbit state-interaction driver.

`x2c_hamiltonian/driver/si_driver.py:75-103` consumes `states = {'|So>': [...], '|S+>': [...], '|S->': [...]}`, each value a
list of `(energy[Ha], X)` tuples; `x2c_hamiltonian/test_SOCSI.py:35-103` shows how the three manifolds are filled from the
XSF_TDA / XTDA / SF_TDA(isf=1) drivers.  The spin-flip-down vector is re-laid out there (`:47-58`): the removed OO vector is
expanded with `vects`, then the OO block is split into its off-diagonal part (no^2 entries, diagonal zeroed) followed by
its diagonal (no entries).  This module does that re-layout for the drop-in drivers of this package; it is host-side
index work on [dim, nstates] arrays.
"""
from __future__ import annotations

from typing import Optional

import numpy as np


def xsf_si_vectors(v: np.ndarray, nc: int, no: int, nv: int, vects: Optional[np.ndarray] = None) -> np.ndarray:
    """Block-order spin-flip-down vectors [dim(-1), nstates] -> [nc*nv + nc*no + no*nv + no^2 + no, nstates] with the OO
    block as (off-diagonal elements with a zeroed diagonal | diagonal) -- test_SOCSI.py:47-58."""
    v = np.asarray(v, dtype=np.float64)
    dim = nc * nv + nc * no + no * nv
    ns = v.shape[1]
    oo = v[dim:] if vects is None else vects @ v[dim:]
    oo = oo.reshape(no, no, ns)
    diag = oo[np.arange(no), np.arange(no), :].copy()            # [no, ns]
    off = oo.copy()
    off[np.arange(no), np.arange(no), :] = 0.0
    out = np.empty((dim + no * no + no, ns))
    out[:dim] = v[:dim]
    out[dim:dim + no * no] = off.reshape(no * no, ns)
    out[dim + no * no:] = diag
    return out


def build_state_dict(xsf=None, xtda=None, sf_up=None) -> dict:
    """{'|S->': [(e, x)], '|So>': [(e, x)], '|S+>': [(e, x)]} from solved driver objects (any may be None).
    Energies in Hartree; `xtda.e` of this package is already in Hartree (the reference's tensor-basis class returns eV and
    test_SOCSI.py:70 divides by ha2eV)."""
    states = {"|So>": [], "|S+>": [], "|S->": []}
    if xsf is not None:
        xm = xsf_si_vectors(xsf.v, xsf.nc, xsf.no, xsf.nv, xsf.vects if getattr(xsf, "re", False) else None)
        states["|S->"] = [(float(e), xm[:, i]) for i, e in enumerate(np.asarray(xsf.e))]
    if xtda is not None:
        states["|So>"] = [(float(e), np.asarray(xtda.v)[:, i]) for i, e in enumerate(np.asarray(xtda.e))]
    if sf_up is not None:
        n = sf_up.nstates
        states["|S+>"] = [(float(e), np.asarray(sf_up.v)[:, i]) for i, e in enumerate(np.asarray(sf_up.e)[:n])]
    return states
