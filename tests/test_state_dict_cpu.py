"""Spin-flip-down vectors in the layout the reference's state-interaction driver consumes (xtddft_b200/state_dict.py)
against a loop restatement of x2c_hamiltonian/test_SOCSI.py:47-58.  CPU only."""
import types

import numpy as np
import pytest

from oracle import layouts
from xtddft_b200 import state_dict as sd


def _restated(xm_, vects, nc, no, nv):
    """test_SOCSI.py:47-58 with explicit loops"""
    dim = nc * nv + nc * no + no * nv
    ns = xm_.shape[1]
    xm = np.zeros((dim + no ** 2 + no, ns))
    xm[:dim, :] = xm_[:dim, :]
    for n in range(ns):
        xo = (vects @ xm_[dim:, n]).reshape(no, no) if vects is not None else xm_[dim:, n].reshape(no, no)
        for i in range(no):
            for j in range(no):
                xm[dim + i * no + j, n] = 0.0 if i == j else xo[i, j]
            xm[dim + no * no + i, n] = xo[i, i]
    return xm


@pytest.mark.parametrize("no", [2, 3])
@pytest.mark.parametrize("remove", [False, True])
def test_si_vector_layout(no, remove):
    nc, nv = 3, 4
    dim = (nc + no) * (no + nv) - int(remove)
    v = np.random.default_rng(no).standard_normal((dim, 5))
    vects = layouts.get_vect(no) if remove else None
    got = sd.xsf_si_vectors(v, nc, no, nv, vects)
    assert np.abs(got - _restated(v, vects, nc, no, nv)).max() < 1e-15     # (matrix-vector vs matrix-matrix summation order)
    # the OO part recombines to the expanded OO block
    d3 = nc * nv + nc * no + no * nv
    oo = (vects @ v[d3:]) if remove else v[d3:]
    rec = got[d3:d3 + no * no].reshape(no, no, -1).copy()
    rec[np.arange(no), np.arange(no)] += got[d3 + no * no:]
    assert np.abs(rec.reshape(no * no, -1) - oo).max() < 1e-15


def test_build_state_dict():
    nc, no, nv = 2, 2, 3
    r = np.random.default_rng(0)
    xsf = types.SimpleNamespace(v=r.standard_normal(((nc + no) * (no + nv) - 1, 3)), e=np.array([0.1, 0.2, 0.3]), nc=nc, no=no, nv=nv,
                                re=True, vects=layouts.get_vect(no))
    xt = types.SimpleNamespace(v=r.standard_normal((10, 2)), e=np.array([0.15, 0.25]))
    up = types.SimpleNamespace(v=r.standard_normal((6, 4)), e=np.array([0.3, 0.4, 0.5, 0.6]), nstates=2)
    st = sd.build_state_dict(xsf, xt, up)
    assert [len(st[k]) for k in ("|S->", "|So>", "|S+>")] == [3, 2, 2]
    assert st["|S->"][1][0] == 0.2 and st["|S->"][1][1].shape == (nc * nv + nc * no + no * nv + no * no + no,)


@pytest.mark.parametrize("nc,no,nv", [(3, 2, 4), (2, 3, 3)])
def test_si_vector_layout_against_the_reference(golden_dir, nc, no, nv):
    """Fixture produced by executing the reference's OWN statements (x2c_hamiltonian/test_SOCSI.py:47-58, read from the reference
    tree by tests/golden/make_golden_properties.py) on a stand-in solved XSF_TDA object with the removed OO vector."""
    import os
    d = np.load(os.path.join(golden_dir, "state_dict.npz"), allow_pickle=False)
    got = sd.xsf_si_vectors(d[f"in_{nc}_{no}_{nv}"], nc, no, nv, layouts.get_vect(no))
    assert np.abs(got - d[f"out_{nc}_{no}_{nv}"]).max() < 1e-15
