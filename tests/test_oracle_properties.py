"""Pin oracle/properties.py against fixtures produced by the reference's own osc_str / rot_str / deltaS2 /
calculate_TDM_R / calculate_TDM_U / deltaS2_U / analyse (tests/golden/make_golden_properties.py).  CPU only."""
import os

import numpy as np
import pytest

from oracle import properties as oprop
from oracle import layouts
from xtddft_b200.synth import make_problem

from golden.make_golden_properties import one_electron

TOL = 1e-11


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "properties.npz"))


def _close(a, b, tol=TOL):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape
    assert float(np.abs(a - b).max()) <= tol * max(1.0, float(np.abs(b).max()))


def _problem(prm, restricted=True):
    nc, no, nv, seed = [int(x) for x in prm]
    return make_problem(nc + no + nv, nc, no, nv, 6, 0, xctype="HF", hyb=1.0, restricted=restricted, seed=seed), seed


@pytest.mark.parametrize("tag", ["xtda_a", "xtda_b"])
def test_xtda_strengths(g, tag):
    p, seed = _problem(g[f"{tag}_params"])
    dip, ipo, rxp, _ = one_electron(p.nao, seed + 1)
    x1, e = g[f"{tag}_x1"], g[f"{tag}_e"]
    _close(oprop.xtda_osc_str(p, e, x1, dip), g[f"{tag}_os"])
    _close(oprop.xtda_rot_str(p, e, x1, ipo, rxp), g[f"{tag}_rs"])
    v_my = x1.T[layouts.order_pyscf2my(p.nc, p.no, p.nv)]
    nc, no, nv = p.nc, p.no, p.nv
    d = v_my[:nc * nv] - v_my[(nc + no) * nv + nc * no:]
    _close(np.einsum("ik,ik->k", d, d), g[f"{tag}_dS2"])


@pytest.mark.parametrize("tag", ["sf_a", "sf_b"])
def test_sf_oscillator_matrix_roks(g, tag):
    p, seed = _problem(g[f"{tag}_params"])
    dip, _, _, _ = one_electron(p.nao, seed + 1)
    c = p.mo_coeff[0]
    ints_mo = np.einsum("xpq,pi,qj->xij", dip, c, c)
    e = g[f"{tag}_e"]
    vects = layouts.get_vect(p.no)
    for re in (0, 1):
        v = g[f"{tag}_v_re{re}"]
        for X in (0, 1, 3):
            tdm = oprop.tdm_r(v, ints_mo, p.nc, p.no, p.nv, X, vects if re else None)
            _close(oprop.osc_matrix(e, tdm), g[f"{tag}_osc_X{X}_re{re}"])
        _close(oprop.delta_s2_roks_sf(v, p.nc, p.no, p.nv, vects if re else None), g[f"{tag}_ds2_re{re}"])


def test_usf_oscillator_matrix_and_s2(g):
    tag = "usf_a"
    p, seed = _problem(g[f"{tag}_params"], restricted=False)
    dip, _, _, ovlp = one_electron(p.nao, seed + 1)
    ca, cb = p.mo_coeff
    aa = np.einsum("xpq,pi,qj->xij", dip, ca, ca)
    bb = np.einsum("xpq,pi,qj->xij", dip, cb, cb)
    v, e = g[f"{tag}_v"], g[f"{tag}_e"]
    _close(oprop.osc_matrix(e, oprop.tdm_u(v, aa, bb, p.nc, p.no, p.nv)), g[f"{tag}_osc"])
    _close(oprop.delta_s2_u(p, v, ovlp) + p.no - 1, g[f"{tag}_pab"])
