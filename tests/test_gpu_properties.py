"""Device property pass (xtddft_b200/properties.py) against the golden fixtures produced by the
reference's own osc_str / rot_str / calculate_TDM_R / calculate_TDM_U / deltaS2_U (tests/golden/properties.npz) and
against the oracle on larger seeded cases.  Needs a B200."""
import os

import numpy as np
import pytest

from oracle import layouts
from oracle import properties as oprop
from xtddft_b200 import plan as planmod
from xtddft_b200.synth import make_problem

from golden.make_golden_properties import one_electron, orthonormal_states

pytestmark = pytest.mark.gpu
RTOL = 1e-9


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "properties.npz"))


def _close(a, b, tol=RTOL):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape
    assert float(np.abs(a - b).max()) <= tol * max(1.0, float(np.abs(b).max())), float(np.abs(a - b).max())


def _problem(prm, restricted=True):
    nc, no, nv, seed = [int(x) for x in prm]
    return make_problem(nc + no + nv, nc, no, nv, 6, 0, xctype="HF", hyb=1.0, restricted=restricted, seed=seed), seed


@pytest.mark.parametrize("tag", ["xtda_a", "xtda_b"])
def test_xtda_strengths_golden(torch_cuda, g, tag):
    from xtddft_b200.properties import PropertyPass
    p, seed = _problem(g[f"{tag}_params"])
    dip, ipo, rxp, _ = one_electron(p.nao, seed + 1)
    pp = PropertyPass(p)
    _close(pp.xtda_osc_str(g[f"{tag}_e"], g[f"{tag}_x1"], dip), g[f"{tag}_os"])
    _close(pp.xtda_rot_str(g[f"{tag}_e"], g[f"{tag}_x1"], ipo, rxp), g[f"{tag}_rs"])


@pytest.mark.parametrize("tag", ["sf_a", "sf_b"])
def test_sf_oscillator_matrix_golden(torch_cuda, g, tag):
    from xtddft_b200.properties import PropertyPass
    p, seed = _problem(g[f"{tag}_params"])
    dip, _, _, _ = one_electron(p.nao, seed + 1)
    pp = PropertyPass(p)
    e = g[f"{tag}_e"]
    for re in (0, 1):
        v = g[f"{tag}_v_re{re}"]
        for X in (0, 1, 3):
            tdm = pp.tdm_r(v.T, dip, X, planmod.LAYOUT_BLOCK, bool(re))
            _close(pp.osc_matrix(e, tdm), g[f"{tag}_osc_X{X}_re{re}"])


def test_usf_golden(torch_cuda, g):
    from xtddft_b200.properties import PropertyPass
    tag = "usf_a"
    p, seed = _problem(g[f"{tag}_params"], restricted=False)
    dip, _, _, ovlp = one_electron(p.nao, seed + 1)
    pp = PropertyPass(p)
    v, e = g[f"{tag}_v"], g[f"{tag}_e"]
    _close(pp.osc_matrix(e, pp.tdm_u(v.T, dip)), g[f"{tag}_osc"])
    _close(pp.delta_s2_u(v.T, ovlp) + p.no - 1, g[f"{tag}_pab"])


@pytest.mark.parametrize("sa,remove", [(0, False), (2, True), (3, True)])
def test_tdm_medium_vs_oracle(torch_cuda, sa, remove):
    """sizes off the 128 tile, odd block sizes, 7 states; PySCF-order rows as the GPU class hands them over"""
    from xtddft_b200.properties import PropertyPass
    nc, no, nv = 37, 3, 141
    p = make_problem(nc + no + nv, nc, no, nv, 5, 0, xctype="HF", hyb=1.0, seed=300 + sa)
    dip, ipo, rxp, _ = one_electron(p.nao, 301)
    ns = 7
    dimf = (nc + no) * (no + nv)
    v = orthonormal_states(dimf - int(remove), ns, 302)                      # block order columns
    c = p.mo_coeff[0]
    ints_mo = np.einsum("xpq,pi,qj->xij", dip, c, c)
    vects = layouts.get_vect(no) if remove else None
    ref = oprop.tdm_r(v, ints_mo, nc, no, nv, sa, vects)
    pp = PropertyPass(p)
    _close(pp.tdm_r(v.T, dip, sa, planmod.LAYOUT_BLOCK, remove), ref)
    # X-TDA moments on the same orbitals
    dimx = (nc + no) * nv + nc * (no + nv)
    x1 = orthonormal_states(dimx, ns, 303).T
    e = np.linspace(0.1, 0.4, ns)
    _close(pp.xtda_osc_str(e, x1, dip), oprop.xtda_osc_str(p, e, x1, dip))
    _close(pp.xtda_rot_str(e, x1, ipo, rxp), oprop.xtda_rot_str(p, e, x1, ipo, rxp))
