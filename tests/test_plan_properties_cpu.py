"""Operator plans of the property pass (xtddft_b200/properties.py) executed by the NumPy plan interpreter reproduce
the golden fixtures of the reference's calculate_TDM_R / calculate_TDM_U / deltaS2_U.  CPU only: checks the plan
construction (block scalings, trace terms, layouts incl. the removed OO vector) without a GPU."""
import os

import numpy as np
import pytest

from plan_interp import PlanInterpreter
from xtddft_b200 import plan as planmod
from xtddft_b200 import properties as prop
from xtddft_b200.synth import make_problem

from golden.make_golden_properties import one_electron

TOL = 1e-11


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "properties.npz"))


def _close(a, b):
    assert a.shape == b.shape
    assert float(np.abs(a - b).max()) <= TOL * max(1.0, float(np.abs(b).max())), float(np.abs(a - b).max())


def _bilinear(plans, p, v):
    return np.stack([v.T @ PlanInterpreter(pl, p).sigma(v.T).T for pl in plans])


def _problem(prm, restricted=True):
    nc, no, nv, seed = [int(x) for x in prm]
    return make_problem(nc + no + nv, nc, no, nv, 6, 0, xctype="HF", hyb=1.0, restricted=restricted, seed=seed), seed


@pytest.mark.parametrize("tag", ["sf_a", "sf_b"])
def test_tdm_r_plans(g, tag):
    p, seed = _problem(g[f"{tag}_params"])
    dip, _, _, _ = one_electron(p.nao, seed + 1)
    c = p.mo_coeff[0]
    d_mo = np.einsum("xpq,pi,qj->xij", dip, c, c)
    e = g[f"{tag}_e"]
    for re in (0, 1):
        v = g[f"{tag}_v_re{re}"]
        for X in (0, 1, 3):
            plans = [prop.tdm_r_operator(p, d, X, planmod.LAYOUT_BLOCK, bool(re)) for d in d_mo]
            tdm = _bilinear(plans, p, v)
            _close(prop.PropertyPass.osc_matrix(e, tdm), g[f"{tag}_osc_X{X}_re{re}"])


def test_usf_plans(g):
    tag = "usf_a"
    p, seed = _problem(g[f"{tag}_params"], restricted=False)
    dip, _, _, ovlp = one_electron(p.nao, seed + 1)
    ca, cb = p.mo_coeff
    aa = np.einsum("xpq,pi,qj->xij", dip, ca, ca)
    bb = np.einsum("xpq,pi,qj->xij", dip, cb, cb)
    v, e = g[f"{tag}_v"], g[f"{tag}_e"]
    tdm = _bilinear([prop.tdm_u_operator(p, a, b, planmod.LAYOUT_BLOCK) for a, b in zip(aa, bb)], p, v)
    _close(prop.PropertyPass.osc_matrix(e, tdm), g[f"{tag}_osc"])
    s_ba = cb.T @ ovlp @ ca[:, :p.nocc_a]
    q = _bilinear([prop.s2_u_operator(p, s_ba[:p.nocc_b], s_ba[p.nocc_b:], planmod.LAYOUT_BLOCK)], p, v)[0]
    _close(np.diag(q), g[f"{tag}_pab"])
