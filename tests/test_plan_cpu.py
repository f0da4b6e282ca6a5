"""Plans (xtddft_b200/plan.py) executed by the NumPy interpreter reproduce the oracle sigma builders.  CPU only."""
import numpy as np
import pytest

from oracle import sigma
from plan_interp import PlanInterpreter
from xtddft_b200 import plan as planmod
from xtddft_b200.synth import make_problem

RTOL = 1e-11


def _check(sig, ref):
    assert sig.shape == ref.shape
    assert np.abs(sig - ref).max() <= RTOL * max(1.0, np.abs(ref).max())


def _z(seed, x, dim):
    return np.random.default_rng(seed).standard_normal((x, dim))


@pytest.mark.parametrize("xct,hyb", [("GGA", 0.2), ("LDA", 0.0), ("HF", 1.0), ("GGA", 0.0)])
@pytest.mark.parametrize("no", [1, 2, 3])
@pytest.mark.parametrize("restricted", [True, False])
def test_xtda_plan(xct, hyb, no, restricted):
    p = make_problem(8 + no, 3, no, 5, 9, 30, xctype=xct, hyb=hyb, restricted=restricted, seed=50 + no)
    vind, hd = sigma.xtda_gen_vind(p)
    pl = planmod.build_xtda_plan(p)
    it = PlanInterpreter(pl, p)
    z = _z(1, 3, hd.size)
    _check(it.sigma(z), vind(z))
    assert np.abs(pl.hdiag - hd).max() < 1e-14


def test_xtda_plan_rsh():
    p = make_problem(10, 3, 2, 5, 9, 30, xctype="GGA", hyb=0.2, seed=9, omega=0.33, alpha=0.65)
    vind, hd = sigma.xtda_gen_vind(p)
    it = PlanInterpreter(planmod.build_xtda_plan(p), p)
    z = _z(2, 2, hd.size)
    _check(it.sigma(z), vind(z))


@pytest.mark.parametrize("isf", [-1, 1])
@pytest.mark.parametrize("method", [0, 1, 2])
@pytest.mark.parametrize("restricted", [True, False])
@pytest.mark.parametrize("no", [2, 3])
def test_sf_plan(isf, method, restricted, no):
    p = make_problem(8 + no, 3, no, 5, 9, 30, xctype="GGA", hyb=0.5, restricted=restricted, seed=60 + no)
    vind, hd = sigma.sf_gen_vind(p, isf, method)
    pl = planmod.build_sf_plan(p, isf=isf, method=method)
    it = PlanInterpreter(pl, p)
    z = _z(3, 3, hd.size)
    _check(it.sigma(z), vind(z))
    assert np.abs(pl.hdiag - hd).max() < 1e-14


@pytest.mark.parametrize("sa", [0, 1, 2, 3])
@pytest.mark.parametrize("remove", [False, True])
@pytest.mark.parametrize("no,nc", [(2, 3), (3, 2), (2, 4)])
@pytest.mark.parametrize("method", [0, 1])
def test_xsf_block_plan(sa, remove, no, nc, method):
    p = make_problem(nc + no + 5, nc, no, 5, 9, 30, xctype="GGA", hyb=0.4, seed=70 + no)
    vind, hd = sigma.xsf_gen_vind(p, sa=sa, method=method, remove=remove, foo=0.8, fglobal=0.7)
    pl = planmod.build_sf_plan(p, isf=-1, method=method, sa=sa, layout=planmod.LAYOUT_BLOCK, remove=remove, foo=0.8,
                               fglobal=0.7, hdiag_kind="xsf")
    it = PlanInterpreter(pl, p)
    z = _z(4, 2, hd.size)
    _check(it.sigma(z), vind(z))
    co_j = ov_j = None
    if pl.j_blocks:
        co_j, ov_j = it.jblock_diag(0), it.jblock_diag(1)
    assert np.abs(planmod.finish_xsf_hdiag(pl, co_j, ov_j) - hd).max() < 1e-12


@pytest.mark.parametrize("x_level", [0, 3])
@pytest.mark.parametrize("remove", [False, True])
def test_xsf_gpu_order_plan(x_level, remove):
    p = make_problem(10, 3, 2, 5, 9, 30, xctype="GGA", hyb=0.4, seed=80)
    vind, hd = sigma.xsf_gpu_gen_vind(p, x_level=x_level, collinear="mcol", extype=1, remove=remove, foo=0.8, fglobal=0.7)
    pl = planmod.build_sf_plan(p, isf=-1, method=1, sa=x_level, layout=planmod.LAYOUT_PYSCF, remove=remove, foo=0.8,
                               fglobal=0.7, hdiag_kind="gpu")
    it = PlanInterpreter(pl, p)
    z = _z(5, 2, hd.size)
    _check(it.sigma(z), vind(z))
    assert np.abs(pl.hdiag - hd).max() < 1e-13


def test_xsf_lda_hf():
    for xct, hyb in (("LDA", 0.3), ("HF", 1.0)):
        p = make_problem(9, 2, 3, 4, 8, 24, xctype=xct, hyb=hyb, seed=90)
        vind, hd = sigma.xsf_gen_vind(p, sa=3, method=0, remove=True)
        pl = planmod.build_sf_plan(p, isf=-1, method=0, sa=3, layout=planmod.LAYOUT_BLOCK, remove=True, hdiag_kind="xsf")
        it = PlanInterpreter(pl, p)
        z = _z(6, 2, hd.size)
        _check(it.sigma(z), vind(z))


@pytest.mark.parametrize("restricted", [True, False])
def test_meta_gga_plans(restricted):
    """tau component of the kernel tables: X-TDA (UKS kernel), spin-flip multicollinear; ALDA0 has no tau part."""
    p = make_problem(11, 3, 2, 6, 9, 30, xctype="MGGA", hyb=0.2, restricted=restricted, seed=77)
    vind, hd = sigma.xtda_gen_vind(p)
    pl = planmod.build_xtda_plan(p)
    assert pl.xc_kind == "uks_tau"
    z = _z(3, 2, hd.size)
    _check(PlanInterpreter(pl, p).sigma(z), vind(z))
    for method, kind in [(1, "mcol_tau"), (0, "alda0")]:
        vind, hd = sigma.sf_gen_vind(p, -1, method)
        pl = planmod.build_sf_plan(p, isf=-1, method=method)
        assert pl.xc_kind == kind
        z = _z(4, 2, hd.size)
        _check(PlanInterpreter(pl, p).sigma(z), vind(z))
