"""Driver classes + device Davidson on the B200 against the oracle operator solved with the oracle's restatement of
the reference solver: excitation energies within 1e-6 Eh, same state ordering, same Delta<S^2> labels."""
import numpy as np
import pytest

from oracle import davidson as odav
from oracle import layouts as olay
from oracle import sigma as osig
from xtddft_b200 import davidson as pdav
from xtddft_b200.adapters import SynthMF
from xtddft_b200.synth import make_problem

pytestmark = pytest.mark.gpu
E_TOL = 1e-6


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def _oracle_solve(vind, hdiag, x0, settings, nroots):
    cfg = pdav.SOLVER[settings]
    precond = odav.make_diag_precond(hdiag.copy(), cfg["level_shift"])
    pick = pdav.pick_positive if cfg["pick_positive"] else None
    return odav.davidson1(lambda xs: np.asarray(vind(xs)), x0, precond, tol=cfg["tol"], tol_residual=cfg["tol_residual"],
                          lindep=cfg["lindep"], max_cycle=cfg["max_cycle"], nroots=nroots, pick=pick)


def test_xtda_driver(torch_cuda):
    from xtddft_b200.XTDA import XTDA
    p = make_problem(34, 8, 1, 25, 30, 500, xctype="GGA", hyb=0.2, seed=200)
    mf = SynthMF(p)
    td = XTDA(mf.mol, mf, nstates=5)
    e, v = td.kernel()
    assert td.converged.all()
    vind, hd = osig.xtda_gen_vind(p)
    conv, e_ref, x_ref, _ = _oracle_solve(vind, hd, td.get_init_guess(mf, 5), "xtda", 5)
    assert conv.all()
    assert np.abs(e - e_ref).max() < E_TOL
    v_ref = np.array(x_ref).T[olay.order_pyscf2my(p.nc, p.no, p.nv)]
    ds2_ref = olay.delta_s2_xtda(v_ref, p.nc, p.no, p.nv)
    assert np.abs(td.dS2 - ds2_ref).max() < 1e-5
    assert v.shape == (hd.size, 5)
    # the operator interface itself
    vind_gpu, hd_gpu = td.gen_vind()
    z = np.random.default_rng(0).standard_normal((2, hd.size))
    assert np.abs(vind_gpu(z) - vind(z)).max() < 1e-10
    assert np.abs(hd_gpu - hd).max() < 1e-13


def test_xtda_gpu_class(torch_cuda):
    from xtddft_b200.XTDA_GPU import XTDA
    p = make_problem(30, 7, 2, 21, 26, 400, xctype="LDA", hyb=0.25, restricted=False, seed=201)
    mf = SynthMF(p)
    td = XTDA(mf.mol, mf, nstates=4)
    e, v = td.kernel()
    vind, hd = osig.xtda_gen_vind(p)
    conv, e_ref, _, _ = _oracle_solve(vind, hd, td.get_init_guess(mf, 4), "gpu_class", 4)
    assert np.abs(e - e_ref).max() < E_TOL


@pytest.mark.parametrize("isf", [-1, 1])
def test_sf_driver(torch_cuda, isf):
    from xtddft_b200.SF_TDA import SF_TDA
    p = make_problem(32, 6, 2, 24, 28, 450, xctype="GGA", hyb=0.5, seed=202)
    td = SF_TDA(SynthMF(p), isf=isf, method=0)
    e_ev, v = td.kernel(nstates=4)
    vind, hd = osig.sf_gen_vind(p, isf, 0)
    x0 = pdav.init_guess(hd, 4, 1e-5)
    conv, e_ref, x_ref, _ = _oracle_solve(vind, hd, x0, "sf_down", 4)
    assert conv.all() and td.converged.all()
    assert np.abs(td.e - e_ref).max() < E_TOL
    assert np.abs(e_ev - e_ref * 27.2113834).max() < 1e-4
    if isf == -1:
        v_ref = olay.deal_v_davidson(np.array(x_ref).T, p.nc, p.no, p.nv)
        ds2 = olay.delta_s2_sf(v_ref, p.nc, p.no, p.nv)
        assert np.abs(td.deltaS2() - ds2).max() < 1e-4


@pytest.mark.parametrize("remove", [True, False])
@pytest.mark.parametrize("sa", [3, 1])
def test_xsf_driver(torch_cuda, remove, sa):
    from xtddft_b200.XSF_TDA import XSF_TDA
    p = make_problem(30, 6, 2, 22, 26, 420, xctype="GGA", hyb=0.5, seed=203)
    td = XSF_TDA(SynthMF(p), SA=sa)
    e_ev, v = td.kernel(nstates=4, remove=remove)
    fg = 0.7 * 0.5 + 0.3
    vind, hd = osig.xsf_gen_vind(p, sa=sa, method=0, remove=remove, foo=1.0, fglobal=fg)
    conv, e_ref, x_ref, _ = _oracle_solve(vind, hd, pdav.init_guess(hd, 4, 1e-5), "xsf", 4)
    assert conv.all() and td.converged.all()
    assert np.abs(td.e - e_ref).max() < E_TOL
    assert np.abs(e_ev - e_ref * 27.21138505).max() < 1e-4
    ds2 = olay.delta_s2_sf(np.array(x_ref).T, p.nc, p.no, p.nv, olay.get_vect(p.no) if remove else None)
    assert np.abs(td.deltaS2() - ds2).max() < 1e-4


def test_xsf_gpu_class(torch_cuda):
    from xtddft_b200.XSF_TDA_GPU import XSF_TDA_GPU
    p = make_problem(30, 6, 2, 22, 26, 420, xctype="GGA", hyb=0.5, seed=204)
    td = XSF_TDA_GPU(SynthMF(p), X=3, collinear="mcol", nstates=4, extype=1, fglobal=0.6, foo=0.9)
    e_ev, v = td.kernel()
    vind, hd = osig.xsf_gpu_gen_vind(p, x_level=3, collinear="mcol", extype=1, remove=True, foo=0.9, fglobal=0.6)
    conv, e_ref, x_ref, _ = _oracle_solve(vind, hd, td.init_guess(), "gpu_class", 4)
    assert np.abs(td.e - e_ref).max() < E_TOL
    assert v.shape[0] == hd.size


def test_device_vector_kernels(torch_cuda):
    """xtd_vec_* primitives against NumPy."""
    torch = torch_cuda
    vb = pdav.CudaVectors(5003)
    rng = np.random.default_rng(1)
    a, b = rng.standard_normal((4, 5003)), rng.standard_normal((7, 5003))
    A, B = vb.from_host(a), vb.from_host(b)
    assert np.abs(vb.dots(A, B) - a @ b.T).max() < 1e-11
    c = rng.standard_normal((4, 7))
    Y = vb.from_host(a)
    vb.lincomb(Y, B, c, beta=0.5)
    assert np.abs(vb.to_host(Y) - (0.5 * a + c @ b)).max() < 1e-12
    e = rng.standard_normal(4)
    R = vb.alloc(4)
    n2 = vb.residual(R, vb.from_host(b[:4]), A, e)
    r = b[:4] - e[:, None] * a
    assert np.abs(vb.to_host(R) - r).max() < 1e-13 and np.abs(n2 - (r * r).sum(1)).max() < 1e-10
    hd = rng.uniform(0.1, 2, 5003); hd[5] = 0.3
    sh = np.array([0.3, 0.1, -0.2, 0.0])
    X = vb.from_host(a)
    n2 = vb.precond(X, vb.from_host(hd[None]), sh)
    d = hd[None] - sh[:, None]; d[abs(d) < 1e-8] = 1e-8
    assert np.abs(vb.to_host(X) - a / d).max() < 1e-7 * np.abs(a / d).max()


def test_driver_property_pass(torch_cuda):
    """kernel() leaves the reference's strengths on the object: XTDA.os / .rs (XTDA.py:816-821) and the state-to-state
    oscillator matrix XSF_TDA_GPU.os (XSF_TDA_GPU.py:1263), evaluated on the device and checked against the oracle."""
    from golden.make_golden_properties import one_electron
    from oracle import properties as oprop
    from xtddft_b200.XSF_TDA import XSF_TDA
    from xtddft_b200.XSF_TDA_GPU import XSF_TDA_GPU
    from xtddft_b200.XTDA import XTDA
    p = make_problem(30, 6, 2, 22, 26, 420, xctype="GGA", hyb=0.5, seed=205)
    dip, ipo, rxp, _ = one_electron(p.nao, 206)
    p.meta["one_electron"] = {"int1e_r": dip, "int1e_ipovlp": ipo, "int1e_cg_irxp": rxp}
    p.meta["chiral"] = True
    mf = SynthMF(p)
    td = XTDA(mf.mol, mf, nstates=4)
    td.kernel()
    x_rows = td._x_rows
    assert np.abs(td.os - oprop.xtda_osc_str(p, td.e, x_rows, dip)).max() < 1e-9 * max(1.0, np.abs(td.os).max())
    assert np.abs(td.rs - oprop.xtda_rot_str(p, td.e, x_rows, ipo, rxp)).max() < 1e-9 * max(1.0, np.abs(td.rs).max())
    c = p.mo_coeff[0]
    ints_mo = np.einsum("xpq,pi,qj->xij", dip, c, c)
    g = XSF_TDA_GPU(mf, X=3, collinear="alda0", nstates=4, extype=1)
    g.kernel()
    ref = oprop.osc_matrix(g.e, oprop.tdm_r(np.asarray(g.v), ints_mo, p.nc, p.no, p.nv, 3, olay.get_vect(p.no)))
    assert g.os.shape == (4, 4) and np.abs(g.os - ref).max() < 1e-9 * max(1.0, np.abs(ref).max())
    x = XSF_TDA(mf, SA=2)
    x.kernel(nstates=3, remove=False)
    tdm, osc = x.calculate_TDM(verbose=False)
    ref = oprop.tdm_r(np.asarray(x.v), ints_mo, p.nc, p.no, p.nv, 2, None)
    assert np.abs(tdm - ref).max() < 1e-9 * max(1.0, np.abs(ref).max())


def test_timing_categories_of_the_reference(torch_cuda):
    """`tc` carries the reference's TimeCounter categories (XTDA_GPU.py:481-499) filled from the CUDA-event phase timers."""
    from xtddft_b200.SF_TDA import SF_TDA
    from xtddft_b200.synth import make_problem
    p = make_problem(290, 40, 2, 248, 30, 3000, xctype="LDA", hyb=0.5, seed=77)       # dim 10 500 < 50 000: totals only
    obj = SF_TDA(p, isf=-1)
    obj.kernel(nstates=3)
    assert obj.tc.Ap > 0 and obj.tc.dv > 0 and obj.tc.Adv == 0.0
    p = make_problem(420, 120, 2, 298, 16, 2000, xctype="LDA", hyb=0.5, seed=78)      # dim 36 600 ... still small; force detail
    obj = SF_TDA(p, isf=-1)
    eng = obj._get_engine()
    assert eng.ext_dim == 122 * 300
    import xtddft_b200.drivers_common as dc
    conv, e, v, cyc, hd = dc.solve(eng, 3, "sf_down", tc=obj.tc)
    assert obj.tc.dv > 0
    # the per-phase categories are collected from 50 000 unknowns on; check the accounting on this engine directly
    import torch
    z = torch.randn((3, eng.ext_dim), dtype=torch.float64, device="cuda")
    eng.sigma(z)
    ms = eng.stats()["ms"]
    assert ms["k1"] > 0 and ms["k2"] > 0 and ms["xc_gemm"] > 0 and ms["allreduce"] == 0.0


@pytest.mark.parametrize("method", ["sf_down", "xtda", "xsf"])
def test_native_davidson_matches_python_solver(torch_cuda, method):
    """`xtd_davidson` (solver loop in C++ inside libxtdsigma) against the Python solver over the same kernels: same roots to
    1e-10 Eh, same convergence flags, Ritz vectors equal up to sign, cycle / sigma-vector counts within one cycle."""
    from xtddft_b200 import davidson as dav
    from xtddft_b200 import plan as planmod
    from xtddft_b200.engine import SigmaEngine
    from xtddft_b200.synth import make_problem
    p = make_problem(60, 12, 2, 46, 30, 400, xctype="LDA", hyb=0.4, seed=91)
    if method == "sf_down":
        plan = planmod.build_sf_plan(p, isf=-1, method=0, hdiag_kind="sf")
    elif method == "xtda":
        plan = planmod.build_xtda_plan(p)
    else:
        plan = planmod.build_sf_plan(p, isf=-1, method=0, sa=3, layout=planmod.LAYOUT_BLOCK, remove=True, hdiag_kind="xsf")
    eng = SigmaEngine.from_problem(plan, p, workspace_bytes=256 << 20, max_nvec=16)
    c1, e1, x1, i1 = dav.davidson_for_engine(eng, 6, method, native=False)
    c2, e2, x2, i2 = dav.davidson_for_engine(eng, 6, method, native=True)
    assert np.all(c1) and np.all(c2)
    assert np.abs(e1 - e2).max() < 1e-10
    ov = np.abs(np.array(x1) @ np.array(x2).T)
    assert np.abs(np.diag(ov) - 1.0).max() < 1e-6, np.diag(ov)
    assert abs(i1[0] - i2[0]) <= 1 and abs(i1[1] - i2[1]) <= 6, (i1, i2)
    eng.close()
