"""Parity of the CUDA sigma path (through the C-ABI) with the oracle on seeded inputs.  Needs a B200."""
import numpy as np
import pytest

from oracle import sigma as osig
from xtddft_b200 import plan as planmod
from xtddft_b200.synth import make_problem

pytestmark = pytest.mark.gpu

RTOL = 1e-9     # north_star: sigma vectors within 1e-9 relative; measured differences are ~1e-13


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def _engine(plan, p, **kw):
    from xtddft_b200.engine import SigmaEngine
    return SigmaEngine.from_problem(plan, p, workspace_bytes=512 << 20, max_nvec=kw.pop("max_nvec", 8), **kw)


def _check(torch, eng, vind, dim, nvec=3, seed=0):
    z = np.random.default_rng(seed).standard_normal((nvec, dim))
    ref = vind(z)
    got = eng.sigma(torch.from_numpy(z).cuda()).cpu().numpy()
    err = np.abs(got - ref).max() / max(1.0, np.abs(ref).max())
    assert err < RTOL, err
    got_h = eng.sigma_host(z)
    assert np.abs(got_h - got).max() == 0.0
    return err


@pytest.mark.parametrize("xct,hyb", [("GGA", 0.2), ("LDA", 0.0), ("HF", 1.0), ("LDA", 0.3)])
@pytest.mark.parametrize("no", [1, 2])
@pytest.mark.parametrize("restricted", [True, False])
def test_xtda(torch_cuda, xct, hyb, no, restricted):
    p = make_problem(20 + no, 5, no, 15, 23, 300, xctype=xct, hyb=hyb, restricted=restricted, seed=100 + no)
    vind, hd = osig.xtda_gen_vind(p)
    pl = planmod.build_xtda_plan(p)
    eng = _engine(pl, p)
    _check(torch_cuda, eng, vind, hd.size)
    assert np.abs(eng.hdiag() - hd).max() < 1e-13
    eng.close()


@pytest.mark.parametrize("isf", [-1, 1])
@pytest.mark.parametrize("method", [0, 1, 2])
@pytest.mark.parametrize("restricted", [True, False])
def test_sf(torch_cuda, isf, method, restricted):
    p = make_problem(21, 4, 2, 15, 19, 260, xctype="GGA", hyb=0.5, restricted=restricted, seed=110)
    vind, hd = osig.sf_gen_vind(p, isf, method)
    pl = planmod.build_sf_plan(p, isf=isf, method=method)
    eng = _engine(pl, p)
    _check(torch_cuda, eng, vind, hd.size)
    eng.close()


@pytest.mark.parametrize("sa", [0, 1, 2, 3])
@pytest.mark.parametrize("remove", [False, True])
@pytest.mark.parametrize("no,nc,method", [(2, 5, 0), (3, 4, 1)])
def test_xsf_block(torch_cuda, sa, remove, no, nc, method):
    p = make_problem(nc + no + 13, nc, no, 13, 21, 280, xctype="GGA", hyb=0.4, seed=120 + no)
    vind, hd = osig.xsf_gen_vind(p, sa=sa, method=method, remove=remove, foo=0.8, fglobal=0.7)
    pl = planmod.build_sf_plan(p, isf=-1, method=method, sa=sa, layout=planmod.LAYOUT_BLOCK, remove=remove, foo=0.8, fglobal=0.7,
                               hdiag_kind="xsf")
    eng = _engine(pl, p)
    _check(torch_cuda, eng, vind, hd.size)
    assert np.abs(eng.hdiag() - hd).max() < 1e-11
    eng.close()


@pytest.mark.parametrize("x_level", [0, 3])
@pytest.mark.parametrize("remove", [False, True])
def test_xsf_gpu_order(torch_cuda, x_level, remove):
    p = make_problem(20, 5, 2, 13, 21, 280, xctype="GGA", hyb=0.4, seed=130)
    vind, hd = osig.xsf_gpu_gen_vind(p, x_level=x_level, collinear="mcol", extype=1, remove=remove, foo=0.8, fglobal=0.7)
    pl = planmod.build_sf_plan(p, isf=-1, method=1, sa=x_level, layout=planmod.LAYOUT_PYSCF, remove=remove, foo=0.8, fglobal=0.7,
                               hdiag_kind="gpu")
    eng = _engine(pl, p)
    _check(torch_cuda, eng, vind, hd.size)
    eng.close()


def test_rsh_and_many_vectors(torch_cuda):
    p = make_problem(24, 6, 2, 16, 25, 300, xctype="GGA", hyb=0.2, seed=140, omega=0.33, alpha=0.65)
    vind, hd = osig.xtda_gen_vind(p)
    eng = _engine(planmod.build_xtda_plan(p), p, max_nvec=5)
    _check(torch_cuda, eng, vind, hd.size, nvec=12)      # more vectors than max_nvec: batched internally
    eng.close()


def test_medium_size_tiles(torch_cuda):
    """Sizes that cross tile boundaries (M, N, K > 128) in every GEMM of the path."""
    p = make_problem(300, 40, 2, 258, 150, 3000, xctype="GGA", hyb=0.5, seed=150)
    vind, hd = osig.xsf_gen_vind(p, sa=3, method=0, remove=True)
    pl = planmod.build_sf_plan(p, isf=-1, method=0, sa=3, layout=planmod.LAYOUT_BLOCK, remove=True, hdiag_kind="xsf")
    eng = _engine(pl, p, max_nvec=4)
    _check(torch_cuda, eng, vind, hd.size, nvec=4)
    eng.close()
    p = make_problem(200, 30, 1, 169, 120, 2000, xctype="GGA", hyb=0.2, seed=151)
    vind, hd = osig.xtda_gen_vind(p)
    eng = _engine(planmod.build_xtda_plan(p), p, max_nvec=4)
    _check(torch_cuda, eng, vind, hd.size, nvec=3)
    eng.close()
