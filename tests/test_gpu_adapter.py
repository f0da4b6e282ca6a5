"""Level-1 drop-in, end to end: the driver classes are constructed from a duck-typed mean-field OBJECT (the stand-in objects that
tests/golden/make_golden.py ran the REFERENCE's own classes on), go through `adapters.from_pyscf` (packed 3-centre tensor
streamed to the device, kernels built with the reference's numint calls) and must reproduce the golden sigma vectors and
preconditioner diagonals the reference produced from the same objects.  Needs a B200."""
import os

import numpy as np
import pytest

from adapter_fakes import fake_scf, packed_df
from xtddft_b200.synth import make_problem

pytestmark = pytest.mark.gpu
RTOL = 1e-9


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def _problem(d):
    prm = d["params"]
    nc, no, nv, naux, ng, seed = [int(v) for v in prm[:6]]
    restricted = bool(prm[6]) if len(prm) > 6 else True
    return make_problem(nc + no + nv, nc, no, nv, naux, ng, xctype=str(d["xctype"]), hyb=float(d["hyb"]), restricted=restricted, seed=seed)


def _close(a, b, tol=RTOL):
    b = np.atleast_2d(b)
    assert a.shape == b.shape
    assert float(np.abs(a - b).max()) <= tol * max(1.0, float(np.abs(b).max()))


def _mf(fakes, p, on_disk=False):
    FakeROKS, FakeUKS = fakes
    mf = (FakeROKS if p.restricted else FakeUKS)(p)
    mf.with_df = packed_df(p, on_disk=on_disk)
    return mf


@pytest.mark.parametrize("tag", ["roks_gga_no1", "roks_lda_no3", "uks_gga_no1", "roks_mgga_no2"])
def test_xtda_from_object(torch_cuda, golden_dir, tag):
    from xtddft_b200.XTDA import XTDA
    d = np.load(os.path.join(golden_dir, f"xtda_{tag}.npz"), allow_pickle=False)
    p = _problem(d)
    with fake_scf() as fakes:
        mf = _mf(fakes, p)
        obj = XTDA(mf.mol, mf, nstates=3)
        vind, hdiag = obj.gen_vind(mf)
        _close(vind(d["z"]), d["hx"])
        assert np.abs(hdiag - d["hdiag"]).max() < 1e-12
        assert np.array_equal(obj.get_init_guess(mf, 3), d["x0"])


@pytest.mark.parametrize("tag", ["down_gga", "up_gga", "down_uks", "down_mgga"])
def test_sf_from_object(torch_cuda, golden_dir, tag):
    """The adapter builds the ALDA0 kernel itself (SF_TDA.py:39-88) -- nothing is taken from the fixture but z / hx."""
    from xtddft_b200.SF_TDA import SF_TDA
    d = np.load(os.path.join(golden_dir, f"sf_{tag}.npz"), allow_pickle=False)
    p = _problem(d)
    with fake_scf() as fakes:
        mf = _mf(fakes, p, on_disk=(tag == "down_gga"))
        obj = SF_TDA(mf, isf=int(d["params"][7]), davidson=True, method=0)
        vind, hdiag = obj.gen_tda_operation_sf()
        _close(vind(d["z"]), d["hx"])
        assert np.abs(hdiag - d["hdiag"]).max() < 1e-12


@pytest.mark.parametrize("tag", ["gga_no2", "lda_no3"])
def test_xsf_from_object(torch_cuda, golden_dir, tag):
    """`XSF_TDA(FakeROKS(p)).gen_tda_operation_sf(0.8, 0.7)` for every SA level, with and without the removed OO vector."""
    from xtddft_b200.XSF_TDA import XSF_TDA
    d = np.load(os.path.join(golden_dir, f"xsf_{tag}.npz"), allow_pickle=False)
    p = _problem(d)
    with fake_scf() as fakes:
        mf = _mf(fakes, p)
        for sa in (0, 1, 2, 3):
            for re in (0, 1):
                obj = XSF_TDA(mf, SA=sa, davidson=True, method=0)
                obj.re = bool(re)
                vind, hdiag = obj.gen_tda_operation_sf(0.8, 0.7)
                _close(vind(d[f"z_sa{sa}_re{re}"]), d[f"hx_sa{sa}_re{re}"])
                assert np.abs(hdiag - d[f"hdiag_sa{sa}_re{re}"]).max() < 1e-10


def test_xsf_gpu_class_from_object_with_cached_mcol_kernel(torch_cuda, golden_dir):
    """The class default collinear='mcol' with a caller-cached multicollinear kernel (the sampling itself is mcfun's)."""
    from xtddft_b200.XSF_TDA_GPU import XSF_TDA_GPU
    d = np.load(os.path.join(golden_dir, "xsfgpu_gga_no2.npz"), allow_pickle=False)
    p = _problem(d)
    with fake_scf() as fakes:
        mf = _mf(fakes, p)
        mf.fxc_sf_mc = p.fxc_mcol
        for X in (0, 3):
            obj = XSF_TDA_GPU(mf, X=X, collinear="mcol", nstates=3, extype=1, remove=True, foo=0.8, fglobal=0.7)
            vind, hdiag = obj.gen_vind()
            _close(vind(d[f"z_X{X}_re1"]), d[f"hx_X{X}_re1"])
            assert np.abs(hdiag - d[f"hdiag_X{X}_re1"]).max() < 1e-10


def test_be_energies_with_pyscf(torch_cuda):
    """XSF_TDA.py:1558-1574: Be / aug-cc-pVTZ, ROKS BHandHLYP triplet reference, 10 XSF-TDA roots (eV) printed in the reference
    source.  Needs a real PySCF (not installable in the build or bench containers): skipped without it."""
    pyscf = pytest.importorskip("pyscf")
    if not hasattr(pyscf, "__version__"):
        pytest.skip("a stub pyscf is installed")
    from pyscf import dft, gto
    from xtddft_b200.XSF_TDA import XSF_TDA
    mol = gto.M(atom="Be 0 0 0", basis="aug-cc-pvtz", spin=2, symmetry="D2h", verbose=0)
    mf = dft.ROKS(mol).density_fit()          # the reference example runs exact integrals; this build is DF-only
    mf.xc = "bhandhlyp"
    mf.kernel()
    e, _ = XSF_TDA(mf).kernel(nstates=10, remove=True)
    ref = np.array([-2.58159612, 1.94501967, 2.0441558, 2.04415705, 3.55556409, 4.0395836, 4.07260624, 4.07260634, 4.09542032,
                    4.09542242])               # XSF_TDA.py:1574 (eV)
    assert np.abs(np.sort(e) - np.sort(ref)).max() < 2e-3       # density-fitting error of the SCF and the response (~1e-4 eV)
