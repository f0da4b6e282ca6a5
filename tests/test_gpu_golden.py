"""CUDA sigma path (through the C-ABI) against the committed golden fixtures -- sigma vectors produced by executing
the REFERENCE's own `vind` closures (tests/golden/make_golden.py) -- so the GPU path is pinned to the reference
itself, not only to the oracle restatement.  Needs a B200."""
import os

import numpy as np
import pytest

from xtddft_b200 import plan as planmod
from xtddft_b200.synth import make_problem

pytestmark = pytest.mark.gpu

RTOL = 1e-9     # north_star tolerance on sigma vectors (relative); observed ~1e-13


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def _problem(d):
    prm = d["params"]
    nc, no, nv, naux, ng, seed = [int(v) for v in prm[:6]]
    restricted = bool(prm[6]) if len(prm) > 6 else True
    return make_problem(nc + no + nv, nc, no, nv, naux, ng, xctype=str(d["xctype"]), hyb=float(d["hyb"]), restricted=restricted, seed=seed)


def _run(plan, p, z):
    from xtddft_b200.engine import SigmaEngine
    eng = SigmaEngine.from_problem(plan, p, workspace_bytes=256 << 20, max_nvec=8)
    try:
        return eng.sigma_host(np.atleast_2d(z)), eng.hdiag()
    finally:
        eng.close()


def _close(a, b, tol=RTOL):
    b = np.atleast_2d(b)
    assert a.shape == b.shape
    err = float(np.abs(a - b).max()) / max(1.0, float(np.abs(b).max()))
    assert err <= tol, err


@pytest.mark.parametrize("tag", ["roks_gga_no1", "roks_gga_no2", "roks_lda_no3", "roks_hf_no1", "uks_gga_no1", "roks_mgga_no2"])
def test_xtda_golden(torch_cuda, golden_dir, tag):
    d = _load(golden_dir, f"xtda_{tag}.npz")
    p = _problem(d)
    hx, hdiag = _run(planmod.build_xtda_plan(p), p, d["z"])
    _close(hx, d["hx"])
    assert np.abs(hdiag - d["hdiag"]).max() < 1e-12


@pytest.mark.parametrize("tag", ["down_gga", "up_gga", "down_lda", "down_uks", "down_mgga"])
def test_sf_golden(torch_cuda, golden_dir, tag):
    d = _load(golden_dir, f"sf_{tag}.npz")
    p = _problem(d)
    p.fxc_alda0 = d["fxc_alda0"]
    isf = int(d["params"][7])
    hx, hdiag = _run(planmod.build_sf_plan(p, isf=isf, method=0), p, d["z"])
    _close(hx, d["hx"])
    assert np.abs(hdiag - d["hdiag"]).max() < 1e-12


@pytest.mark.parametrize("tag", ["gga_no2", "lda_no3"])
def test_xsf_block_golden(torch_cuda, golden_dir, tag):
    d = _load(golden_dir, f"xsf_{tag}.npz")
    p = _problem(d)
    p.fxc_alda0 = d["fxc_alda0"]
    for sa in (0, 1, 2, 3):
        for re in (0, 1):
            plan = planmod.build_sf_plan(p, isf=-1, method=0, sa=sa, layout=planmod.LAYOUT_BLOCK, remove=bool(re), foo=0.8, fglobal=0.7,
                                         hdiag_kind="xsf")
            hx, hdiag = _run(plan, p, d[f"z_sa{sa}_re{re}"])
            _close(hx, d[f"hx_sa{sa}_re{re}"])
            assert np.abs(hdiag - d[f"hdiag_sa{sa}_re{re}"]).max() < 1e-10


@pytest.mark.parametrize("tag", ["gga_no2", "gga_no3"])
def test_xsf_gpu_order_golden(torch_cuda, golden_dir, tag):
    d = _load(golden_dir, f"xsfgpu_{tag}.npz")
    p = _problem(d)
    for X in (0, 1, 2, 3):
        for re in (0, 1):
            plan = planmod.build_sf_plan(p, isf=-1, method=1, sa=X, layout=planmod.LAYOUT_PYSCF, remove=bool(re), foo=0.8, fglobal=0.7,
                                         hdiag_kind="gpu")
            hx, hdiag = _run(plan, p, d[f"z_X{X}_re{re}"])
            _close(hx, d[f"hx_X{X}_re{re}"])
            assert np.abs(hdiag - d[f"hdiag_X{X}_re{re}"]).max() < 1e-10
    # spin-flip-up branch of the GPU class (extype=0, XSF_TDA_GPU.py:422-439)
    plan = planmod.build_sf_plan(p, isf=1, method=1, hdiag_kind="gpu")
    hx, hdiag = _run(plan, p, d["z_up"])
    _close(hx, d["hx_up"])
    assert np.abs(hdiag - d["hdiag_up"]).max() < 1e-12
