"""Pin the oracle against fixtures produced by executing the reference's own code
(tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest

from oracle import layouts, numint, sigma
from xtddft_b200.synth import make_problem

TOL = 1e-11


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def _problem(d, **kw):
    prm = d["params"]
    nc, no, nv, naux, ng, seed = [int(v) for v in prm[:6]]
    restricted = bool(prm[6]) if len(prm) > 6 else True
    return make_problem(nc + no + nv, nc, no, nv, naux, ng, xctype=str(d["xctype"]), hyb=float(d["hyb"]),
                        restricted=restricted, seed=seed, **kw)


def _close(a, b, tol=TOL):
    scale = max(1.0, float(np.abs(b).max()))
    assert a.shape == b.shape
    assert float(np.abs(a - b).max()) <= tol * scale, float(np.abs(a - b).max())


def test_helpers(golden_dir):
    h = _load(golden_dir, "helpers.npz")
    for (nc, no, nv) in [(2, 1, 3), (3, 2, 4), (1, 3, 2), (4, 1, 1)]:
        assert np.array_equal(layouts.order_pyscf2my(nc, no, nv), h[f"order_{nc}_{no}_{nv}"])
        dim = (nc + no) * nv + nc * (no + nv)
        v = np.random.default_rng(7).standard_normal((dim, 3))
        _close(layouts.so2st(v, nc, no, nv), h[f"so2st_{nc}_{no}_{nv}"], 1e-15)
        _close(layouts.st2so(v, nc, no, nv), h[f"st2so_{nc}_{no}_{nv}"], 1e-15)
    for no in (2, 3, 4):
        _close(layouts.get_vect(no), h[f"vects_{no}"], 1e-15)


@pytest.mark.parametrize("tag", ["roks_gga_no1", "roks_gga_no2", "roks_lda_no3", "roks_hf_no1", "uks_gga_no1", "roks_mgga_no2"])
def test_xtda_sigma(golden_dir, tag):
    d = _load(golden_dir, f"xtda_{tag}.npz")
    p = _problem(d)
    vind, hdiag = sigma.xtda_gen_vind(p)
    _close(hdiag, d["hdiag"], 1e-14)
    _close(vind(d["z"]), d["hx"])


@pytest.mark.parametrize("tag", ["down_gga", "up_gga", "down_lda", "down_uks", "down_mgga"])
def test_sf_sigma(golden_dir, tag):
    d = _load(golden_dir, f"sf_{tag}.npz")
    p = _problem(d)
    p.fxc_alda0 = d["fxc_alda0"]          # kernel the reference built (input of the path)
    isf = int(d["params"][7])
    vind, hdiag = sigma.sf_gen_vind(p, isf, 0)
    _close(hdiag, d["hdiag"], 1e-14)
    _close(vind(d["z"]), d["hx"])
    if isf == -1:
        out = layouts.deal_v_davidson(d["deal_in"], p.nc, p.no, p.nv)
        _close(out, d["deal_out"], 1e-15)


def test_sf_mcol_contraction(golden_dir):
    d = _load(golden_dir, "sf_mcol_contraction.npz")
    p = make_problem(9, 3, 2, 4, 10, 36, xctype="GGA", hyb=0.5, seed=25)
    _close(numint.nr_uks_fxc_sf_mc(p.ao, p.weights, p.fxc_mcol, d["dms"]), d["v"])


def test_sf_mcol_contraction_mgga(golden_dir):
    """tau component of the multicollinear kernel (SF_TDA.py:1028-1040; the shipped file needs MGGA_DENSITY_LAPL defined)"""
    d = _load(golden_dir, "sf_mcol_contraction_mgga.npz")
    p = make_problem(9, 3, 2, 4, 10, 36, xctype="MGGA", hyb=0.5, seed=27)
    assert p.fxc_mcol.shape[0] == 5 and p.ao.shape[0] == 4
    _close(numint.nr_uks_fxc_sf_mc(p.ao, p.weights, p.fxc_mcol, d["dms"]), d["v"])


@pytest.mark.parametrize("tag", ["gga_no2", "lda_no3"])
def test_xsf_block_sigma(golden_dir, tag):
    d = _load(golden_dir, f"xsf_{tag}.npz")
    p = _problem(d)
    p.fxc_alda0 = d["fxc_alda0"]
    for sa in (0, 1, 2, 3):
        for re in (0, 1):
            vind, hdiag = sigma.xsf_gen_vind(p, sa=sa, method=0, remove=bool(re), foo=0.8, fglobal=0.7)
            _close(hdiag, d[f"hdiag_sa{sa}_re{re}"], 1e-12)
            _close(vind(d[f"z_sa{sa}_re{re}"]), d[f"hx_sa{sa}_re{re}"])


@pytest.mark.parametrize("tag", ["gga_no2", "gga_no3"])
def test_xsf_gpu_order_sigma(golden_dir, tag):
    d = _load(golden_dir, f"xsfgpu_{tag}.npz")
    p = _problem(d)
    for X in (0, 1, 2, 3):
        for re in (0, 1):
            vind, hdiag = sigma.xsf_gpu_gen_vind(p, x_level=X, collinear="mcol", extype=1, remove=bool(re), foo=0.8, fglobal=0.7)
            _close(hdiag, d[f"hdiag_X{X}_re{re}"], 1e-12)
            _close(vind(d[f"z_X{X}_re{re}"]), d[f"hx_X{X}_re{re}"])
    vind, hdiag = sigma.xsf_gpu_gen_vind(p, x_level=0, collinear="mcol", extype=0, remove=False)
    _close(hdiag, d["hdiag_up"], 1e-14)
    _close(vind(d["z_up"]), d["hx_up"])
