"""Explicit-A == sigma-build identities (the reference's own dense-vs-iterative validation)
and Davidson known-answer tests on the oracle.  CPU only."""
import numpy as np
import pytest

from oracle import amat, davidson, layouts, sigma
from xtddft_b200.synth import make_problem


def _apply(vind, dim):
    return np.asarray(vind(np.eye(dim)))


@pytest.mark.parametrize("xct", ["GGA", "LDA", "HF"])
@pytest.mark.parametrize("no", [1, 2, 3])
def test_xtda_sigma_equals_explicit_a(xct, no):
    p = make_problem(9 + no, 3, no, 6, 11, 40, xctype=xct, hyb=0.3, seed=3 + no)
    vind, hd = sigma.xtda_gen_vind(p)
    a = _apply(vind, hd.size)
    assert np.abs(a - amat.xtda_amat(p)).max() < 1e-12
    assert np.abs(a - a.T).max() < 1e-12


@pytest.mark.parametrize("isf", [-1, 1])
@pytest.mark.parametrize("method", [0, 1, 2])
def test_sf_sigma_equals_explicit_a(isf, method):
    p = make_problem(11, 3, 2, 6, 11, 40, xctype="GGA", hyb=0.5, seed=5)
    vind, hd = sigma.sf_gen_vind(p, isf, method)
    a = _apply(vind, hd.size)
    assert np.abs(a - amat.sf_amat_pyscf(p, isf, method)).max() < 1e-12


@pytest.mark.parametrize("sa", [0, 1, 2, 3])
@pytest.mark.parametrize("remove", [False, True])
@pytest.mark.parametrize("no", [2, 3])
def test_xsf_sigma_equals_explicit_a(sa, remove, no):
    p = make_problem(8 + no, 3, no, 5, 10, 36, xctype="GGA", hyb=0.4, seed=7)
    vind, hd = sigma.xsf_gen_vind(p, sa=sa, method=0, remove=remove, foo=0.8, fglobal=0.7)
    a = _apply(vind, hd.size)
    assert np.abs(a - amat.xsf_amat(p, sa=sa, method=0, foo=0.8, fglobal=0.7, remove=remove)).max() < 1e-12
    # the PySCF-order (GPU class) builder is the same operator in another basis
    vg, hg = sigma.xsf_gpu_gen_vind(p, x_level=sa, collinear="alda0", remove=remove, foo=0.8, fglobal=0.7)
    ag = _apply(vg, hg.size)
    assert np.abs(np.linalg.eigvalsh(a) - np.linalg.eigvalsh(ag)).max() < 1e-11


def test_removed_layout_roundtrip():
    nc, no, nv = 3, 3, 4
    vects = layouts.get_vect(no)
    assert np.abs(vects.T @ vects - np.eye(no * no - 1)).max() < 1e-14
    trace_vec = np.eye(no).ravel() / np.sqrt(no)
    assert np.abs(trace_vec @ vects).max() < 1e-14
    z = np.random.default_rng(0).standard_normal((2, (nc + no) * (no + nv) - 1))
    full = layouts.gpu_order_expand(z, nc, no, nv, vects)
    assert np.abs(layouts.gpu_order_compress(full, nc, no, nv, vects) - z).max() < 1e-14


@pytest.mark.parametrize("nroots", [1, 5, 10])
def test_davidson_kat(nroots):
    rng = np.random.default_rng(nroots)
    n = 300
    a = np.diag(np.sort(rng.uniform(0.1, 5.0, n))) + 0.01 * rng.standard_normal((n, n))
    a = 0.5 * (a + a.T)
    a[3, 3] = a[4, 4]                       # a near-degenerate pair
    hd = a.diagonal().copy()
    idx = np.argsort(hd)[:nroots + 2]
    x0 = np.zeros((idx.size, n)); x0[np.arange(idx.size), idx] = 1
    conv, e, x, info = davidson.davidson1(lambda xs: xs @ a, x0, hd, tol=1e-10, nroots=nroots, max_cycle=100, lindep=1e-14)
    assert conv.all()
    assert np.abs(e - np.linalg.eigvalsh(a)[:nroots]).max() < 1e-8
    assert info[0] >= 1 and info[1] >= nroots


def test_davidson_restart_and_pick():
    rng = np.random.default_rng(1)
    n = 200
    a = np.diag(np.linspace(0.05, 4.0, n)) + 0.02 * rng.standard_normal((n, n))
    a = 0.5 * (a + a.T)
    hd = a.diagonal().copy()
    ref = np.linalg.eigvalsh(a)
    ref_pos = ref[ref > 1e-3][:6]

    def pick(w, v, nroots, envs):
        idx = np.where(w > 1e-3)[0]
        return w[idx], v[:, idx], idx
    pos = np.where(hd > 1e-3)[0]
    x0 = np.zeros((8, n)); x0[np.arange(8), pos[:8]] = 1
    conv, e, x, info = davidson.davidson1(lambda xs: xs @ a, x0, hd, tol=1e-9, nroots=6, max_cycle=200, max_space=12,
                                          lindep=1e-12, pick=pick)
    assert conv.all()
    assert np.abs(e - ref_pos).max() < 1e-7
