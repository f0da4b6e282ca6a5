"""Control flow of the product Davidson (device-resident subspace) against the oracle restatement of the
reference solver, using a NumPy vector backend.  CPU only."""
import numpy as np
import pytest

from numpy_vectors import NumpyVectors
from oracle import davidson as odav
from oracle import sigma as osig
from xtddft_b200 import davidson as pdav
from xtddft_b200.synth import make_problem


def _matrix(n, seed, gap=0.01):
    rng = np.random.default_rng(seed)
    a = np.diag(np.sort(rng.uniform(0.05, 4.0, n))) + gap * rng.standard_normal((n, n))
    return 0.5 * (a + a.T)


@pytest.mark.parametrize("nroots", [1, 5, 10])
@pytest.mark.parametrize("max_space", [12, 6])
def test_same_roots_and_cycles_as_reference_solver(nroots, max_space):
    a = _matrix(400, nroots)
    hd = a.diagonal().copy()
    x0 = pdav.init_guess(hd, nroots, 1e-5)
    conv_o, e_o, x_o, info_o = odav.davidson1(lambda xs: xs @ a, x0, hd.copy(), tol=1e-9, nroots=nroots, max_cycle=200,
                                              max_space=max_space, lindep=1e-14)
    conv_p, e_p, x_p, info_p = pdav.davidson1(lambda xs: xs @ a, x0, hd.copy(), tol=1e-9, nroots=nroots, max_cycle=200,
                                              max_space=max_space, lindep=1e-14, backend=NumpyVectors(400))
    assert conv_o.all() and conv_p.all()
    assert np.abs(e_o - e_p).max() < 1e-10
    assert np.abs(e_p - np.linalg.eigvalsh(a)[:nroots]).max() < 1e-8
    assert abs(info_o[0] - info_p[0]) <= 1 and abs(info_o[1] - info_p[1]) <= nroots
    for xo, xp in zip(x_o, x_p):
        assert min(np.abs(xo - xp).max(), np.abs(xo + xp).max()) < 1e-6


def test_pick_and_tol_residual():
    a = _matrix(300, 7)
    hd = a.diagonal().copy()
    x0 = pdav.init_guess(hd, 6, 1e-3)
    conv, e, x, info = pdav.davidson1(lambda xs: xs @ a, x0, hd, tol=1e-12, tol_residual=1e-5, lindep=1e-12, nroots=6, max_cycle=100,
                                      level_shift=0.0, pick=pdav.pick_positive, backend=NumpyVectors(300))
    assert conv.all()
    assert np.abs(e - np.linalg.eigvalsh(a)[:6]).max() < 1e-9


def test_davidson_on_oracle_operator_matches_dense():
    """Solve the XSF-TDA problem with the product solver over the ORACLE operator and compare with dense eigh."""
    p = make_problem(14, 4, 2, 8, 12, 40, xctype="GGA", hyb=0.4, seed=5)
    vind, hd = osig.xsf_gen_vind(p, sa=3, method=0, remove=True)
    a = np.asarray(vind(np.eye(hd.size)))
    x0 = pdav.init_guess(hd, 4, 1e-5)
    conv, e, x, info = pdav.davidson1(lambda xs: np.asarray(vind(xs)), x0, hd, tol=1e-8, lindep=1e-9, nroots=4, max_cycle=200,
                                      backend=NumpyVectors(hd.size))
    assert conv.all()
    assert np.abs(e - np.linalg.eigvalsh(a)[:4]).max() < 1e-7
