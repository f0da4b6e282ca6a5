"""Host-only pieces of the native Davidson solver (`xtd_davidson`, csrc/davidson.cuh) called through the C-ABI without a device:
the Householder + QL eigensolver of the projected matrix against numpy.linalg.eigh, and the Gram-Schmidt coefficient routine
against the Python one the device solver uses.  CPU only."""
import ctypes as C

import numpy as np
import pytest

from xtddft_b200 import _lib
from xtddft_b200.davidson import _gs_coefficients


@pytest.mark.parametrize("n", [1, 2, 5, 31, 98])
def test_sym_eig_against_numpy(n):
    lib = _lib.load()
    r = np.random.default_rng(n)
    a = r.standard_normal((n, n))
    a = a + a.T + np.diag(np.linspace(0, 3 * n, n))
    if n == 31:                                  # clustered / degenerate eigenvalues
        q, _ = np.linalg.qr(r.standard_normal((n, n)))
        a = q @ np.diag(np.repeat([0.1, 0.1000001, 2.0, 2.0, 7.5], [7, 6, 6, 6, 6])) @ q.T
        a = 0.5 * (a + a.T)
    v = np.ascontiguousarray(a.copy())
    w = np.zeros(n)
    _lib.check(lib.xtd_host_sym_eig(C.c_void_p(v.ctypes.data), n, C.c_void_p(w.ctypes.data)), "xtd_host_sym_eig")
    wr = np.linalg.eigvalsh(a)
    scale = max(1.0, np.abs(wr).max())
    assert np.abs(w - wr).max() < 1e-13 * scale
    assert np.abs(v.T @ v - np.eye(n)).max() < 1e-13
    assert np.abs(a @ v - v * w).max() < 1e-12 * scale


def test_gs_coefficients_against_python():
    lib = _lib.load()
    r = np.random.default_rng(3)
    wv = r.standard_normal((7, 40))
    wv[3] = 0.5 * wv[0] - 2.0 * wv[2]            # a dependent vector: dropped
    g = np.ascontiguousarray(wv @ wv.T)
    t = np.zeros((7, 7))
    nk = lib.xtd_host_gs_coefficients(C.c_void_p(g.ctypes.data), 7, 1e-12, C.c_void_p(t.ctypes.data))
    ref = _gs_coefficients(g, 1e-12)
    assert nk == ref.shape[0] == 6
    assert np.abs(t[:nk] - ref).max() < 1e-12
    q = t[:nk] @ wv
    assert np.abs(q @ q.T - np.eye(nk)).max() < 1e-10
