"""NumPy stand-in for the device vector backend (TEST INFRASTRUCTURE): lets the CPU suite exercise the
Davidson control flow of xtddft_b200/davidson.py without a GPU.  Never imported by the product."""
import numpy as np


class NumpyVectors:
    def __init__(self, dim):
        self.dim = dim

    def alloc(self, rows):
        return np.zeros((rows, self.dim))

    def from_host(self, a):
        return np.array(a, dtype=float, copy=True)

    def to_host(self, m):
        return np.array(m, copy=True)

    def copy(self, dst, src):
        dst[...] = src

    def dots(self, a, b):
        return a @ b.T

    def lincomb(self, y, x, c, beta=0.0):
        y[...] = (beta * y if beta != 0.0 else 0.0) + c @ x

    def residual(self, r, ax, x, e):
        r[...] = ax - e[:, None] * x
        return np.einsum("ij,ij->i", r, r)

    def precond(self, x, hdiag, shift):
        d = hdiag.reshape(1, -1) - shift[:, None]
        d[abs(d) < 1e-8] = 1e-8
        x[...] = x / d
        return np.einsum("ij,ij->i", x, x)

    def scale(self, x, s):
        x *= np.asarray(s)[:, None]

    def transform(self, work, n_in, t):
        n_out = t.shape[0]
        if n_out:
            work[:n_out] = t @ work[:n_in]
