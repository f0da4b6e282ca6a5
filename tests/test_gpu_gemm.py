"""The DMMA + TMA GEMM through the C-ABI test entry `xtd_dgemm` against NumPy: all four operand layouts, every tile
configuration (main 128x128 tile; short N tail <= 64 / <= 32; short M tail <= 64 / <= 16; corner with both), split-K with
the deterministic reduction, alpha and accumulate.  fp64: tolerance 1e-12 relative (summation order differs)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def _pad(n):
    return (n + 15) // 16 * 16


def _dev(torch, a):
    """host [r, c] -> device buffer with padded leading dimension; returns (tensor, ld)"""
    buf = torch.zeros((a.shape[0], _pad(a.shape[1])), dtype=torch.float64, device="cuda")
    buf[:, :a.shape[1]] = torch.from_numpy(a).cuda()
    return buf, buf.stride(0)


SHAPES = [
    (128, 128, 64),     # one full tile
    (300, 17 + 256, 40),    # N tail 17 (<= 32), M tail 44 (<= 64)
    (256 + 9, 128 + 53, 100),   # M tail 9 (<= 16), N tail 53 (<= 64)
    (137, 821, 1000),   # config-4 shapes: M tail 9, N tail 53, split-K
    (2 * 128 + 64, 3 * 128 + 32, 33),   # tails exactly 64 and 32
    (128 + 65, 128 + 100, 50),  # tails > 64: main configuration only
    (100, 50, 70),      # single partial tile
    (1, 1, 1),
]


@pytest.mark.parametrize("a_kc,b_kc", [(1, 1), (1, 0), (0, 1), (0, 0)])
@pytest.mark.parametrize("m,n,k", SHAPES)
def test_dgemm_layouts_and_tails(torch_cuda, monkeypatch, m, n, k, a_kc, b_kc):
    torch = torch_cuda
    from xtddft_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(m * 7 + n * 3 + k)
    a, b = rng.standard_normal((m, k)), rng.standard_normal((n, k))
    c0 = rng.standard_normal((m, n))
    ad, lda = _dev(torch, a if a_kc else np.ascontiguousarray(a.T))
    bd, ldb = _dev(torch, b if b_kc else np.ascontiguousarray(b.T))
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for acc, alpha in [(0, 1.0), (1, -0.75)]:
        cd, ldc = _dev(torch, c0)
        _lib.check(lib.xtd_dgemm(s, m, n, k, alpha, C.c_void_p(ad.data_ptr()), lda, a_kc, C.c_void_p(bd.data_ptr()), ldb, b_kc,
                                 C.c_void_p(cd.data_ptr()), ldc, acc), "xtd_dgemm")
        ref = alpha * (a @ b.T) + (c0 if acc else 0.0)
        got = cd[:, :n].cpu().numpy()
        assert np.abs(got - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max()) * max(1, k) ** 0.5
        # nothing written past the view
        assert float(cd[:, n:].abs().max()) == 0.0 if cd.shape[1] > n else True
