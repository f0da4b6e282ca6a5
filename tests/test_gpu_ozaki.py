"""FP64 contraction emulated on the INT8 tensor cores (tcgen05.mma kind::i8 + TMEM, csrc/ozaki.cuh) against an fp64
reference of the same contraction (numpy on the host: the oracle for a floating-point kernel is the plain fp64 product).
Error model: each operand row is cut into S signed 7-bit digits below a power-of-two row scale, products with i + j >= S are
dropped, so |C - C_ref| <~ S 2^(-7 S + 2) * sum_k |a||b| -- with S = 8 that is the fp64 rounding level.  Needs a B200."""
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


def _ozaki(torch, a, b, slices, group=0, alpha=1.0, c0=None):
    """a [nq, M, K], b [nq, N, K] (numpy) -> C [M, N] through xtd_ozaki_gemm with padded leading dimensions."""
    from xtddft_b200 import _lib
    lib = _lib.load()
    nq, m, k = a.shape
    n = b.shape[1]
    ld = (k + 15) // 16 * 16
    ad = torch.full((nq, m, ld), 7.0, dtype=torch.float64, device="cuda")      # garbage in the padding: must be ignored
    bd = torch.full((nq, n, ld), -3.0, dtype=torch.float64, device="cuda")
    ad[:, :, :k] = torch.from_numpy(a).cuda()
    bd[:, :, :k] = torch.from_numpy(b).cuda()
    ldc = (n + 1) // 2 * 2
    cd = torch.zeros((m, ldc), dtype=torch.float64, device="cuda")
    if c0 is not None:
        cd[:, :n] = torch.from_numpy(c0).cuda()
    ms = (C.c_double * 3)()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.xtd_ozaki_gemm(st, m, n, k, nq, slices, group, C.c_void_p(ad.data_ptr()), ld, m * ld, C.c_void_p(bd.data_ptr()), ld, n * ld,
                                  C.c_void_p(cd.data_ptr()), ldc, alpha, 1 if c0 is not None else 0, ms), "xtd_ozaki_gemm")
    return cd[:, :n].cpu().numpy(), list(ms)


def _ref(a, b):
    return np.einsum("qmk,qnk->mn", a, b, optimize=True)


@pytest.mark.parametrize("m,n,k,nq", [(128, 64, 32, 1), (100, 50, 70, 3), (300, 200, 129, 5), (17, 9, 5, 2), (257, 130, 64, 20)])
def test_ozaki_small_exact_level(torch_cuda, m, n, k, nq):
    """8 slices: 55 bits below the row scale -> agreement at the fp64 rounding level; ragged M, N, K and several q-slices."""
    rng = np.random.default_rng(m + n + k)
    a = rng.standard_normal((nq, m, k))
    b = rng.standard_normal((nq, n, k))
    c, _ = _ozaki(torch_cuda, a, b, 8)
    ref = _ref(a, b)
    bound = np.einsum("qmk,qnk->mn", np.abs(a), np.abs(b)).max()
    assert np.abs(c - ref).max() <= 1e-13 * bound


def test_ozaki_integer_inputs_are_exact(torch_cuda):
    """Small integers are represented exactly by the first two digits: the int8 pipeline must reproduce the product bit
    for bit (catches any descriptor / layout / level-weight mistake without tolerance)."""
    rng = np.random.default_rng(5)
    a = rng.integers(-50, 51, (4, 256, 96)).astype(np.float64)
    b = rng.integers(-50, 51, (4, 128, 96)).astype(np.float64)
    for s in (3, 5, 8):
        c, _ = _ozaki(torch_cuda, a, b, s)
        assert np.array_equal(c, _ref(a, b)), s


@pytest.mark.parametrize("slices", [4, 5, 6, 7, 8])
def test_ozaki_error_vs_slices(torch_cuda, slices):
    """Truncation error drops by 2^-7 per slice; rows with a wide dynamic range (1e-6 .. 1) inside one scale group."""
    rng = np.random.default_rng(9)
    nq, m, n, k = 6, 200, 150, 300
    a = rng.standard_normal((nq, m, k)) * 10.0 ** rng.uniform(-6, 0, (nq, m, k))
    b = rng.standard_normal((nq, n, k)) * 10.0 ** rng.uniform(-6, 0, (nq, n, k))
    c, _ = _ozaki(torch_cuda, a, b, slices, group=2)
    ref = _ref(a, b)
    # row / column maxima over a group bound every element; products below 2^(-7 S) of (row max * col max) are dropped
    amax = np.abs(a).reshape(3, 2, m, k).max(axis=(1, 3))          # [groups, m]
    bmax = np.abs(b).reshape(3, 2, n, k).max(axis=(1, 3))
    bound = (slices + 1) * 2.0 ** (-7 * slices + 2) * 2 * k * np.einsum("gm,gn->mn", amax, bmax)
    assert (np.abs(c - ref) <= bound + 1e-15 * np.abs(ref).max()).all(), float((np.abs(c - ref) / bound).max())


def test_ozaki_alpha_accumulate_and_zero_rows(torch_cuda):
    rng = np.random.default_rng(11)
    a = rng.standard_normal((3, 140, 40))
    b = rng.standard_normal((3, 70, 40))
    a[:, 5] = 0.0                   # an all-zero row: scale 0
    a[1] = 0.0                      # an all-zero q-slice
    c0 = rng.standard_normal((140, 70))
    c, _ = _ozaki(torch_cuda, a, b, 8, group=1, alpha=-0.5, c0=c0)
    ref = c0 - 0.5 * _ref(a, b)
    assert np.abs(c - ref).max() <= 1e-13 * np.abs(ref).max()
    assert np.array_equal(c[5], c0[5])


def test_ozaki_many_groups_and_splits(torch_cuda):
    """More groups than stages and a split group range (tiles x splits > 148 SMs is not needed for splitting to kick in)."""
    rng = np.random.default_rng(13)
    a = rng.standard_normal((64, 128, 64))
    b = rng.standard_normal((64, 64, 64))
    c, _ = _ozaki(torch_cuda, a, b, 7, group=4)
    ref = _ref(a, b)
    assert np.abs(c - ref).max() <= 1e-11 * np.abs(ref).max()
