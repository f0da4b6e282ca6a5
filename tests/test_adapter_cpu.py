"""`adapters.from_pyscf` executed on the stand-in mean-field objects of make_golden.py (the objects the reference's own classes
were run on to produce the fixtures): what it extracts must be what the reference's closures captured.  CPU only."""
import os

import numpy as np
import pytest

from adapter_fakes import fake_scf, packed_df
from xtddft_b200 import adapters
from xtddft_b200.synth import make_problem


def _unpack(packed, n):
    il = np.tril_indices(n)
    out = np.zeros((packed.shape[0], n, n))
    out[:, il[0], il[1]] = packed
    out[:, il[1], il[0]] = packed
    return out


@pytest.mark.parametrize("restricted", [True, False])
@pytest.mark.parametrize("xct", ["LDA", "GGA", "MGGA"])
def test_from_mean_field_object(restricted, xct):
    p = make_problem(12, 3, 2, 7, 9, 40, xctype=xct, hyb=0.3, restricted=restricted, seed=50)
    with fake_scf() as (FakeROKS, FakeUKS):
        mf = (FakeROKS if restricted else FakeUKS)(p)
        mf.with_df = packed_df(p)
        q = adapters.from_pyscf(mf, kernel="uks")
        qa = adapters.from_pyscf(mf, kernel="alda0")
        qm = adapters.from_pyscf(mf, kernel="mcol", fxc_mcol=p.fxc_mcol)
    assert (q.nao, q.nc, q.no, q.nv, q.restricted) == (p.nao, p.nc, p.no, p.nv, restricted)
    assert np.array_equal(q.mo_coeff, p.mo_coeff) and np.array_equal(q.mo_energy, p.mo_energy)
    assert np.abs(q.fock_ks - p.fock_ks).max() < 1e-12
    if restricted:
        assert np.abs(q.fock_hf - p.fock_hf).max() < 1e-12
    else:
        assert q.fock_hf is None
    assert (q.hyb, q.omega, q.alpha) == (p.hyb, p.omega, p.alpha)
    assert q.cderi is None and q.naux == p.naux and q.has_df
    assert np.array_equal(_unpack(q.cderi_packed, p.nao), p.cderi)
    assert q.xctype == xct and np.array_equal(q.ao, p.ao) and np.array_equal(q.weights, p.weights)
    assert np.array_equal(q.fxc_uks, p.fxc_uks)
    assert qa.fxc_alda0.shape == (p.ng,) and np.isfinite(qa.fxc_alda0).all()
    assert np.array_equal(qm.fxc_mcol, p.fxc_mcol)


def test_alda0_kernel_matches_the_reference(golden_dir):
    """The ALDA0 kernel the adapter builds == the one the reference's own cache_xc_kernel_sf built from the same object."""
    d = np.load(os.path.join(golden_dir, "sf_down_gga.npz"), allow_pickle=False)
    nc, no, nv, naux, ng, seed = [int(v) for v in d["params"][:6]]
    p = make_problem(nc + no + nv, nc, no, nv, naux, ng, xctype=str(d["xctype"]), hyb=float(d["hyb"]), seed=seed)
    with fake_scf() as (FakeROKS, _):
        mf = FakeROKS(p)
        mf.with_df = packed_df(p)
        q = adapters.from_pyscf(mf, kernel="alda0")
    assert np.abs(q.fxc_alda0 - d["fxc_alda0"]).max() <= 1e-14 * np.abs(d["fxc_alda0"]).max()


def test_on_disk_tensor_and_missing_df():
    p = make_problem(10, 3, 1, 6, 17, 20, xctype="LDA", hyb=0.3, seed=51)
    with fake_scf() as (FakeROKS, _):
        mf = FakeROKS(p)
        with pytest.raises(adapters.NotDensityFittedError):
            adapters.from_pyscf(mf, kernel="uks")               # no with_df, no auxbasis: refuse instead of fitting silently
        mf.with_df = packed_df(p, on_disk=True)
        q = adapters.from_pyscf(mf, kernel="none")
    assert callable(q.cderi_packed) and q.naux == 17 and q.xctype == "HF"
    rows = np.concatenate(list(q.cderi_packed()))
    assert np.array_equal(_unpack(rows, p.nao), p.cderi)


def test_stubs_do_not_leak():
    import sys
    with fake_scf():
        assert "pyscf" in sys.modules
    assert "pyscf" not in sys.modules and "cupy" not in sys.modules
