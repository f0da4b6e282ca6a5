"""Parity of the CUDA sigma path (through the C-ABI) with the oracle on seeded inputs.  Needs a B200."""
import numpy as np
import pytest

from oracle import sigma as osig
from xtddft_b200 import plan as planmod
from xtddft_b200.synth import make_problem

pytestmark = pytest.mark.gpu

RTOL = 1e-9     # north_star: sigma vectors within 1e-9 relative; measured differences are ~1e-13


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def _engine(plan, p, **kw):
    from xtddft_b200.engine import SigmaEngine
    return SigmaEngine.from_problem(plan, p, workspace_bytes=512 << 20, max_nvec=kw.pop("max_nvec", 8), **kw)


def _check(torch, eng, vind, dim, nvec=3, seed=0):
    z = np.random.default_rng(seed).standard_normal((nvec, dim))
    ref = vind(z)
    got = eng.sigma(torch.from_numpy(z).cuda()).cpu().numpy()
    err = np.abs(got - ref).max() / max(1.0, np.abs(ref).max())
    assert err < RTOL, err
    got_h = eng.sigma_host(z)
    assert np.abs(got_h - got).max() == 0.0
    return err


@pytest.mark.parametrize("xct,hyb", [("GGA", 0.2), ("LDA", 0.0), ("HF", 1.0), ("LDA", 0.3)])
@pytest.mark.parametrize("no", [1, 2])
@pytest.mark.parametrize("restricted", [True, False])
def test_xtda(torch_cuda, xct, hyb, no, restricted):
    p = make_problem(20 + no, 5, no, 15, 23, 300, xctype=xct, hyb=hyb, restricted=restricted, seed=100 + no)
    vind, hd = osig.xtda_gen_vind(p)
    pl = planmod.build_xtda_plan(p)
    eng = _engine(pl, p)
    _check(torch_cuda, eng, vind, hd.size)
    assert np.abs(eng.hdiag() - hd).max() < 1e-13
    eng.close()


@pytest.mark.parametrize("isf", [-1, 1])
@pytest.mark.parametrize("method", [0, 1, 2])
@pytest.mark.parametrize("restricted", [True, False])
def test_sf(torch_cuda, isf, method, restricted):
    p = make_problem(21, 4, 2, 15, 19, 260, xctype="GGA", hyb=0.5, restricted=restricted, seed=110)
    vind, hd = osig.sf_gen_vind(p, isf, method)
    pl = planmod.build_sf_plan(p, isf=isf, method=method)
    eng = _engine(pl, p)
    _check(torch_cuda, eng, vind, hd.size)
    eng.close()


@pytest.mark.parametrize("sa", [0, 1, 2, 3])
@pytest.mark.parametrize("remove", [False, True])
@pytest.mark.parametrize("no,nc,method", [(2, 5, 0), (3, 4, 1)])
def test_xsf_block(torch_cuda, sa, remove, no, nc, method):
    p = make_problem(nc + no + 13, nc, no, 13, 21, 280, xctype="GGA", hyb=0.4, seed=120 + no)
    vind, hd = osig.xsf_gen_vind(p, sa=sa, method=method, remove=remove, foo=0.8, fglobal=0.7)
    pl = planmod.build_sf_plan(p, isf=-1, method=method, sa=sa, layout=planmod.LAYOUT_BLOCK, remove=remove, foo=0.8, fglobal=0.7,
                               hdiag_kind="xsf")
    eng = _engine(pl, p)
    _check(torch_cuda, eng, vind, hd.size)
    assert np.abs(eng.hdiag() - hd).max() < 1e-11
    eng.close()


@pytest.mark.parametrize("x_level", [0, 3])
@pytest.mark.parametrize("remove", [False, True])
def test_xsf_gpu_order(torch_cuda, x_level, remove):
    p = make_problem(20, 5, 2, 13, 21, 280, xctype="GGA", hyb=0.4, seed=130)
    vind, hd = osig.xsf_gpu_gen_vind(p, x_level=x_level, collinear="mcol", extype=1, remove=remove, foo=0.8, fglobal=0.7)
    pl = planmod.build_sf_plan(p, isf=-1, method=1, sa=x_level, layout=planmod.LAYOUT_PYSCF, remove=remove, foo=0.8, fglobal=0.7,
                               hdiag_kind="gpu")
    eng = _engine(pl, p)
    _check(torch_cuda, eng, vind, hd.size)
    eng.close()


def test_rsh_and_many_vectors(torch_cuda):
    p = make_problem(24, 6, 2, 16, 25, 300, xctype="GGA", hyb=0.2, seed=140, omega=0.33, alpha=0.65)
    vind, hd = osig.xtda_gen_vind(p)
    eng = _engine(planmod.build_xtda_plan(p), p, max_nvec=5)
    _check(torch_cuda, eng, vind, hd.size, nvec=12)      # more vectors than max_nvec: batched internally
    eng.close()


def test_medium_size_tiles(torch_cuda):
    """Sizes that cross tile boundaries (M, N, K > 128) in every GEMM of the path."""
    p = make_problem(300, 40, 2, 258, 150, 3000, xctype="GGA", hyb=0.5, seed=150)
    vind, hd = osig.xsf_gen_vind(p, sa=3, method=0, remove=True)
    pl = planmod.build_sf_plan(p, isf=-1, method=0, sa=3, layout=planmod.LAYOUT_BLOCK, remove=True, hdiag_kind="xsf")
    eng = _engine(pl, p, max_nvec=4)
    _check(torch_cuda, eng, vind, hd.size, nvec=4)
    eng.close()
    p = make_problem(200, 30, 1, 169, 120, 2000, xctype="GGA", hyb=0.2, seed=151)
    vind, hd = osig.xtda_gen_vind(p)
    eng = _engine(planmod.build_xtda_plan(p), p, max_nvec=4)
    _check(torch_cuda, eng, vind, hd.size, nvec=3)
    eng.close()


@pytest.mark.parametrize("naux,ng", [(1, 1), (3, 129), (17, 5)])
def test_ragged_sizes(torch_cuda, naux, ng):
    """One auxiliary function / one grid point / sizes off every tile and chunk boundary; 1 and 7 vectors."""
    p = make_problem(21, 5, 2, 14, naux, ng, xctype="GGA", hyb=0.5, seed=160 + naux)
    vind, hd = osig.sf_gen_vind(p, -1, 1)
    eng = _engine(planmod.build_sf_plan(p, isf=-1, method=1), p, max_nvec=8)
    _check(torch_cuda, eng, vind, hd.size, nvec=1)
    _check(torch_cuda, eng, vind, hd.size, nvec=7, seed=2)
    eng.close()
    vind, hd = osig.xtda_gen_vind(p)
    eng = _engine(planmod.build_xtda_plan(p), p, max_nvec=8)
    _check(torch_cuda, eng, vind, hd.size, nvec=2)
    eng.close()


def test_zero_vector_and_bad_arguments(torch_cuda):
    torch = torch_cuda
    from xtddft_b200 import _lib
    p = make_problem(16, 4, 2, 10, 9, 40, xctype="LDA", hyb=0.3, seed=170)
    vind, hd = osig.sf_gen_vind(p, -1, 0)
    eng = _engine(planmod.build_sf_plan(p, isf=-1, method=0), p, max_nvec=4)
    z = torch.zeros((2, hd.size), dtype=torch.float64, device="cuda")
    assert eng.sigma(z).abs().max().item() == 0.0                       # A.0 = 0 exactly
    with pytest.raises(AssertionError):
        eng.sigma(torch.zeros((2, hd.size + 1), dtype=torch.float64, device="cuda"))
    with pytest.raises(_lib.XtdError):                                   # the C-ABI rejects nvec outside 1..max_nvec
        _lib.check(eng.lib.xtd_sigma(eng._h, 9, z.data_ptr(), z.data_ptr()), "xtd_sigma")
    assert b"nvec" in eng.lib.xtd_last_error()
    eng.close()


def test_empty_shards(torch_cuda):
    """world=4 with 2 auxiliary functions and 3 grid points: ranks 2 and 3 own no tensor slice, rank 3 no grid point; the
    partial buffers still add up to the unsharded result."""
    import ctypes as C
    from xtddft_b200 import _lib
    from xtddft_b200.engine import SigmaEngine, _as_tensor
    torch = torch_cuda
    p = make_problem(14, 3, 2, 9, 2, 3, xctype="LDA", hyb=0.5, seed=180)
    plan = planmod.build_sf_plan(p, isf=-1, method=0)
    vind, hd = osig.sf_gen_vind(p, -1, 0)
    z = torch.from_numpy(np.random.default_rng(1).standard_normal((2, hd.size))).cuda()
    acc = None
    for r in range(4):
        eng = SigmaEngine.from_problem(plan, p, workspace_bytes=256 << 20, max_nvec=4, rank=r, world=4)
        _lib.check(eng.lib.xtd_sigma_partial(eng._h, 2, C.c_void_p(z.data_ptr())), "partial")
        ptr, n = C.c_void_p(), C.c_long()
        _lib.check(eng.lib.xtd_partial_buffer(eng._h, 2, C.byref(ptr), C.byref(n)), "buffer")
        part = _as_tensor(torch, ptr.value, n.value, eng.device).clone()
        acc = part if acc is None else acc + part
        if r == 3:
            # finish on the last engine with the summed partials: must equal the oracle
            _as_tensor(torch, ptr.value, n.value, eng.device).copy_(acc)
            out = torch.empty_like(z)
            _lib.check(eng.lib.xtd_sigma_finish(eng._h, 2, C.c_void_p(out.data_ptr())), "finish")
            ref = vind(z.cpu().numpy())
            assert np.abs(out.cpu().numpy() - ref).max() < RTOL * max(1.0, np.abs(ref).max())
        eng.close()


@pytest.mark.parametrize("split", ["1", "0"])
@pytest.mark.parametrize("nc,no,nv", [(5, 1, 15), (6, 2, 17), (130, 3, 140)])
def test_gga_split_gradient_form(torch_cuda, monkeypatch, split, nc, no, nv):
    """Value + gradient kernels in both GEMM arrangements (XTD_XC_SPLIT forces one): four components on the virtual side,
    or value on the virtual side + value on the occupied side with the gradient halves streamed (run_xc).  UKS kernel on
    two channels (X-TDA) and the multicollinear kernel on one (spin flip); odd / even orbital counts, > 1 tile."""
    monkeypatch.setenv("XTD_XC_SPLIT", split)
    p = make_problem(nc + no + nv, nc, no, nv, 7, 300, xctype="GGA", hyb=0.2, seed=190 + no)
    vind, hd = osig.xtda_gen_vind(p)
    eng = _engine(planmod.build_xtda_plan(p), p, max_nvec=8)
    _check(torch_cuda, eng, vind, hd.size, nvec=3)
    _check(torch_cuda, eng, vind, hd.size, nvec=1, seed=5)
    eng.close()
    vind, hd = osig.sf_gen_vind(p, -1, 1)
    eng = _engine(planmod.build_sf_plan(p, isf=-1, method=1), p, max_nvec=8)
    _check(torch_cuda, eng, vind, hd.size, nvec=4)
    eng.close()


@pytest.mark.parametrize("nc,no,nv", [(6, 2, 17), (70, 3, 150), (33, 1, 140)])
def test_gga_split_streaming_kernels_by_batch_size(torch_cuda, monkeypatch, nc, no, nv):
    """The split-gradient streaming step has two kernels: warps over trial vectors (more than 8 vectors) and, for small batches, the
    warps of a CTA sharing the grid point by orbital range (`xc_weight_split_op_kernel`, 1 / 2 / 4 vectors per pass with a ragged last
    pass).  Every batch size class against the oracle, UKS kernel on two channels and the multicollinear kernel on one, odd / even
    occupied counts (8- and 16-byte accesses)."""
    monkeypatch.setenv("XTD_XC_SPLIT", "1")
    p = make_problem(nc + no + nv, nc, no, nv, 7, 300, xctype="GGA", hyb=0.2, seed=230 + no)
    vind, hd = osig.xtda_gen_vind(p)
    eng = _engine(planmod.build_xtda_plan(p), p, max_nvec=12)
    for nvec in (1, 2, 5, 8, 11):
        _check(torch_cuda, eng, vind, hd.size, nvec=nvec, seed=nvec)
    eng.close()
    vind, hd = osig.sf_gen_vind(p, -1, 1)
    eng = _engine(planmod.build_sf_plan(p, isf=-1, method=1), p, max_nvec=12)
    for nvec in (1, 2, 6, 9):
        _check(torch_cuda, eng, vind, hd.size, nvec=nvec, seed=nvec)
    eng.close()


@pytest.mark.parametrize("narrow", ["on", "off"])
@pytest.mark.parametrize("nc,no,nv", [(131, 3, 150), (40, 4, 33), (7, 2, 260), (20, 3, 141)])
def test_xsf_narrow_open_block(torch_cuda, monkeypatch, narrow, nc, no, nv):
    """Block-weighted exchange of the XSF Delta A with the open-shell output columns taken by the narrow-output pass
    (contract with the few open rows of Lvv first) or by the general pass (XTD_NO_NARROW); > 1 tile of closed orbitals,
    odd / even open-shell counts, removed OO vector.  nv = 260 and 141 leave 4 / 13 columns past the last full 128-column
    tile of the virtual block: that tail takes the narrow pass too."""
    if narrow == "off":
        monkeypatch.setenv("XTD_NO_NARROW", "1")
    p = make_problem(nc + no + nv, nc, no, nv, 9, 200, xctype="LDA", hyb=0.3, seed=210 + no)
    for sa, remove in [(3, True), (2, False)]:
        vind, hd = osig.xsf_gen_vind(p, sa=sa, method=0, remove=remove, foo=0.8, fglobal=0.7)
        pl = planmod.build_sf_plan(p, isf=-1, method=0, sa=sa, layout=planmod.LAYOUT_BLOCK, remove=remove, foo=0.8, fglobal=0.7,
                                   hdiag_kind="xsf")
        eng = _engine(pl, p)
        _check(torch_cuda, eng, vind, hd.size, nvec=3)
        _check(torch_cuda, eng, vind, hd.size, nvec=1, seed=3)
        eng.close()


@pytest.mark.parametrize("route", ["gemm", "stream"])
def test_xtda_coulomb_routes(torch_cuda, monkeypatch, route):
    """Full-width Coulomb blocks (X-TDA J[Da] + J[Db]) as two GEMMs over the flattened pair index, or with the streaming
    kernels (XTD_J_STREAM); odd auxiliary count, 9 vectors."""
    if route == "stream":
        monkeypatch.setenv("XTD_J_STREAM", "1")
    p = make_problem(150, 20, 2, 128, 37, 100, xctype="LDA", hyb=0.0, seed=220)
    vind, hd = osig.xtda_gen_vind(p)
    eng = _engine(planmod.build_xtda_plan(p), p, max_nvec=16)
    _check(torch_cuda, eng, vind, hd.size, nvec=9)
    eng.close()


@pytest.mark.parametrize("nc,no,nv", [(5, 1, 15), (6, 2, 17), (130, 3, 140)])
@pytest.mark.parametrize("restricted", [True, False])
def test_meta_gga(torch_cuda, nc, no, nv, restricted):
    """Kernel tables with a tau component (meta-GGA, no Laplacian): UKS kernel of X-TDA, multicollinear spin flip; the
    ALDA0 kernel of a meta-GGA has no tau part.  Odd / even occupied counts (8- and 16-byte access variants)."""
    p = make_problem(nc + no + nv, nc, no, nv, 7, 260, xctype="MGGA", hyb=0.2, restricted=restricted, seed=230 + no)
    vind, hd = osig.xtda_gen_vind(p)
    eng = _engine(planmod.build_xtda_plan(p), p, max_nvec=8)
    _check(torch_cuda, eng, vind, hd.size, nvec=3)
    eng.close()
    for method in (1, 0):
        vind, hd = osig.sf_gen_vind(p, -1, method)
        eng = _engine(planmod.build_sf_plan(p, isf=-1, method=method), p, max_nvec=8)
        _check(torch_cuda, eng, vind, hd.size, nvec=5)
        eng.close()


@pytest.mark.parametrize("method", ["xtda", "xsf"])
def test_graph_replay(torch_cuda, method):
    """Launch-bound calls are replayed as CUDA graphs from the third call with the same buffers on (eager, capture, replay):
    every call must read the CURRENT contents of the trial-vector buffer and count its launches; a different vector
    count or buffer falls back to the eager path."""
    torch = torch_cuda
    p = make_problem(26, 6, 2, 18, 15, 200, xctype="GGA", hyb=0.3, seed=240)
    if method == "xtda":
        vind, hd = osig.xtda_gen_vind(p)
        plan = planmod.build_xtda_plan(p)
    else:
        vind, hd = osig.xsf_gen_vind(p, sa=3, method=0, remove=True)
        plan = planmod.build_sf_plan(p, isf=-1, method=0, sa=3, layout=planmod.LAYOUT_BLOCK, remove=True, hdiag_kind="xsf")
    eng = _engine(plan, p, max_nvec=8)
    z = torch.zeros((3, hd.size), dtype=torch.float64, device="cuda")
    out = torch.empty_like(z)
    launches = []
    for it in range(5):
        zh = np.random.default_rng(it).standard_normal((3, hd.size))
        z.copy_(torch.from_numpy(zh))
        eng.reset_stats()
        eng.sigma(z, out)
        ref = vind(zh)
        assert np.abs(out.cpu().numpy() - ref).max() < RTOL * max(1.0, np.abs(ref).max()), it
        st = eng.stats()
        launches.append(st["launches"])
        assert st["ms"]["total"] > 0
    assert len(set(launches)) == 1 and launches[0] > 10, launches          # replays account for the recorded launches
    assert eng.stats()["ms"]["k2"] == 0.0                                   # graph replay: only the total is timed
    # other shapes / buffers still work (eager)
    _check(torch, eng, vind, hd.size, nvec=2, seed=9)
    eng.close()
