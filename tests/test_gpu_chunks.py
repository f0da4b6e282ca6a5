"""Parity of the code path every BASELINE-size run takes -- several auxiliary-function chunks in the exchange build
(`run_k`), several grid chunks in the grid path (`run_xc`), the full-size orbital shapes of configs 3 / 4 / 5 (tail-tile
layouts, narrow-output passes, `tail_w`) -- against the oracle, through the C-ABI.  The oracle finishes these in seconds
because only the orbital dimensions are at full size (a handful of auxiliary functions, ~2000 grid points) or only the
chunked dimension is long.  Also: an engine built and run on a non-blocking side stream, and the multicollinear
contraction fixtures of the reference against the CUDA path.  Needs a B200."""
import os

import numpy as np
import pytest

from oracle import sigma as osig
from xtddft_b200 import plan as planmod
from xtddft_b200.synth import make_problem

pytestmark = pytest.mark.gpu

RTOL = 1e-9     # north_star: sigma vectors within 1e-9 relative


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def _rel(got, ref):
    return float(np.abs(got - ref).max() / max(1.0, np.abs(ref).max()))


def _cases(p):
    return {
        "sf": (lambda: osig.sf_gen_vind(p, -1, 0), lambda: planmod.build_sf_plan(p, isf=-1, method=0)),
        "sf_mcol": (lambda: osig.sf_gen_vind(p, -1, 1), lambda: planmod.build_sf_plan(p, isf=-1, method=1)),
        "xtda": (lambda: osig.xtda_gen_vind(p), lambda: planmod.build_xtda_plan(p)),
        "xsf": (lambda: osig.xsf_gen_vind(p, sa=3, method=0, remove=True, foo=0.8, fglobal=0.7),
                lambda: planmod.build_sf_plan(p, isf=-1, method=0, sa=3, layout=planmod.LAYOUT_BLOCK, remove=True, foo=0.8,
                                              fglobal=0.7, hdiag_kind="xsf")),
    }


@pytest.mark.parametrize("method", ["sf", "xtda", "xsf"])
def test_many_aux_chunks(torch_cuda, monkeypatch, method):
    """naux = 61 with at most 7 auxiliary functions per chunk: 9 chunks (the last one short) in every exchange term,
    on top of several tiles in every GEMM dimension."""
    from xtddft_b200.engine import SigmaEngine
    monkeypatch.setenv("XTD_CHUNK_AUX", "7")
    p = make_problem(200, 30, 2, 168, 61, 700, xctype="GGA", hyb=0.4, seed=300)
    mk_vind, mk_plan = _cases(p)[method]
    vind, hd = mk_vind()
    eng = SigmaEngine.from_problem(mk_plan(), p, workspace_bytes=256 << 20, max_nvec=6)
    z = np.random.default_rng(1).standard_normal((5, hd.size))
    got = eng.sigma(torch_cuda.from_numpy(z).cuda()).cpu().numpy()
    assert eng.last_chunks()[0] == 9
    assert _rel(got, vind(z)) < RTOL
    eng.close()


def test_aux_chunks_from_small_workspace(torch_cuda):
    """The 64 MiB minimum workspace forces the chunking by itself (no test knob): naux = 500 at N = 200, 6 vectors."""
    from xtddft_b200.engine import SigmaEngine
    p = make_problem(200, 30, 2, 168, 500, 300, xctype="LDA", hyb=0.5, seed=301)
    vind, hd = osig.sf_gen_vind(p, -1, 0)
    eng = SigmaEngine.from_problem(planmod.build_sf_plan(p, isf=-1, method=0), p, workspace_bytes=64 << 20, max_nvec=6)
    z = np.random.default_rng(2).standard_normal((6, hd.size))
    got = eng.sigma(torch_cuda.from_numpy(z).cuda()).cpu().numpy()
    assert eng.last_chunks()[0] >= 3, eng.last_chunks()
    assert _rel(got, vind(z)) < RTOL
    eng.close()


@pytest.mark.parametrize("method", ["sf", "sf_mcol", "xtda", "xsf"])
def test_many_grid_chunks(torch_cuda, monkeypatch, method):
    """ng = 1700 in chunks of 256 points: 7 grid chunks (the last one ragged) for the ALDA0, multicollinear and UKS
    kernels, in both GEMM arrangements where the kernel has gradients."""
    from xtddft_b200.engine import SigmaEngine
    monkeypatch.setenv("XTD_CHUNK_GRID", "256")
    p = make_problem(40, 8, 2, 30, 9, 1700, xctype="GGA", hyb=0.3, seed=310)
    mk_vind, mk_plan = _cases(p)[method]
    vind, hd = mk_vind()
    z = np.random.default_rng(3).standard_normal((3, hd.size))
    ref = vind(z)
    forms = ("0", "1") if method in ("sf_mcol", "xtda") else ("0",)
    for split in forms:
        monkeypatch.setenv("XTD_XC_SPLIT", split)
        eng = SigmaEngine.from_problem(mk_plan(), p, workspace_bytes=128 << 20, max_nvec=4)
        got = eng.sigma(torch_cuda.from_numpy(z).cuda()).cpu().numpy()
        assert eng.last_chunks()[1] == 7, eng.last_chunks()
        assert _rel(got, ref) < RTOL, split
        eng.close()


def test_grid_chunks_long_grid(torch_cuda):
    """N = 20 with 200 000 grid points: the 65 536-point chunk cap of run_xc alone gives >= 4 chunks."""
    from xtddft_b200.engine import SigmaEngine
    p = make_problem(20, 5, 2, 13, 5, 200000, xctype="GGA", hyb=0.3, seed=311)
    for method in ("sf", "xtda"):
        mk_vind, mk_plan = _cases(p)[method]
        vind, hd = mk_vind()
        z = np.random.default_rng(4).standard_normal((2, hd.size))
        eng = SigmaEngine.from_problem(mk_plan(), p, workspace_bytes=512 << 20, max_nvec=4)
        got = eng.sigma(torch_cuda.from_numpy(z).cuda()).cpu().numpy()
        assert eng.last_chunks()[1] >= 4, eng.last_chunks()
        assert _rel(got, vind(z)) < RTOL
        eng.close()


# BASELINE configs 3 / 4 / 5: the ORBITAL shapes at full size (nocc 176 / 137 / 277, nvir 1552 / 822 / 1777 with their pad
# orbitals: tail tiles, narrow-output passes, tail_w), with few auxiliary functions and grid points so the oracle finishes.
FULL_SHAPES = {
    "cfg5_sf": dict(nao=2052, nc=275, no=2, nv=1775, naux=5, ng=1536, hyb=0.5, method="sf"),
    "cfg4_xtda": dict(nao=958, nc=136, no=1, nv=821, naux=6, ng=2048, hyb=0.2, method="xtda"),
    "cfg3_xsf": dict(nao=1725, nc=173, no=3, nv=1549, naux=3, ng=1536, hyb=0.25, method="xsf"),
}


@pytest.mark.parametrize("slices", [0, 6])
@pytest.mark.parametrize("name", list(FULL_SHAPES))
def test_full_orbital_shapes(torch_cuda, monkeypatch, name, slices):
    """slices = 0: FP64 DMMA GEMMs everywhere (the chunk loops of run_k / run_xc); slices = 6: the engine's default at these
    sizes -- uniform-weight exchange and one-component grid GEMMs emulated on the INT8 tensor cores (block-weighted XSF
    exchange terms and value + gradient grid kernels stay on the DMMA GEMM)."""
    from xtddft_b200.engine import SigmaEngine
    c = FULL_SHAPES[name]
    monkeypatch.setenv("XTD_CHUNK_AUX", "2")            # 2-3 aux chunks
    monkeypatch.setenv("XTD_CHUNK_GRID", "640")         # 3-4 grid chunks
    p = make_problem(c["nao"], c["nc"], c["no"], c["nv"], c["naux"], c["ng"], xctype="GGA", hyb=c["hyb"], seed=320)
    mk_vind, mk_plan = _cases(p)[c["method"]]
    vind, hd = mk_vind()
    z = np.random.default_rng(5).standard_normal((2, hd.size))
    ref = vind(z)
    eng = SigmaEngine.from_problem(mk_plan(), p, workspace_bytes=4 << 30, max_nvec=4, df_chunk=2, exchange_slices=slices)
    got = eng.sigma(torch_cuda.from_numpy(z).cuda()).cpu().numpy()
    ca, cg = eng.last_chunks()
    assert ca >= 2 and (cg >= 3 or slices), (ca, cg)
    assert _rel(got, ref) < RTOL
    assert np.abs(eng.hdiag() - hd).max() < 1e-10
    eng.close()


@pytest.mark.parametrize("method", ["xtda", "xsf"])
def test_side_stream(torch_cuda, method):
    """Engine built and run inside a non-blocking side stream: every zero fill, upload and kernel is ordered on that
    stream (a legacy-stream memset would not be)."""
    torch = torch_cuda
    from xtddft_b200.engine import SigmaEngine
    p = make_problem(150, 25, 2, 123, 40, 900, xctype="GGA", hyb=0.3, seed=330)
    mk_vind, mk_plan = _cases(p)[method]
    vind, hd = mk_vind()
    z = np.random.default_rng(6).standard_normal((3, hd.size))
    ref = vind(z)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        eng = SigmaEngine.from_problem(mk_plan(), p, workspace_bytes=256 << 20, max_nvec=4)
        zt = torch.from_numpy(z).cuda()
        for _ in range(3):                               # eager, captured, replayed
            out = eng.sigma(zt)
        side.synchronize()
        got = out.cpu().numpy()
    assert _rel(got, ref) < RTOL
    eng.close()


@pytest.mark.parametrize("name,xct,seed", [("sf_mcol_contraction.npz", "GGA", 25), ("sf_mcol_contraction_mgga.npz", "MGGA", 27)])
def test_mcol_contraction_fixture(torch_cuda, golden_dir, name, xct, seed):
    """The reference's own multicollinear contraction (`nr_uks_fxc_sf_tda_mc`, SF_TDA.py:976-1047) against the CUDA grid
    path: for a hybrid-free problem without Fock terms sigma is exactly Co^T V[D(z)] Cv, so projecting the fixture's
    AO potentials must reproduce what the engine adds to the orbital-energy part."""
    from xtddft_b200.engine import SigmaEngine
    d = np.load(os.path.join(golden_dir, name), allow_pickle=False)
    p = make_problem(9, 3, 2, 4, 10, 36, xctype=xct, hyb=0.5, seed=seed)
    co, cv = p.mo_coeff[0][:, :p.nocc_a], p.mo_coeff[1][:, p.nocc_b:]
    # the trial vectors that produced the fixture's densities: D = Co z Cv^T with orthonormal C
    z = np.einsum("xpq,po,qv->xov", d["dms"], co, cv).reshape(2, -1)
    assert np.abs(np.einsum("xov,qv,po->xpq", z.reshape(2, p.nocc_a, p.nvir_b), cv, co) - d["dms"]).max() < 1e-12
    grid_ref = np.einsum("xpq,po,qv->xov", d["v"], co, cv).reshape(2, -1)
    full = SigmaEngine.from_problem(planmod.build_sf_plan(p, isf=-1, method=1), p, workspace_bytes=128 << 20, max_nvec=4)
    nogrid = SigmaEngine.from_problem(planmod.build_sf_plan(p, isf=-1, method=2), p, workspace_bytes=128 << 20, max_nvec=4)
    got = full.sigma_host(z) - nogrid.sigma_host(z)
    assert _rel(got, grid_ref) < RTOL
    full.close()
    nogrid.close()


def test_rohf_fock_difference_on_device(torch_cuda):
    """Setup-stage piece: F_beta^HF - F_alpha^HF = K[D_open] accumulated from the streamed tensor in the MO
    basis, against the oracle's `get_k` on the open-shell density; then the full flow -- a ROKS problem WITHOUT fock_hf gets its
    spin-adaptation couplings from the device and must give the sigma vectors of the same problem with fock_hf = [0, K] supplied."""
    from oracle import jk
    from xtddft_b200.engine import SigmaEngine
    p = make_problem(40, 8, 3, 29, 21, 200, xctype="LDA", hyb=0.3, seed=340)
    kref = jk.rohf_fock_difference(p.cderi, p.mo_coeff[0], np.arange(p.nc, p.nc + p.no))
    p.fock_hf = np.stack([np.zeros_like(kref), kref])
    builder = lambda q: planmod.build_sf_plan(q, isf=-1, method=0, sa=3, layout=planmod.LAYOUT_BLOCK, remove=True, hdiag_kind="xsf")
    vind, hd = osig.xsf_gen_vind(p, sa=3, method=0, remove=True)
    z = np.random.default_rng(8).standard_normal((3, hd.size))
    ref = vind(z)
    import copy
    q = copy.copy(p)
    q.fock_hf = None
    for world in (1, 2):
        q.fock_hf = None
        parts = []
        for r in range(world):
            eng = SigmaEngine.from_problem(builder(p), q if world == 1 else copy.copy(q), workspace_bytes=256 << 20, max_nvec=4, df_chunk=8,
                                           plan_builder=builder, rank=r, world=world)
            if world == 1:
                assert np.abs(q.fock_hf[1] - kref).max() < 1e-11 * np.abs(kref).max()
                assert _rel(eng.sigma_host(z), ref) < RTOL
                assert np.abs(eng.hdiag() - hd).max() < 1e-10
            else:
                parts.append(eng.kopen())
            eng.close()
        if world == 2:          # aux shards of K[D_open] add up (the all-reduce of the real multi-rank run)
            assert np.abs(parts[0] + parts[1] - kref).max() < 1e-11 * np.abs(kref).max()
