"""SURVEY 8f row f3 on the CPU: the oracle restatement of the Z-vector operator and right-hand side against the fixtures produced by the
reference's own `grad_elec` functions (tests/golden/make_golden_zvector.py), the engine plan of the operator executed by the NumPy plan
interpreter against the oracle, and the device solver's control flow (NumPy vector backend) against dense solves."""
import copy
import os

import numpy as np
import pytest

from numpy_vectors import NumpyVectors
from oracle import zvector as ozv
from plan_interp import PlanInterpreter
from xtddft_b200 import plan as planmod
from xtddft_b200.synth import make_problem
from xtddft_b200.zvector import solve_linear

TAGS = ["roks_gga_no2", "roks_lda_no3", "roks_hf_no2", "uks_gga_no2", "uks_lda_pure_no2"]


def load_case(golden_dir, tag):
    d = np.load(os.path.join(golden_dir, f"zvector_{tag}.npz"), allow_pickle=False)
    nc, no, nv, naux, ng, seed, restricted = [int(v) for v in d["params"]]
    p = make_problem(nc + no + nv, nc, no, nv, naux, ng, xctype=str(d["xctype"]), hyb=float(d["hyb"]), restricted=bool(restricted), seed=seed)
    return d, p


def _rel(a, b):
    return float(np.abs(a - b).max() / max(1.0, np.abs(b).max()))


@pytest.mark.parametrize("tag", TAGS)
def test_oracle_against_the_reference_closures(golden_dir, tag):
    d, p = load_case(golden_dir, tag)
    v = d["amp"].reshape(p.nc, p.nv)
    if p.restricted:
        op, rhs = ozv.roks_matvec(p), ozv.roks_rhs(p, v)
    else:
        op = ozv.uks_fvind(p)
        rhs = np.hstack([w.ravel() for w in ozv.uks_rhs(p, v)])
    assert _rel(np.stack([op(x) for x in d["x"]]), d["ax"]) < 1e-12
    assert _rel(rhs, d["rhs"]) < 1e-12


@pytest.mark.parametrize("tag", TAGS)
def test_plan_against_the_reference_closures(golden_dir, tag):
    d, p = load_case(golden_dir, tag)
    it = PlanInterpreter(planmod.build_zvector_plan(p, with_diag=False), p)
    assert _rel(it.sigma(d["x"]), d["ax"]) < 1e-12


@pytest.mark.parametrize("xct,hyb,kw", [("GGA", 0.2, {}), ("LDA", 0.0, {}), ("HF", 1.0, {}), ("MGGA", 0.2, {}),
                                        ("GGA", 0.25, dict(omega=0.33, alpha=0.65))])
@pytest.mark.parametrize("no", [1, 2, 3])
@pytest.mark.parametrize("restricted", [True, False])
def test_plan_against_oracle(xct, hyb, kw, no, restricted):
    p = make_problem(8 + no, 3, no, 5, 9, 30, xctype=xct, hyb=hyb, restricted=restricted, seed=70 + no, **kw)
    pl = planmod.build_zvector_plan(p, with_diag=False)
    it = PlanInterpreter(pl, p)
    op = ozv.roks_matvec(p) if restricted else ozv.uks_fvind(p)
    z = np.random.default_rng(1).standard_normal((3, pl.ext_dim))
    assert _rel(it.sigma(z), np.stack([op(x) for x in z])) < 1e-12


@pytest.mark.parametrize("restricted", [True, False])
def test_plan_preconditioner_diagonal(restricted):
    """plan.hdiag = the diagonal of the operator without its response part (Fock couplings / orbital-energy gaps)."""
    p = make_problem(10, 3, 2, 5, 9, 0, xctype="HF", hyb=1.0, restricted=restricted, seed=81)
    pl = planmod.build_zvector_plan(p)
    if restricted:
        q = copy.copy(p)
        q.cderi = np.zeros_like(p.cderi)
        ref = np.diagonal(ozv.dense_operator(ozv.roks_matvec(q), pl.ext_dim))
    else:
        ref = ozv.uks_gaps(p)
    assert np.abs(pl.hdiag - ref).max() < 1e-14


@pytest.mark.parametrize("restricted", [True, False])
def test_solver_control_flow_against_dense_solve(restricted):
    p = make_problem(12, 3, 2, 7, 11, 40, xctype="GGA", hyb=0.25, restricted=restricted, seed=83)
    pl = planmod.build_zvector_plan(p, with_diag=True)
    it = PlanInterpreter(pl, p)
    v = np.random.default_rng(5).standard_normal((p.nc, p.nv))
    v /= np.linalg.norm(v)
    if restricted:
        w = ozv.roks_rhs(p, v)
        ref, b = ozv.roks_solve(p, w), w
    else:
        wa, wb = ozv.uks_rhs(p, v)
        ref, b = ozv.uks_solve(p, wa, wb), -np.hstack([wa.ravel(), wb.ravel()])
    z, conv, cycles, res = solve_linear(it.sigma, b, pl.hdiag, tol=1e-11, max_cycle=pl.ext_dim, backend=NumpyVectors(pl.ext_dim))
    assert conv and cycles < pl.ext_dim and res < 1e-11
    assert np.abs(z - ref).max() < 1e-9 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("tag", ["roks_gga_no2", "uks_gga_no2"])
def test_from_mean_field_object(golden_dir, tag):
    """`ZVector(mf)` takes the mean-field object through `adapters.from_pyscf(kernel="uks")`: on the stand-in object the reference's
    `grad_elec` was run on, what the adapter extracts, compiled to the Z-vector plan, reproduces the reference's closure."""
    from adapter_fakes import fake_scf, packed_df
    from xtddft_b200 import adapters
    d, p = load_case(golden_dir, tag)
    with fake_scf() as (FakeROKS, FakeUKS):
        mf = (FakeROKS if p.restricted else FakeUKS)(p)
        mf.with_df = packed_df(p)
        q = adapters.problem_from_mf(mf, kernel="uks", rohf_fock="device")       # the call `ZVector(mf)` makes
    assert q.fock_hf is None
    il = np.tril_indices(p.nao)
    full = np.zeros((q.naux, p.nao, p.nao))
    full[:, il[0], il[1]] = q.cderi_packed
    full[:, il[1], il[0]] = q.cderi_packed
    q.cderi, q.cderi_packed = full, None
    it = PlanInterpreter(planmod.build_zvector_plan(q, with_diag=False), q)
    assert _rel(it.sigma(d["x"]), d["ax"]) < 1e-12


def test_solver_edge_cases():
    """Zero right-hand side, an exhausted cycle budget (reported, not raised), a non-symmetric operator, and a right-hand side that is
    an eigenvector of the preconditioned operator (one cycle)."""
    n = 40
    rng = np.random.default_rng(11)
    a = np.diag(np.linspace(1.0, 3.0, n)) + 0.05 * rng.standard_normal((n, n))          # non-symmetric, diagonally dominant
    op = lambda x: x @ a.T
    vb = NumpyVectors(n)
    z, conv, cycles, res = solve_linear(op, np.zeros(n), np.diagonal(a), tol=1e-12, backend=vb)
    assert conv and cycles == 0 and res == 0.0 and not z.any()
    b = rng.standard_normal(n)
    z, conv, cycles, res = solve_linear(op, b, np.diagonal(a), tol=1e-12, max_cycle=3, backend=vb)
    assert not conv and cycles == 3 and res > 1e-12
    z, conv, cycles, res = solve_linear(op, b, np.diagonal(a), tol=1e-12, max_cycle=n, backend=vb)
    assert conv and np.abs(z - np.linalg.solve(a, b)).max() < 1e-10
    d = np.linspace(1.0, 3.0, n)
    z, conv, cycles, res = solve_linear(lambda x: x * d, b, d, tol=1e-12, backend=vb)     # exact preconditioner
    assert conv and cycles == 1 and np.abs(z - b / d).max() < 1e-13


@pytest.mark.parametrize("tag", TAGS)
def test_rhs_and_w_matrix_plans_against_the_reference(golden_dir, tag):
    """The right-hand side and the W matrix assembled from engine plans on the whole MO space (general J / K / f_xc response of the
    relaxed difference densities and of the symmetrised Z-vector density, spin-flip exchange of the transition density) reproduce the
    `w` / `(wvoa, wvob)` and the `im0` that the reference's own `grad_elec` built (tests/golden/make_golden_zvector.py)."""
    from xtddft_b200.zvector import assemble_rhs, assemble_w, rhs_intermediates
    d, p = load_case(golden_dir, tag)
    resp = PlanInterpreter(planmod.build_mo_response_plan(p, range_separated=False), p)
    sfx = PlanInterpreter(planmod.build_mo_sf_exchange_plan(p), p) if p.hyb != 0.0 else None
    inter = rhs_intermediates(p, d["amp"].reshape(p.nc, p.nv), lambda t: resp.sigma(t.reshape(1, -1))[0],
                              (lambda x: sfx.sigma(x.reshape(1, -1))[0]) if sfx is not None else None)
    assert _rel(assemble_rhs(p, None, None, None, inter=inter), d["rhs"]) < 1e-12
    full = PlanInterpreter(planmod.build_mo_response_plan(p), p)
    assert _rel(assemble_w(p, d["z"], inter, lambda t: full.sigma(t.reshape(1, -1))[0]), d["im0"]) < 1e-12


@pytest.mark.parametrize("tag", TAGS)
def test_oracle_solution_and_w_matrix_against_the_reference(golden_dir, tag):
    d, p = load_case(golden_dir, tag)
    v = d["amp"].reshape(p.nc, p.nv)
    if p.restricted:
        z, w = ozv.roks_solve(p, d["rhs"]), ozv.roks_w_matrix(p, v, d["z"])
    else:
        n0 = p.nv * (p.nc + p.no)
        z = ozv.uks_solve(p, d["rhs"][:n0].reshape(p.nv, -1), d["rhs"][n0:].reshape(p.no + p.nv, -1))
        w = ozv.uks_w_matrix(p, v, d["z"])
    assert _rel(z, d["z"]) < 1e-11 and _rel(w, d["im0"]) < 1e-12


def test_mo_response_plan_against_oracle():
    """The whole-MO-space plan is the general `vresp`: for arbitrary (non-symmetric) MO-basis densities, all blocks, with the
    range-separated part."""
    from oracle.sigma import response_uks
    p = make_problem(11, 3, 2, 6, 9, 30, xctype="GGA", hyb=0.25, restricted=False, seed=91, omega=0.33, alpha=0.65)
    it = PlanInterpreter(planmod.build_mo_response_plan(p), p)
    t = np.random.default_rng(2).standard_normal((2, p.nmo, p.nmo))
    dm = np.stack([p.mo_coeff[s] @ t[s] @ p.mo_coeff[s].T for s in (0, 1)])
    v1 = response_uks(p, dm[:, None])[:, 0]
    ref = np.stack([p.mo_coeff[s].T @ v1[s] @ p.mo_coeff[s] for s in (0, 1)])
    assert _rel(it.sigma(t.reshape(1, -1))[0].reshape(2, p.nmo, p.nmo), ref) < 1e-12
