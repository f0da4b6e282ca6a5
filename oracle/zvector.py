"""Z-vector (coupled-perturbed) equation of the spin-flip-up TDA nuclear gradients -- SURVEY 8f row f3 -- restating

  * ROKS reference  xtddft/grad_hb/tdroks_sfu.py:184-333 (`grad_elec`): internal variables (:207-234), the XC pieces of the
                    right-hand side (`_contract_xc_kernel`, :59-181, collinear kernel), the Q matrix / right-hand side `w`
                    (:236-274), the ROHF orbital-Hessian `matvec` (:283-321), the `lib.solve` call (:324-327) and the W matrix
                    `im0` (:328-356)
  * UKS reference   xtddft/grad_hb/tduks_sfu.py:184-299: right-hand side (wvoa, wvob) (:205-244), `fvind` (:246-258), the
                    `ucphf.solve` call (:261-263) and the W matrix (:266-299)

on a `ProblemData` (NumPy), by the reference's AO route: back-transform the rotation blocks to AO densities, symmetrise, apply the
hermi = 1 response `vresp` (grid f_xc + J - hyb K), project on the occupied-virtual blocks.  TEST INFRASTRUCTURE ONLY (see
oracle/__init__.py).

Third-party arithmetic restated from its published definition (PySCF is not installed): `mf.gen_response(hermi=1)` is
`_gen_uhf_response`: v1 = nr_uks_fxc(dm1) + J[dm1_a + dm1_b] - hyb K[dm1_s] (+ range-separated part) -- the same expression
`oracle.sigma.response_uks` restates for XTDA.py:510-556; `ucphf.solve` (pyscf/scf/ucphf.py `solve_nos1`) solves
(e_a - e_i) z_ai + fvind(z)_ai = -h1_ai with a Krylov solver; `lib.solve` solves aop(x) = b.  Both solutions are unique, so the
oracle solves the dense systems with LAPACK.
"""
import numpy as np

from . import jk, numint
from .sigma import response_uks

es = lambda *a: np.einsum(*a, optimize=True)


def _orbitals(p):
    ca, cb = p.mo_coeff
    na, nb = p.nocc_a, p.nocc_b
    return ca[:, :na], ca[:, na:], cb[:, :nb], cb[:, nb:]


def internal_densities(p, v):
    """tdroks_sfu.py:207-214 / tduks_sfu.py:205-212: T_ab, T_ij and the AO densities of the spin-flip-up amplitudes v[nc, nv]."""
    oa, va, ob, vb = _orbitals(p)
    dvva = es("ia,ib->ab", v, v)
    doob = -es("ia,ja->ij", v, v)
    return dvva, doob, va @ dvva @ va.T, ob @ doob @ ob.T, ob @ v @ va.T


def _rhs_intermediates(p, v):
    """veff0doo (AO, both spins) and veff0mo of rhs_common, for the W matrix."""
    dvva, doob, dmzvva, dmzoob, dmt = internal_densities(p, v)
    dmoo = np.stack([dmzvva, dmzoob])
    f1oo = numint.nr_uks_fxc(p.ao, p.weights, p.fxc_uks, dmoo[:, None])[:, 0] if p.xctype != "HF" else np.zeros_like(dmoo)
    vj = jk.get_j(p.cderi, dmoo)
    veff0doo = (vj[0] + vj[1])[None] + f1oo
    vk1 = np.zeros_like(dmt)
    if p.hyb != 0.0:
        veff0doo = veff0doo - p.hyb * jk.get_k(p.cderi, dmoo)
        vk1 = p.hyb * jk.get_k(p.cderi, dmt)
    return veff0doo, p.mo_coeff[1].T @ (-vk1) @ p.mo_coeff[0]


def rhs_common(p, v):
    """The part of the Q matrix both references share (tdroks_sfu.py:241-255,257 / tduks_sfu.py:219-234), collinear kernel
    (`collinear_samples <= 0`: f1vo = 0, k1ao = 0): veff0doo = J[T_a + T_b] - hyb K[T_s] + f_xc[T]; veff0mo = C_b^T (-hyb K[X]) C_a."""
    nc, no = p.nc, p.no
    oa, va, ob, vb = _orbitals(p)
    dvva, doob, dmzvva, dmzoob, dmt = internal_densities(p, v)
    dmoo = np.stack([dmzvva, dmzoob])
    if p.xctype != "HF":
        f1oo = numint.nr_uks_fxc(p.ao, p.weights, p.fxc_uks, dmoo[:, None])[:, 0]     # _contract_xc_kernel :156-167
    else:
        f1oo = np.zeros_like(dmoo)
    vj = jk.get_j(p.cderi, dmoo)
    veff0doo = (vj[0] + vj[1])[None] + f1oo
    vk1 = np.zeros_like(dmt)
    if p.hyb != 0.0:
        veff0doo = veff0doo - p.hyb * jk.get_k(p.cderi, dmoo)
        vk1 = p.hyb * jk.get_k(p.cderi, dmt)
    wvoa = va.T @ veff0doo[0] @ oa
    wvob = vb.T @ veff0doo[1] @ ob
    veff0mo = p.mo_coeff[1].T @ (-vk1) @ p.mo_coeff[0]
    wvoa = wvoa - es("jk,jc->ck", veff0mo[:nc, :nc + no], v)
    wvob = wvob + es("ac,ka->ck", veff0mo.T[nc + no:, nc:], v)
    return wvoa, wvob, dvva, doob


def uks_rhs(p, v):
    """(wvoa [nv, nocc_a], wvob [nvir_b, nocc_b]) of tduks_sfu.py:223-234."""
    wvoa, wvob, _, _ = rhs_common(p, v)
    return wvoa, wvob


def roks_rhs(p, v):
    """w of tdroks_sfu.py:246-274, hybrid branch (:246-258); with hyb = 0 it is the intended pure-functional branch.  The SHIPPED
    pure branch cannot run: its last line (:269) writes `einsum('ac,ka->ck', doob, fockbmo[:nc, nc:])`, whose label `a` is the closed
    index of doob [nc, nc] and the virtual index of the Fock block [nc, no+nv] at once (shape error unless nc == no+nv); the hybrid
    branch's 'jk,jc->ck' (:258) is the same term written correctly and is what this restates for both."""
    nc, no = p.nc, p.no
    fa, fb = p.fock_ks
    wvoa, wvob, dvva, doob = rhs_common(p, v)
    wvoa = wvoa - es("ac,ka->ck", dvva, fa[:nc + no, nc + no:])
    wvob = wvob + es("jk,jc->ck", doob, fb[:nc, nc:])
    wvc = wvoa[:, :nc] + wvob[no:, :]
    wvo = wvoa[:, nc:]
    woc = wvob[:no, :]
    return np.hstack([wvc.ravel(), wvo.ravel(), woc.ravel()]) * 2


def sym_fock(p):
    """tdroks_sfu.py:224-234."""
    nc, no = p.nc, p.no
    fa, fb = p.fock_ks
    C, O, V = slice(0, nc), slice(nc, nc + no), slice(nc + no, None)
    sym = lambda f, r, c: (f[r, c] + f[c, r].T) / 2
    return dict(acc=sym(fa, C, C), aoc=sym(fa, O, C), avc=sym(fa, V, C), avv=sym(fa, V, V), aoo=sym(fa, O, O),
                bcc=sym(fb, C, C), bvc=sym(fb, V, C), bvv=sym(fb, V, V), bvo=sym(fb, V, O), boo=sym(fb, O, O))


def roks_matvec(p):
    """`matvec` of tdroks_sfu.py:284-321 on x = [vc (nv,nc) | vo (nv,no) | oc (no,nc)]."""
    nc, no, nv = p.nc, p.no, p.nv
    oa, va, ob, vb = _orbitals(p)
    f = sym_fock(p)

    def matvec(x):
        x = np.asarray(x, dtype=float).ravel()
        xvc = x[:nv * nc].reshape(nv, nc)
        xvo = x[nv * nc:nv * nc + nv * no].reshape(nv, no)
        xoc = x[nv * nc + nv * no:].reshape(no, nc)
        xa = np.hstack((xvc, xvo))
        xb = np.vstack((xoc, xvc))
        dma = va @ xa @ oa.T
        dmb = vb @ xb @ ob.T
        dm1 = np.stack([(dma + dma.T) / 2, (dmb + dmb.T) / 2])
        v1 = response_uks(p, dm1[:, None])[:, 0]
        v1a = va.T @ v1[0] @ oa
        v1b = vb.T @ v1[1] @ ob
        vvc = v1a[:, :nc] + v1b[no:, :]
        voc = v1b[:no, :]
        vvo = v1a[:, nc:]
        fvc = (-es("bi,ab->ai", xvc, f["avv"]) - es("bi,ab->ai", xvc, f["bvv"]) + es("aj,ji->ai", xvc, f["acc"])
               + es("aj,ji->ai", xvc, f["bcc"]) - es("ti,at->ai", xoc, f["bvo"]) + es("at,ti->ai", xvo, f["aoc"]) - vvc * 2)
        fvo = (-es("ti,ai->at", xoc, f["bvc"]) + es("ai,ti->at", xvc, f["aoc"]) - es("bt,ba->at", xvo, f["avv"])
               + es("au,tu->at", xvo, f["aoo"]) - vvo * 2)
        foc = (-es("ui,tu->ti", xoc, f["boo"]) + es("tj,ij->ti", xoc, f["bcc"]) - es("ai,at->ti", xvc, f["bvo"])
               + es("at,ai->ti", xvo, f["avc"]) - voc * 2)
        return np.hstack((fvc.ravel(), fvo.ravel(), foc.ravel()))
    return matvec


def uks_fvind(p):
    """`fvind` of tduks_sfu.py:249-258 on x = [alpha (nv, nocc_a) | beta (nvir_b, nocc_b)]."""
    na, nb, nva, nvb = p.nocc_a, p.nocc_b, p.nvir_a, p.nvir_b
    oa, va, ob, vb = _orbitals(p)

    def fvind(x):
        x = np.asarray(x, dtype=float).reshape(1, -1)
        xa = x[0, :na * nva].reshape(nva, na)
        xb = x[0, na * nva:].reshape(nvb, nb)
        dma = va @ xa @ oa.T
        dmb = vb @ xb @ ob.T
        dm1 = np.stack((dma + dma.T, dmb + dmb.T))
        v1 = response_uks(p, dm1[:, None])[:, 0]
        return np.hstack(((va.T @ v1[0] @ oa).ravel(), (vb.T @ v1[1] @ ob).ravel()))
    return fvind


def uks_gaps(p):
    """e_a - e_i in the vector order of `fvind` (pyscf/scf/ucphf.py solve_nos1)."""
    ea, eb = p.mo_energy
    na, nb = p.nocc_a, p.nocc_b
    return np.hstack([(ea[na:, None] - ea[None, :na]).ravel(), (eb[nb:, None] - eb[None, :nb]).ravel()])


def dense_operator(op, dim):
    return np.stack([op(e) for e in np.eye(dim)], axis=1)


def roks_solve(p, w):
    """tdroks_sfu.py:324-327: z with matvec(z) = w."""
    a = dense_operator(roks_matvec(p), w.size)
    return np.linalg.solve(a, w)


def uks_solve(p, wvoa, wvob):
    """tduks_sfu.py:261-263, `ucphf.solve`: (e_a - e_i) z + fvind(z) = -(wvoa, wvob); returns the stacked vector."""
    h1 = np.hstack([wvoa.ravel(), wvob.ravel()])
    a = dense_operator(uks_fvind(p), h1.size) + np.diag(uks_gaps(p))
    return np.linalg.solve(a, -h1)


def roks_w_matrix(p, v, z):
    """W matrix `im0` (AO basis) of tdroks_sfu.py:328-356 from the solution z = [zvc | zvo | zoc] of the Z-vector equation."""
    nc, no, nv = p.nc, p.no, p.nv
    na = nc + no
    oa, va, ob, vb = _orbitals(p)
    fa, fb = p.fock_ks
    f = sym_fock(p)
    dvva, doob, _, _, _ = internal_densities(p, v)
    veff0doo, veff0mo = _rhs_intermediates(p, v)
    z = np.asarray(z).ravel()
    zvc = z[:nv * nc].reshape(nv, nc)
    zvo = z[nv * nc:nv * nc + nv * no].reshape(nv, no)
    zoc = z[nv * nc + nv * no:].reshape(no, nc)
    z1a, z1b = np.hstack((zvc, zvo)), np.vstack((zoc, zvc))
    z1ao = np.stack([va @ z1a @ oa.T, vb @ z1b @ ob.T])
    veff = response_uks(p, ((z1ao + z1ao.transpose(0, 2, 1)) / 2)[:, None])[:, 0]
    n = p.nmo
    im0a, im0b = np.zeros((n, n)), np.zeros((n, n))
    im0a[:na, :na] += fa[:na, :na]
    im0b[:nc, :nc] += fb[:nc, :nc]
    im0a[:na, :na] += oa.T @ (veff0doo[0] + veff[0]) @ oa
    im0a[na:, na:] = es("ac,bc->ab", dvva, fa[na:, na:])
    im0a[na:, na:] += es("ia,ib->ab", v, veff0mo[:nc, na:])
    im0a[na:, :na] = es("aj,ij->ai", z1a, fa[:na, :na])
    im0a[na:, :na] += es("ac,ic->ai", dvva, fa[:na, na:]) * 2
    im0a[na:, :na] += es("ia,ij->aj", v, veff0mo[:nc, :na]) * 2
    im0b[:nc, :nc] += ob.T @ (veff0doo[1] + veff[1]) @ ob
    im0b[:nc, :nc] += es("ik,kj->ij", doob, fb[:nc, :nc])
    im0b[:nc, :nc] += es("ia,ja->ij", v, veff0mo[:nc, na:])
    im0b[nc:, :nc] = es("aj,ij->ai", z1b, fb[:nc, :nc])
    im0b[nc:na, :nc] += es("bt,bi->ti", zvo, f["avc"])
    return p.mo_coeff[0] @ (im0a + im0b) @ p.mo_coeff[0].T


def uks_w_matrix(p, v, z):
    """W matrix `im0` (AO basis) of tduks_sfu.py:266-299 from the stacked solution z = [z1a (nv, nocc_a) | z1b (nvir_b, nocc_b)] that
    `ucphf.solve` returns (one half of Z)."""
    nc, no, nv = p.nc, p.no, p.nv
    na, nb = nc + no, nc
    oa, va, ob, vb = _orbitals(p)
    ea, eb = p.mo_energy
    dvva, doob, _, _, _ = internal_densities(p, v)
    veff0doo, veff0mo = _rhs_intermediates(p, v)
    z = np.asarray(z).ravel()
    z1a = z[:nv * na].reshape(nv, na)
    z1b = z[nv * na:].reshape(no + nv, nb)
    z1ao = np.stack([va @ z1a @ oa.T, vb @ z1b @ ob.T])
    veff = response_uks(p, (z1ao + z1ao.transpose(0, 2, 1))[:, None])[:, 0]
    n = p.nmo
    im0a = np.zeros((n, n))
    im0a[:na, :na] = oa.T @ (veff0doo[0] + veff[0]) @ oa
    im0a[na:, na:] = es("jd,jc->dc", veff0mo[:nc, na:], v)
    im0a[:na, na:] = es("jk,jc->kc", veff0mo[:nc, :na], v) * 2
    im0b = np.zeros((n, n))
    im0b[:nc, :nc] = ob.T @ (veff0doo[1] + veff[1]) @ ob
    im0b[:nc, :nc] += es("al,ka->lk", veff0mo.T[na:, :nc], v)
    zeta_a = (ea[:, None] + ea) * 0.5
    zeta_a[:na, na:] = ea[na:]
    zeta_a[na:, :na] = ea[:na]
    dm1a = np.zeros((n, n))
    dm1a[na:, na:] = dvva
    dm1a[na:, :na] = z1a * 2
    dm1a[:na, :na] += np.eye(na)
    im0a = p.mo_coeff[0] @ (im0a + zeta_a * dm1a) @ p.mo_coeff[0].T
    zeta_b = (eb[:, None] + eb) * 0.5
    zeta_b[nc:, :nc] = eb[:nc]
    zeta_b[:nc, nc:] = eb[nc:]
    dm1b = np.zeros((n, n))
    dm1b[:nc, :nc] = doob
    dm1b[nc:, :nc] = z1b * 2
    dm1b[:nc, :nc] += np.eye(nc)
    im0b = p.mo_coeff[1] @ (im0b + zeta_b * dm1b) @ p.mo_coeff[1].T
    return im0a + im0b
