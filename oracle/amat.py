"""Explicit (dense) definitions of A, restating the reference's own dense builders that it uses to
validate the iterative path (SURVEY 4 "Dense-vs-iterative"):

  * X-TDA    xtddft/XTDA.py:85-398   (`XTDA.full_diag`), order CV(aa) | OV(aa) | CO(bb) | CV(bb)
  * SF-TDA   xtddft/SF_TDA.py:624-804 (`SF_TDA_down.get_Amat`), xtddft/SF_TDA.py:448-560 (`SF_TDA_up.get_Amat`)
  * XSF-TDA  xtddft/XSF_TDA.py:265-395 (`XSF_TDA.get_Amat`) and :416-427 (`remove`), block order cv | co | ov | oo

O(dim^2) memory: small problems only.  MO two-electron integrals come from the DF tensor
(`jk.mo_eri`), the grid kernel from the same cached f_xc the sigma builders use.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
import math

import numpy as np

from . import jk
from .layouts import block_dims, get_vect
from .sigma import xsf_factors, xsf_fglobal, xtda_coeffs

es = lambda *a: np.einsum(*a, optimize=True)


def _pair_density(ao, co, cv):
    """rho_ov[c,g,i,a]: phi_i phi_a and, for GGA, d(phi_i phi_a) (XTDA.py:217-224)."""
    po = es("cgp,pi->cgi", ao, co)
    pv = es("cgp,pa->cga", ao, cv)
    r = es("cgi,ga->cgia", po, pv[0])
    if ao.shape[0] > 1:
        r[1:4] += es("gi,cga->cgia", po[0], pv[1:4])
    return r


def xtda_amat(p):
    """Dense X-TDA matrix in the reference's own order (XTDA.py:277-398)."""
    assert p.restricted
    c = p.mo_coeff[0]
    nc, no, nv = p.nc, p.no, p.nv
    na, nb, nva, nvb = p.nocc_a, p.nocc_b, p.nvir_a, p.nvir_b
    oa, va, ob, vb = c[:, :na], c[:, na:], c[:, :nb], c[:, nb:]
    fa, fb = p.fock_ks
    fha, fhb = p.fock_hf
    aa = np.zeros((na, nva, na, nva))
    ab = np.zeros((na, nva, nb, nvb))
    bb = np.zeros((nb, nvb, nb, nvb))
    if p.cderi is not None:
        # (ia|jb) Coulomb, -(ij|ab) exchange (XTDA.py:142-148)
        aa += jk.mo_eri(p.cderi, oa, va, oa, va)
        bb += jk.mo_eri(p.cderi, ob, vb, ob, vb)
        ab += jk.mo_eri(p.cderi, oa, va, ob, vb)
        if p.hybrid:
            aa -= p.hyb * jk.mo_eri(p.cderi, oa, oa, va, va).transpose(0, 3, 1, 2)
            bb -= p.hyb * jk.mo_eri(p.cderi, ob, ob, vb, vb).transpose(0, 3, 1, 2)
            if p.omega != 0.0 and p.cderi_lr is not None:
                kf = p.alpha - p.hyb
                aa -= kf * jk.mo_eri(p.cderi_lr, oa, oa, va, va).transpose(0, 3, 1, 2)
                bb -= kf * jk.mo_eri(p.cderi_lr, ob, ob, vb, vb).transpose(0, 3, 1, 2)
    if p.xctype != "HF":
        wf = p.fxc_uks * p.weights
        ra = _pair_density(p.ao, oa, va)
        rb = _pair_density(p.ao, ob, vb)
        aa += es("xyg,xgia,ygjb->iajb", wf[0, :, 0], ra, ra)
        ab += es("xyg,xgia,ygjb->iajb", wf[0, :, 1], ra, rb)
        bb += es("xyg,xgia,ygjb->iajb", wf[1, :, 1], rb, rb)
    # Fock parts, PySCF index ranges
    aa += es("ij,ab->iajb", np.eye(na), fa[na:, na:]) - es("ij,ab->iajb", fa[:na, :na], np.eye(nva))
    bb += es("ij,ab->iajb", np.eye(nb), fb[nb:, nb:]) - es("ij,ab->iajb", fb[:nb, :nb], np.eye(nvb))
    # spin-adaptation corrections on the CV blocks (XTDA.py:298-307, 324-331, 389-398)
    c1, c2, c3 = xtda_coeffs(p.spin_s)
    dvv = fhb[na:, na:] - fha[na:, na:]
    dcc = fhb[:nb, :nb] - fha[:nb, :nb]
    vv = es("ij,ab->iajb", np.eye(nc), dvv)
    cc = es("ij,ab->iajb", dcc, np.eye(nv))
    aa[:nc, :, :nc, :] += c1 * vv + c2 * cc
    bb[:, no:, :, no:] += c2 * vv + c1 * cc
    ab[:nc, :, :, no:] -= c3 * (vv + cc)
    # assemble in PySCF order then permute to [CVa | OVa | COb | CVb]
    da, db = na * nva, nb * nvb
    a = np.zeros((da + db, da + db))
    a[:da, :da] = aa.reshape(da, da)
    a[:da, da:] = ab.reshape(da, db)
    a[da:, :da] = ab.reshape(da, db).T
    a[da:, da:] = bb.reshape(db, db)
    return a          # PySCF order; use layouts.order_pyscf2my to reorder


def sf_amat_pyscf(p, isf=-1, method=0):
    """Dense SF-TDA matrix in PySCF order (rows (i,a), a fastest).  Collects the terms of
    SF_TDA.py:624-735 (down) / :448-560 (up): -hyb (ij|ab) + ALDA0 kernel + Fock blocks."""
    ca, cb = p.mo_coeff
    na, nb = p.nocc_a, p.nocc_b
    fa, fb = p.fock_ks
    if isf == -1:
        co, cv, foo, fvv = ca[:, :na], cb[:, nb:], fa[:na, :na], fb[nb:, nb:]
    else:
        co, cv, foo, fvv = cb[:, :nb], ca[:, na:], fb[:nb, :nb], fa[na:, na:]
    n0, n1 = co.shape[1], cv.shape[1]
    a = np.zeros((n0, n1, n0, n1))
    if p.hybrid and p.cderi is not None:
        a -= p.hyb * jk.mo_eri(p.cderi, co, co, cv, cv).transpose(0, 3, 1, 2)
        if p.omega != 0.0 and p.cderi_lr is not None:
            a -= (p.alpha - p.hyb) * jk.mo_eri(p.cderi_lr, co, co, cv, cv).transpose(0, 3, 1, 2)
    if p.xctype != "HF" and method != 2:
        r = _pair_density(p.ao, co, cv)
        if method == 0:
            a += es("g,gia,gjb->iajb", p.fxc_alda0, r[0], r[0])
        else:
            a += es("xyg,xgia,ygjb->iajb", 2.0 * p.fxc_mcol * p.weights, r, r)
    a += es("ij,ab->iajb", np.eye(n0), fvv) - es("ij,ab->iajb", foo, np.eye(n1))
    return a.reshape(n0 * n1, n0 * n1)


def xsf_amat(p, sa=3, method=0, foo=1.0, fglobal=None, remove=False):
    """Dense XSF-TDA matrix in block order cv|co|ov|oo: SF-TDA(down) + fglobal * Delta A (XSF_TDA.py:341-393)."""
    from .layouts import pyscf_to_block_index
    nc, no, nv = p.nc, p.no, p.nv
    perm = pyscf_to_block_index(nc, no, nv)
    a_sf = sf_amat_pyscf(p, -1, method)[np.ix_(perm, perm)]
    if sa == 0 or not p.restricted:
        a = a_sf
    else:
        if fglobal is None:
            fglobal = xsf_fglobal(p, method)
        s = no / 2.0
        c = p.mo_coeff[0]
        fha, fhb = p.fock_hf
        fs = (fhb - fha) / 2
        C, O, V = slice(0, nc), slice(nc, nc + no), slice(nc + no, None)
        eri = jk.mo_eri(p.cderi, c, c, c, c)
        ic, io, iv = np.eye(nc), np.eye(no), np.eye(nv)
        d1, d2, d3 = block_dims(nc, no, nv)
        dim = d3 + no * no
        d = np.zeros((dim, dim))
        f1, f2, f3, f4 = xsf_factors(s)
        d[:d1, :d1] += (es("ij,ab->iajb", ic, fs[V, V]) + es("ji,ab->iajb", fs[C, C], iv)).reshape(d1, d1) / s
        d[d1:d2, d1:d2] += (2 * es("ji,uv->iujv", fs[C, C], io) - es("uijv->iujv", eri[O, C, C, O])).reshape(nc * no, nc * no) / (2 * s - 1)
        d[d2:d3, d2:d3] += (2 * es("uv,ab->uavb", io, fs[V, V]) - es("auvb->uavb", eri[V, O, O, V])).reshape(no * nv, no * nv) / (2 * s - 1)
        if sa > 1:
            t = f1 * (es("ij,av->iajv", ic, fhb[V, O]) - es("avji->iajv", eri[V, O, C, C])).reshape(d1, nc * no)
            d[:d1, d1:d2] += t; d[d1:d2, :d1] += t.T
            t = f1 * (-es("iv,ab->iavb", fha[C, O], iv) - es("abvi->iavb", eri[V, V, O, C])).reshape(d1, no * nv)
            d[:d1, d2:d3] += t; d[d2:d3, :d1] += t.T
            t = (es("uivb->iuvb", eri[O, C, O, V]) - es("ubvi->iuvb", eri[O, V, O, C])).reshape(nc * no, no * nv) / (2 * s - 1)
            d[d1:d2, d2:d3] += t; d[d2:d3, d1:d2] += t.T
        if sa > 2:
            t = (-(f2 - 1) * es("avwi->iawv", eri[V, O, O, C]) + (f2 / s) * es("ia,wv->iawv", fs[C, V], io)).reshape(d1, no * no)
            d[:d1, d3:] += foo * t; d[d3:, :d1] += foo * t.T
            t = (f3 * (-es("wi,uv->iuwv", fha[O, C], io) - es("uvwi->iuwv", eri[O, O, O, C]))
                 + f4 * es("iu,wv->iuwv", fhb[C, O], io)).reshape(nc * no, no * no)
            d[d1:d2, d3:] += foo * t; d[d3:, d1:d2] += foo * t.T
            t = (f3 * (es("wu,av->uawv", io, fhb[V, O]) - es("avwu->uawv", eri[V, O, O, O]))
                 - f4 * es("ua,wv->uawv", fha[O, V], io)).reshape(no * nv, no * no)
            d[d2:d3, d3:] += foo * t; d[d3:, d2:d3] += foo * t.T
        a = a_sf + fglobal * d
    if remove:
        vects = get_vect(no)
        d3 = block_dims(nc, no, nv)[2]
        t = np.eye(a.shape[0], a.shape[0] - 1)
        t[d3:, d3:] = vects
        a = t.T @ a @ t
    return a
