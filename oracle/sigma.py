"""Sigma builders  vind(zs[x,dim]) -> [x,dim]  and their preconditioner diagonals, restating

  * X-TDA            xtddft/XTDA.py:558-692  (`_gen_tda_operation`), response xtddft/XTDA.py:482-556
  * SF-TDA up/down   xtddft/SF_TDA.py:162-244 (`gen_tda_operation_sf`), response xtddft/SF_TDA.py:246-286, 855-904
  * XSF-TDA (block)  xtddft/XSF_TDA.py:1029-1290 (`gen_tda_operation_sf`), hdiag xtddft/XSF_TDA.py:859-1009
  * XSF-TDA (PySCF order, GPU class) xtddft/XSF_TDA_GPU.py:357-729 (`gen_vind`)

on a `ProblemData` (NumPy).  The algorithm is the reference's AO route: back-transform the trial
vectors to AO transition densities, apply the response (grid f_xc + DF J/K), project to the MO
basis, add Fock and spin-adaptation (Delta A) terms.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
import math

import numpy as np

from . import jk, numint
from .layouts import (block_dims, get_vect, gpu_order_compress, gpu_order_expand, join_blocks, split_blocks)

es = lambda *a: np.einsum(*a, optimize=True)


# ------------------------------------------------------------------------------------------------
# response kernels (L1 in SURVEY 1)
# ------------------------------------------------------------------------------------------------
def _k_scaled(p, dms):
    """hyb*K (+ (alpha-hyb)*K_omega for range-separated hybrids): XTDA.py:523-539, SF_TDA.py:273-276."""
    vk = jk.get_k(p.cderi, dms) * p.hyb
    if p.omega != 0.0 and p.cderi_lr is not None:
        vk = vk + jk.get_k(p.cderi_lr, dms) * (p.alpha - p.hyb)
    return vk


def response_uks(p, dms):
    """XTDA.gen_response.vind (XTDA.py:510-556): v1 = fxc[D] + J[Da]+J[Db] - hyb*K[D_s]."""
    dms = np.asarray(dms)
    if p.xctype != "HF":
        v1 = numint.nr_uks_fxc(p.ao, p.weights, p.fxc_uks, dms)
    else:
        v1 = np.zeros_like(dms)
    if p.cderi is None:
        return v1
    vj = jk.get_j(p.cderi, dms)
    v1 = v1 + (vj[0] + vj[1])[None]
    if p.hybrid:
        v1 = v1 - _k_scaled(p, dms)
    return v1


def response_sf(p, dms, method=0):
    """SF_TDA.gen_response_sf.vind (SF_TDA.py:261-281) / _gen_uhf_tda_response_sf.vind (:890-903):
    method 0 ALDA0, 1 multicollinear, 2 collinear (no grid term).  No Coulomb term in spin flip."""
    dms = np.asarray(dms)
    if p.xctype == "HF" or method == 2:
        v1 = np.zeros_like(dms)
    elif method == 0:
        v1 = numint.nr_uks_fxc_sf(p.ao, p.fxc_alda0, dms)
    elif method == 1:
        v1 = numint.nr_uks_fxc_sf_mc(p.ao, p.weights, p.fxc_mcol, dms)
    else:
        raise ValueError(method)
    if p.hybrid and p.cderi is not None:
        v1 = v1 - _k_scaled(p, dms)
    return v1


def response_jk(p, dms):
    """XSF_TDA.gen_response_sf_delta_A.vind (XSF_TDA.py:990-998): plain Coulomb and exchange images."""
    return jk.get_jk(p.cderi, np.asarray(dms))


# ------------------------------------------------------------------------------------------------
# X-TDA  
# ------------------------------------------------------------------------------------------------
def xtda_coeffs(s):
    """c1, c2, c3 of XTDA.py:637-650."""
    r = math.sqrt((s + 1.0) / s)
    return 0.5 * (1 - r + 1 / (2 * s)), 0.5 * (-1 + r + 1 / (2 * s)), 0.5 / (2 * s)


def xtda_hdiag(p):
    """XTDA.py:588-602: Fock-diagonal gaps (ROKS) or orbital-energy gaps (UKS), PySCF order."""
    na, nb = p.nocc_a, p.nocc_b
    if p.restricted:
        da, db = p.fock_ks[0].diagonal(), p.fock_ks[1].diagonal()
    else:
        da, db = p.mo_energy
    e_a = da[na:] - da[:na, None]
    e_b = db[nb:] - db[:nb, None]
    return np.hstack([e_a.ravel(), e_b.ravel()]), e_a, e_b


def xtda_gen_vind(p):
    ca, cb = p.mo_coeff
    na, nb, nva, nvb = p.nocc_a, p.nocc_b, p.nvir_a, p.nvir_b
    oa, va, ob, vb = ca[:, :na], ca[:, na:], cb[:, :nb], cb[:, nb:]
    fa, fb = p.fock_ks
    hdiag, e_a, e_b = xtda_hdiag(p)
    if p.restricted:
        fha, fhb = p.fock_hf
        c1, c2, c3 = xtda_coeffs(p.spin_s)
        dvv = fhb[na:, na:] - fha[na:, na:]          # F~b_vv - F~a_vv on the nv x nv virtual block
        dcc = fhb[:nb, :nb] - fha[:nb, :nb]          # closed-closed block

    def vind(zs):
        zs = np.asarray(zs)
        nz = len(zs)
        za = zs[:, :na * nva].reshape(nz, na, nva)
        zb = zs[:, na * nva:].reshape(nz, nb, nvb)
        dmsa = es("xov,pv,qo->xpq", za, va, oa)
        dmsb = es("xov,pv,qo->xpq", zb, vb, ob)
        v1ao = response_uks(p, np.stack([dmsa, dmsb]))
        v1a = es("xpq,qo,pv->xov", v1ao[0], oa, va)
        v1b = es("xpq,qo,pv->xov", v1ao[1], ob, vb)
        if p.restricted:
            v1a += es("xib,ab->xia", za, fa[na:, na:]) - es("xja,ij->xia", za, fa[:na, :na])
            v1b += es("xib,ab->xia", zb, fb[nb:, nb:]) - es("xja,ij->xia", zb, fb[:nb, :nb])
            a = za[:, :nb, :]            # CV(aa)
            b = zb[:, :, -nva:]          # CV(bb)
            a_vv, a_cc = es("xib,ab->xia", a, dvv), es("xja,ij->xia", a, dcc)
            b_vv, b_cc = es("xib,ab->xia", b, dvv), es("xja,ij->xia", b, dcc)
            v1a[:, :nb, :] += c1 * a_vv + c2 * a_cc - c3 * (b_vv + b_cc)
            v1b[:, :, -nva:] += c2 * b_vv + c1 * b_cc - c3 * (a_vv + a_cc)
        else:
            v1a += za * e_a
            v1b += zb * e_b
        return np.hstack([v1a.reshape(nz, -1), v1b.reshape(nz, -1)])

    return vind, hdiag


# ------------------------------------------------------------------------------------------------
# SF-TDA  
# ------------------------------------------------------------------------------------------------
def sf_gen_vind(p, isf=-1, method=0):
    ca, cb = p.mo_coeff
    na, nb = p.nocc_a, p.nocc_b
    fa, fb = p.fock_ks
    ea, eb = p.mo_energy
    if isf == -1:
        orbo, orbv = ca[:, :na], cb[:, nb:]
        hdiag = (eb[nb:, None] - ea[:na]).T.ravel()                 # SF_TDA.py:209-210 (orbital energies)
        fvv, foo = fb[nb:, nb:], fa[:na, :na]
    elif isf == 1:
        orbo, orbv = cb[:, :nb], ca[:, na:]
        hdiag = (ea[na:, None] - eb[:nb]).T.ravel()
        fvv, foo = fa[na:, na:], fb[:nb, :nb]
    else:
        raise ValueError(isf)
    n0, n1 = orbo.shape[1], orbv.shape[1]

    def vind(zs0):
        zs = np.asarray(zs0).reshape(-1, n0, n1)
        dmov = es("xov,qv,po->xpq", zs, orbv, orbo)
        v1ao = response_sf(p, dmov, method)
        vs = es("xpq,po,qv->xov", v1ao, orbo, orbv)
        vs += es("ab,xib->xia", fvv, zs) - es("ij,xja->xia", foo, zs)
        return vs.reshape(zs.shape[0], -1)

    return vind, hdiag


# ------------------------------------------------------------------------------------------------
# XSF-TDA, block layout  
# ------------------------------------------------------------------------------------------------
def xsf_factors(s):
    """factor1..4 of XSF_TDA.py:1115-1118."""
    return (math.sqrt((2 * s + 1) / (2 * s)) - 1, math.sqrt((2 * s + 1) / (2 * s - 1)),
            math.sqrt((2 * s) / (2 * s - 1)) - 1, 1 / math.sqrt(2 * s * (2 * s - 1)))


def xsf_fglobal(p, method=0, d_lda=0.3, fit=True):
    """Default global Delta-A scaling (XSF_TDA.py:1511-1518)."""
    cx = p.hyb if p.omega == 0 else p.hyb + (p.alpha - p.hyb) * math.erf(p.omega)
    f = (1 - d_lda) * cx + d_lda
    if method == 1 and fit:
        f = f * 4 * (cx - 0.5) ** 2
    return f


def xsf_j_diagonals(p):
    """(iu|iu) and (ua|ua) via unit-vector densities and get_j, XSF_TDA.py:859-913."""
    c = p.mo_coeff[0]
    nc, no, nv = p.nc, p.no, p.nv
    orbca, orbo = c[:, :nc], c[:, nc:nc + no]
    cb = p.mo_coeff[1]
    orbbo, orbvv = cb[:, nc:nc + no], cb[:, nc + no:]
    co_j = np.zeros(nc * no)
    ov_j = np.zeros(no * nv)
    for k in range(nc * no):
        t = np.zeros((nc, no)); t.flat[k] = 1
        dm = es("ov,qv,po->pq", t, orbbo, orbca)
        vj = jk.get_j(p.cderi, dm)
        co_j[k] = es("pq,pi,qu->iu", vj, orbca, orbbo).flat[k]
    for k in range(no * nv):
        t = np.zeros((no, nv)); t.flat[k] = 1
        dm = es("ov,qv,po->pq", t, orbvv, orbo)
        vj = jk.get_j(p.cderi, dm)
        ov_j[k] = es("pq,pu,qa->ua", vj, orbo, orbvv).flat[k]
    return co_j.reshape(nc, no), ov_j.reshape(no, nv)


def xsf_hdiag(p, sa, fglobal, vects=None):
    """Preconditioner diagonal in block order, XSF_TDA.py:915-961 (+ :999-1009 when the S_f=S_i OO vector is removed)."""
    nc, no, nv = p.nc, p.no, p.nv
    s = no / 2.0
    da, db = p.fock_ks[0].diagonal(), p.fock_ks[1].diagonal()
    h = db[p.nocc_b:][None, :] - da[:p.nocc_a, None]              # [(c,o), (o,v)]
    h = h.copy()
    if sa > 0:
        ds = ((p.fock_hf[1] - p.fock_hf[0]) * 0.5).diagonal()
        h[:nc, no:] += fglobal * (ds[nc + no:] + ds[:nc, None]) / s
        co_j, ov_j = xsf_j_diagonals(p)
        h[:nc, :no] += fglobal * (2.0 * ds[:nc, None] - co_j) / (2 * s - 1)
        h[nc:, no:] += fglobal * (2.0 * ds[nc + no:] - ov_j) / (2 * s - 1)
    hd = np.hstack([h[:nc, no:].ravel(), h[:nc, :no].ravel(), h[nc:, no:].ravel(), h[nc:, :no].ravel()])
    if vects is not None:
        d3 = block_dims(nc, no, nv)[2]
        hd = np.hstack([hd[:d3], es("x,xy,xy->y", hd[d3:], vects, vects)])
    return hd


def xsf_gen_vind(p, sa=3, method=0, remove=True, foo=1.0, fglobal=None):
    """Block-layout XSF-TDA / USF-TDA sigma build (XSF_TDA.py:1131-1276)."""
    nc, no, nv = p.nc, p.no, p.nv
    ca, cb = p.mo_coeff
    orbca, orbo = ca[:, :nc], ca[:, nc:nc + no]
    orbbo, orbvv = cb[:, nc:nc + no], cb[:, nc + no:]
    fa, fb = p.fock_ks
    fa_cc, fa_co, fa_oc, fa_oo = fa[:nc, :nc], fa[:nc, nc:nc + no], fa[nc:nc + no, :nc], fa[nc:nc + no, nc:nc + no]
    fb_oo, fb_ov = fb[nc:nc + no, nc:nc + no], fb[nc:nc + no, nc + no:]
    fb_vo, fb_vv = fb[nc + no:, nc:nc + no], fb[nc + no:, nc + no:]
    if fglobal is None:
        fglobal = xsf_fglobal(p, method)
    vects = get_vect(no) if remove else None
    hdiag = xsf_hdiag(p, sa, fglobal, vects)
    s = no / 2.0
    if sa > 0:
        fha, fhb = p.fock_hf
        fs = (fhb - fha) * 0.5
        f1, f2, f3, f4 = xsf_factors(s)
        fs_cc, fs_vv, fs_cv = fs[:nc, :nc], fs[nc + no:, nc + no:], fs[:nc, nc + no:]
        fhb_vo = fhb[nc + no:, nc:nc + no]
        fha_oc = fha[nc:nc + no, :nc]
        fha_co = fha[:nc, nc:nc + no]
        fhb_co = fhb[:nc, nc:nc + no]
        fha_vo = fha[nc + no:, nc:nc + no]

    def project(v):
        return (es("xpq,pi,qa->xia", v, orbca, orbvv), es("xpq,pi,qu->xiu", v, orbca, orbbo),
                es("xpq,pu,qa->xua", v, orbo, orbvv), es("xpq,pu,qv->xuv", v, orbo, orbbo))

    def vind(zs0):
        cv, co, ov, oo = split_blocks(zs0, nc, no, nv, vects)
        d_cv = es("xia,qa,pi->xpq", cv, orbvv, orbca)
        d_co = es("xiu,qu,pi->xpq", co, orbbo, orbca)
        d_ov = es("xua,qa,pu->xpq", ov, orbvv, orbo)
        d_oo = es("xuv,qv,pu->xpq", oo, orbbo, orbo)
        v1ao = response_sf(p, d_cv + d_co + d_ov + d_oo, method)
        vs_cv, vs_co, vs_ov, vs_oo = project(v1ao)
        vs_cv += es("xiu,ua->xia", co, fb_ov) + es("xib,ba->xia", cv, fb_vv) - es("ij,xja->xia", fa_cc, cv) - es("iu,xua->xia", fa_co, ov)
        vs_co += es("xiv,vu->xiu", co, fb_oo) + es("xia,au->xiu", cv, fb_vo) - es("ij,xju->xiu", fa_cc, co) - es("iv,xvu->xiu", fa_co, oo)
        vs_ov += es("xuv,va->xua", oo, fb_ov) + es("xub,ba->xua", ov, fb_vv) - es("ui,xia->xua", fa_oc, cv) - es("uv,xva->xua", fa_oo, ov)
        vs_oo += es("xuw,wv->xuv", oo, fb_oo) + es("xua,av->xuv", ov, fb_vo) - es("ui,xiv->xuv", fa_oc, co) - es("uw,xwv->xuv", fa_oo, oo)
        if sa > 0:
            x = cv.shape[0]
            vj, vk = response_jk(p, np.concatenate([d_cv, d_co, d_ov, d_oo]))
            k_cv, k_co, k_ov, k_oo = vk[:x], vk[x:2 * x], vk[2 * x:3 * x], vk[3 * x:]
            j_co, j_ov = vj[x:2 * x], vj[2 * x:3 * x]
            _, co_co_j, ov_co_j, _ = project(j_co)
            _, co_ov_j, ov_ov_j, _ = project(j_ov)
            _, co_cv_k, ov_cv_k, oo_cv_k = project(k_cv)
            cv_co_k, _, ov_co_k, oo_co_k = project(k_co)
            cv_ov_k, co_ov_k, _, oo_ov_k = project(k_ov)
            cv_oo_k, co_oo_k, ov_oo_k, _ = project(k_oo)
            dcv = (es("ab,xib->xia", fs_vv, cv) + es("ji,xja->xia", fs_cc, cv)) / s
            dco = (-co_co_j + 2.0 * es("ji,xju->xiu", fs_cc, co)) / (2 * s - 1)
            dov = (-ov_ov_j + 2.0 * es("ab,xub->xua", fs_vv, ov)) / (2 * s - 1)
            doo = np.zeros_like(oo)
            if sa > 1:
                dcv += f1 * (-cv_co_k + es("av,xiv->xia", fhb_vo, co))
                dco += f1 * (-co_cv_k + es("av,xja->xjv", fhb_vo, cv))
                dcv += f1 * (-cv_ov_k - es("vi,xva->xia", fha_oc, ov))
                dov += f1 * (-ov_cv_k - es("vi,xib->xvb", fha_oc, cv))
                dco += (co_ov_j - co_ov_k) / (2 * s - 1)
                dov += (ov_co_j - ov_co_k) / (2 * s - 1)
            if sa > 2:
                tr_oo = es("xvv->x", oo)
                eye = np.eye(no)
                dcv += foo * (-(f2 - 1) * cv_oo_k + (f2 / s) * fs_cv[None] * tr_oo[:, None, None])
                doo += foo * (-(f2 - 1) * oo_cv_k + (f2 / s) * eye[None] * es("ia,xia->x", fs_cv, cv)[:, None, None])
                dco += foo * (f3 * (-co_oo_k - es("iw,xwu->xiu", fha_co, oo)) + f4 * fhb_co[None] * tr_oo[:, None, None])
                doo += foo * (f3 * (-oo_co_k - es("iw,xiv->xwv", fha_co, co)) + f4 * eye[None] * es("iu,xiu->x", fhb_co, co)[:, None, None])
                dov += foo * (f3 * (-ov_oo_k + es("av,xuv->xua", fhb_vo, oo)) - f4 * fha_vo.T[None] * tr_oo[:, None, None])
                doo += foo * (f3 * (-oo_ov_k + es("av,xwa->xwv", fhb_vo, ov)) - f4 * eye[None] * es("au,xua->x", fha_vo, ov)[:, None, None])
            vs_cv += fglobal * dcv
            vs_co += fglobal * dco
            vs_ov += fglobal * dov
            vs_oo += fglobal * doo
        return join_blocks(vs_cv, vs_co, vs_ov, vs_oo, vects)

    return vind, hdiag


# ------------------------------------------------------------------------------------------------
# XSF-TDA in PySCF vector order (the reference's GPU class)
# ------------------------------------------------------------------------------------------------
def xsf_gpu_hdiag(p, extype=1, remove=True):
    """XSF_TDA_GPU.py:385-439: Fock-diagonal (ROKS) / orbital-energy (UKS) gaps; removed layout by `oo @ vects`."""
    nc, no, nv = p.nc, p.no, p.nv
    if p.restricted:
        da, db = p.fock_ks[0].diagonal(), p.fock_ks[1].diagonal()
    else:
        da, db = p.mo_energy
    if extype == 0:
        return (da[p.nocc_a:] - db[:p.nocc_b, None]).ravel()
    h = (db[p.nocc_b:] - da[:p.nocc_a, None])
    if not remove:
        return h.ravel()
    vects = get_vect(no)
    oo = h[nc:, :no].ravel()
    new_oo = oo @ vects
    full = h.ravel().copy()
    idx = np.arange(h.size).reshape(h.shape)
    oo_pos = idx[nc:, :no].ravel()
    full[oo_pos[:-1]] = new_oo
    return np.delete(full, oo_pos[-1])


def xsf_gpu_gen_vind(p, x_level=3, collinear="alda0", extype=1, remove=True, foo=1.0, fglobal=None):
    """XSF_TDA_GPU.gen_vind.vind (XSF_TDA_GPU.py:478-727): same mathematics as the block builder, PySCF order."""
    nc, no, nv = p.nc, p.no, p.nv
    method = {"alda0": 0, "mcol": 1, "col": 2}[collinear]
    ca, cb = p.mo_coeff
    na, nb = p.nocc_a, p.nocc_b
    fa, fb = p.fock_ks
    if not p.restricted:
        x_level = 0
    if extype == 0:
        orbo, orbv = cb[:, :nb], ca[:, na:]
        remove = False
    else:
        orbo, orbv = ca[:, :na], cb[:, nb:]
    n0, n1 = orbo.shape[1], orbv.shape[1]
    hdiag = xsf_gpu_hdiag(p, extype, remove)
    vects = get_vect(no) if (remove and extype == 1) else None
    if fglobal is None:
        fglobal = xsf_fglobal(p, method)
    s = no / 2.0
    if x_level > 0 and extype == 1:
        fha, fhb = p.fock_hf
        f1, f2, f3, f4 = xsf_factors(s)
    if not p.restricted:
        ea, eb = p.mo_energy
        e_ia = (ea[na:] - eb[:nb, None]) if extype == 0 else (eb[nb:] - ea[:na, None])

    def proj(v):
        return es("xpq,qo,pv->xov", v, orbo, orbv)

    def vind(zs0):
        zs0 = np.asarray(zs0)
        full = gpu_order_expand(zs0, nc, no, nv, vects) if vects is not None else zs0
        zs = full.reshape(-1, n0, n1)
        dms = es("xov,pv,qo->xpq", zs, orbv, orbo)                # mo1 = z Cv^T ; D = mo1 Co^T (:506-507)
        v1 = proj(response_sf(p, dms, method))
        if p.restricted:
            if extype == 0:
                v1 += es("ab,xib->xia", fa[na:, na:], zs) - es("ij,xja->xia", fb[:nb, :nb], zs)
            else:
                v1 += es("ab,xib->xia", fb[nb:, nb:], zs) - es("ij,xja->xia", fa[:na, :na], zs)
                if x_level > 0:
                    cv1, co1, ov1, oo1 = zs[:, :nc, no:], zs[:, :nc, :no], zs[:, nc:, no:], zs[:, nc:, :no]
                    ob_v, ob_o = orbv[:, no:], orbv[:, :no]
                    oa_c, oa_o = orbo[:, :nc], orbo[:, nc:]
                    d4 = np.stack([es("xov,pv,qo->xpq", cv1, ob_v, oa_c), es("xov,pv,qo->xpq", co1, ob_o, oa_c),
                                   es("xov,pv,qo->xpq", ov1, ob_v, oa_o), es("xov,pv,qo->xpq", oo1, ob_o, oa_o)])
                    vj, vk = response_jk(p, d4)
                    k_cv, k_co, k_ov, k_oo = proj(vk[0]), proj(vk[1]), proj(vk[2]), proj(vk[3])
                    j_co, j_ov = proj(vj[1]), proj(vj[2])
                    C, O, V = slice(0, nc), slice(nc, None), slice(no, None)
                    o_ = slice(0, no)
                    dfv = fhb[nc + no:, nc + no:] - fha[nc + no:, nc + no:]
                    dfc = fhb[:nc, :nc] - fha[:nc, :nc]
                    v1[:, C, V] += fglobal * (es("ab,xib->xia", dfv, cv1) + es("ji,xja->xia", dfc, cv1)) / (2 * s)
                    v1[:, C, o_] += fglobal * (-j_co[:, C, o_] + es("ji,xju->xiu", dfc, co1)) / (2 * s - 1)
                    v1[:, O, V] += fglobal * (-j_ov[:, O, V] + es("ab,xub->xua", dfv, ov1)) / (2 * s - 1)
                    if x_level > 1:
                        fhb_vo, fha_oc = fhb[nc + no:, nc:nc + no], fha[nc:nc + no, :nc]
                        v1[:, C, V] += fglobal * f1 * (-k_co[:, C, V] + es("av,xiv->xia", fhb_vo, co1))
                        v1[:, C, o_] += fglobal * f1 * (-k_cv[:, C, o_] + es("av,xja->xjv", fhb_vo, cv1))
                        v1[:, C, V] += fglobal * f1 * (-k_ov[:, C, V] - es("vi,xva->xia", fha_oc, ov1))
                        v1[:, O, V] += fglobal * f1 * (-k_cv[:, O, V] - es("vi,xib->xvb", fha_oc, cv1))
                        v1[:, C, o_] += fglobal * (j_ov[:, C, o_] - k_ov[:, C, o_]) / (2 * s - 1)
                        v1[:, O, V] += fglobal * (j_co[:, O, V] - k_co[:, O, V]) / (2 * s - 1)
                    if x_level > 2:
                        eye = np.eye(no)
                        dcv = fhb[:nc, nc + no:] - fha[:nc, nc + no:]
                        fha_co, fhb_co = fha[:nc, nc:nc + no], fhb[:nc, nc:nc + no]
                        fha_vo = fha[nc + no:, nc:nc + no]
                        tr = es("xvv->x", oo1)[:, None, None]
                        v1[:, C, V] += fglobal * foo * (-(f2 - 1) * k_oo[:, C, V] + (f2 / (2 * s)) * dcv[None] * tr)
                        v1[:, O, o_] += fglobal * foo * (-(f2 - 1) * k_cv[:, O, o_]
                                                          + (f2 / (2 * s)) * eye[None] * es("ia,xia->x", dcv, cv1)[:, None, None])
                        v1[:, C, o_] += fglobal * foo * (f3 * (-k_oo[:, C, o_] - es("iw,xwu->xiu", fha_co, oo1)) + f4 * fhb_co[None] * tr)
                        v1[:, O, o_] += fglobal * foo * (f3 * (-k_co[:, O, o_] - es("iw,xiv->xwv", fha_co, co1))
                                                          + f4 * eye[None] * es("iu,xiu->x", fhb_co, co1)[:, None, None])
                        v1[:, O, V] += fglobal * foo * (f3 * (-k_oo[:, O, V] + es("av,xuv->xua", fhb_vo, oo1)) - f4 * fha_vo.T[None] * tr)
                        v1[:, O, o_] += fglobal * foo * (f3 * (-k_ov[:, O, o_] + es("av,xwa->xwv", fhb_vo, ov1))
                                                          - f4 * eye[None] * es("au,xua->x", fha_vo, ov1)[:, None, None])
        else:
            v1 += zs * e_ia
        hx = v1.reshape(zs.shape[0], -1)
        if vects is not None:
            hx = gpu_order_compress(hx, nc, no, nv, vects)
        return hx

    return vind, hdiag
