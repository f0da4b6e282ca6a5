"""TEST INFRASTRUCTURE (CPU oracle) -- post-Davidson property pass of the reference, restated in NumPy.

SURVEY 8(f) row f1.  Every function cites the reference lines it follows; inputs that the reference obtains from
libcint (`mol.intor*`) are plain arrays here.

  xtda_transition_moments / xtda_osc_str / xtda_rot_str    xtddft/XTDA.py:838-882
  tdm_r / osc_matrix_r        state-to-state transition dipoles, ROKS reference    xtddft/XSF_TDA.py:481-592,
                                                                                  xtddft/XSF_TDA_GPU.py:993-1116
  tdm_u / osc_matrix_u        ... UKS reference                                   xtddft/XSF_TDA.py:435-478,
                                                                                  xtddft/XSF_TDA_GPU.py:943-991
  delta_s2_u                  D<S^2> label of a spin-flip-down state on a UKS reference   xtddft/XSF_TDA.py:613-649
  delta_s2_roks_sf            ... on a ROKS reference (SA=0)                      xtddft/XSF_TDA.py:771-779
"""
from __future__ import annotations

import math

import numpy as np

CGS2AU = 1.0 / (235.7220 * 2)          # xtddft/utils/unit.py:10


# ---------------------------------------------------------------------------------------------------------
# X-TDA: ground -> excited transition moments (vectors in PySCF order [alpha (i,a) | beta (i,a)])
# ---------------------------------------------------------------------------------------------------------
def xtda_transition_moments(p, x_pyscf: np.ndarray, ints_ao: np.ndarray) -> np.ndarray:
    """trans[s, k] = sum_ia <i|O_k|a>_alpha X^alpha_s[i,a] + the beta term  (XTDA.py:849-857, 865-880).
    x_pyscf: [nstates, dim];  ints_ao: [k, nao, nao]."""
    ca, cb = p.mo_coeff
    na, nb = p.nocc_a, p.nocc_b
    ma = np.einsum("xpq,pi,qj->xij", ints_ao, ca[:, :na], ca[:, na:]).reshape(len(ints_ao), -1)
    mb = np.einsum("xpq,pi,qj->xij", ints_ao, cb[:, :nb], cb[:, nb:]).reshape(len(ints_ao), -1)
    n_a = ma.shape[1]
    return x_pyscf[:, :n_a] @ ma.T + x_pyscf[:, n_a:] @ mb.T


def xtda_osc_str(p, e: np.ndarray, x_pyscf: np.ndarray, dip_ao: np.ndarray) -> np.ndarray:
    """Length-form oscillator strengths f = 2/3 w |<0|r|n>|^2 (XTDA.py:838-858)."""
    td = xtda_transition_moments(p, x_pyscf, dip_ao)
    return 2.0 / 3.0 * e * np.einsum("sx,sx->s", td, td)


def xtda_rot_str(p, e: np.ndarray, x_pyscf: np.ndarray, ipovlp_ao: np.ndarray, irxp_ao: np.ndarray) -> np.ndarray:
    """Rotatory strengths in cgs units (XTDA.py:860-890): -<nabla> . (1/2)<r x p> / w / cgs2au."""
    ele = -xtda_transition_moments(p, x_pyscf, ipovlp_ao)
    mag = 0.5 * xtda_transition_moments(p, x_pyscf, irxp_ao)
    return np.einsum("s,sx,sx->s", 1.0 / e, ele, mag) / CGS2AU


# ---------------------------------------------------------------------------------------------------------
# spin-flip-down states: block-order vectors  cv | co | ov | oo
# ---------------------------------------------------------------------------------------------------------
def split_blocks(v: np.ndarray, nc: int, no: int, nv: int, vects=None):
    d1, d2, d3 = nc * nv, nc * nv + nc * no, nc * nv + nc * no + no * nv
    oo = v[d3:]
    if vects is not None:
        oo = vects @ oo
    return v[:d1].reshape(nc, nv), v[d1:d2].reshape(nc, no), v[d2:d3].reshape(no, nv), oo.reshape(no, no)


def sa_factors(no: int, sa: int):
    """factor1..3 of calculate_TDM_R (XSF_TDA.py:498-505): 1, 1, 0 without spin adaptation."""
    si = no / 2.0
    if sa == 0:
        return 1.0, 1.0, 0.0
    return math.sqrt((2 * si + 1) / (2 * si)), math.sqrt((2 * si) / (2 * si - 1)), 1.0 / math.sqrt(2 * si * (2 * si - 1))


def tdm_r(v: np.ndarray, ints_mo: np.ndarray, nc: int, no: int, nv: int, sa: int, vects=None) -> np.ndarray:
    """tdm[x, i, j] between spin-adapted spin-flip states i and j on a ROKS reference; the 16 block couplings of
    XSF_TDA.py:526-590 written as traces.  v: [dim, nstates] block order (reduced OO block if vects is given)."""
    f1, f2, f3 = sa_factors(no, sa)
    C, O, V = slice(0, nc), slice(nc, nc + no), slice(nc + no, None)
    ns = v.shape[1]
    blocks = [split_blocks(v[:, k], nc, no, nv, vects) for k in range(ns)]
    out = np.zeros((len(ints_mo), ns, ns))
    for x, d in enumerate(ints_mo):
        for i, (cv0, co0, ov0, oo0) in enumerate(blocks):
            for j, (cv1, co1, ov1, oo1) in enumerate(blocks):
                t = np.sum(cv0 * (cv1 @ d[V, V].T)) - np.sum(cv0 * (d[C, C] @ cv1))                       # CV-CV
                t += f1 * (np.sum(cv0 * (co1 @ d[V, O].T)) + np.sum(co0 * (cv1 @ d[V, O])))               # CV-CO
                t -= f1 * (np.sum(cv0 * (d[C, O] @ ov1)) + np.sum(ov0 * (d[C, O].T @ cv1)))               # CV-OV
                t += np.sum(co0 * (co1 @ d[O, O].T)) - np.sum(co0 * (d[C, C] @ co1))                       # CO-CO
                t -= f2 * (np.sum(co0 * (d[C, O] @ oo1)) + np.sum(oo0 * (d[C, O].T @ co1)))               # CO-OO
                t += f3 * (np.sum(co0 * d[C, O]) * np.trace(oo1) + np.trace(oo0) * np.sum(d[C, O] * co1))
                t += np.sum(ov0 * (ov1 @ d[V, V].T)) - np.sum(ov0 * (d[O, O] @ ov1))                       # OV-OV
                t += f2 * (np.sum(ov0 * (oo1 @ d[V, O].T)) + np.sum(oo0 * (ov1 @ d[V, O])))               # OV-OO
                t -= f3 * (np.sum(ov0 * d[O, V]) * np.trace(oo1) + np.trace(oo0) * np.sum(d[O, V] * ov1))
                t += np.sum(oo0 * (oo1 @ d[O, O].T)) - np.sum(oo0 * (d[O, O] @ oo1))                       # OO-OO
                out[x, i, j] = t
    return out


def tdm_u(v: np.ndarray, ints_aa: np.ndarray, ints_bb: np.ndarray, nc: int, no: int, nv: int) -> np.ndarray:
    """UKS reference (XSF_TDA.py:451-476): tdm = tr(c0 D^bb_vir c1^T) - tr(c0^T D^aa_occ c1), c = [[co, cv], [oo, ov]]."""
    ns = v.shape[1]
    cs = []
    for k in range(ns):
        cv, co, ov, oo = split_blocks(v[:, k], nc, no, nv)
        cs.append(np.block([[co, cv], [oo, ov]]))
    na = nc + no
    out = np.zeros((len(ints_aa), ns, ns))
    for x in range(len(ints_aa)):
        for i, c0 in enumerate(cs):
            for j, c1 in enumerate(cs):
                out[x, i, j] = np.trace(c0 @ ints_bb[x][nc:, nc:] @ c1.T) - np.trace(c0.T @ ints_aa[x][:na, :na] @ c1)
    return out


def osc_matrix(e: np.ndarray, tdm: np.ndarray) -> np.ndarray:
    """osc[i,j] = 2/3 |e_i - e_j| |tdm_ij|^2 (XSF_TDA_GPU.py:989,1114)."""
    return 2.0 / 3.0 * np.abs(e[:, None] - e[None, :]) * np.einsum("xij,xij->ij", tdm, tdm)


def delta_s2_u(p, v: np.ndarray, ovlp: np.ndarray) -> np.ndarray:
    """D<S^2> of spin-flip-down states on a UKS reference (XSF_TDA.py:613-649, 781-784): P_ab - no + 1."""
    ca, cb = p.mo_coeff
    na, nb = p.nocc_a, p.nocc_b
    nc, no, nv = p.nc, p.no, p.nv
    sba_oo = cb[:, :nb].T @ ovlp @ ca[:, :na]             # [nocc_b, nocc_a]
    sba_vo = cb[:, nb:].T @ ovlp @ ca[:, :na]             # [nvir_b, nocc_a]
    out = np.zeros(v.shape[1])
    for k in range(v.shape[1]):
        cv, co, ov, oo = split_blocks(v[:, k], nc, no, nv)
        x = np.block([[co, cv], [oo, ov]]).T              # [nvir_b, nocc_a]
        pab = (np.einsum("ai,aj,jk,ki", x, x, sba_oo.T, sba_oo) - np.einsum("ai,bi,kb,ak", x, x, sba_vo.T, sba_vo)
               + np.einsum("ai,bj,jb,ai", x, x, sba_vo.T, sba_vo))
        out[k] = pab - no + 1
    return out


def delta_s2_roks_sf(v: np.ndarray, nc: int, no: int, nv: int, vects=None) -> np.ndarray:
    """ROKS reference without spin adaptation (XSF_TDA.py:771-779): -2S + 1 + |cv|^2 - |oo|^2 + (tr oo)^2."""
    out = np.zeros(v.shape[1])
    for k in range(v.shape[1]):
        cv, _, _, oo = split_blocks(v[:, k], nc, no, nv, vects)
        out[k] = -no + 1 + np.sum(cv * cv) - np.sum(oo * oo) + np.trace(oo) ** 2
    return out
