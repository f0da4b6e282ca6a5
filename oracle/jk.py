"""Density-fitted Coulomb / exchange, restating the PySCF `get_jk` convention the reference calls
(xtddft/XTDA.py:520-539, xtddft/SF_TDA.py:273-276, xtddft/XSF_TDA.py:857,996):

    J_kl = sum_ij (ij|kl) D_ji ,   K_il = sum_jk (ij|kl) D_jk ,   (ij|kl) ~= sum_P L_P,ij L_P,kl

with hermi=0 (no symmetry assumed for D) and batching over leading dimensions of `dms`.
PySCF (pyscf/df/df_jk.py, 2.11/2.12, not vendored) streams aux blocks and forms K as (L_P D) L_P.
"""
import numpy as np


def get_j(cderi, dms):
    dms = np.asarray(dms)
    shape = dms.shape
    d = dms.reshape(-1, shape[-2], shape[-1])
    rho = np.einsum("Pij,xji->xP", cderi, d, optimize=True)
    vj = np.einsum("xP,Pkl->xkl", rho, cderi, optimize=True)
    return vj.reshape(shape)


def get_k(cderi, dms):
    dms = np.asarray(dms)
    shape = dms.shape
    d = dms.reshape(-1, shape[-2], shape[-1])
    vk = np.zeros_like(d)
    for x in range(d.shape[0]):
        tmp = np.matmul(cderi, d[x])                 # (L_P D)[i,k]
        vk[x] = np.einsum("Pik,Pkl->il", tmp, cderi, optimize=True)
    return vk.reshape(shape)


def get_jk(cderi, dms, with_j=True, with_k=True):
    vj = get_j(cderi, dms) if with_j else None
    vk = get_k(cderi, dms) if with_k else None
    return vj, vk


def mo_eri(cderi, c1, c2, c3, c4):
    """(pq|rs) in the MO basis from the DF tensor; plays the role of `ao2mo.general` in the
    reference's explicit-matrix builders (XTDA.py:120-121, SF_TDA.py:646-649, XSF_TDA.py:338-339)."""
    l12 = np.einsum("Pij,ip,jq->Ppq", cderi, c1, c2, optimize=True)
    l34 = np.einsum("Pij,ip,jq->Ppq", cderi, c3, c4, optimize=True)
    return np.einsum("Ppq,Prs->pqrs", l12, l34, optimize=True)


def rohf_fock_difference(cderi, mo_coeff, open_idx):
    """F_beta - F_alpha of the ROHF-form Fock matrices in the MO basis.  The reference builds both with
    `scf.ROHF(mol).get_veff(mol, dm)` (xtddft/XTDA.py:607-613, xtddft/XSF_TDA.py:1103-1111) and uses only their difference:
    h and J[D_alpha + D_beta] cancel, -K[D_beta] + K[D_alpha] = K[D_open], K[D]_pq = sum_P sum_u L^P_pu L^P_qu."""
    c = np.asarray(mo_coeff)
    d_open = c[:, open_idx] @ c[:, open_idx].T
    return c.T @ get_k(cderi, d_open) @ c
