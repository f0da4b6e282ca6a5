"""Grid (exchange-correlation kernel) contractions, restating the PySCF numint routines the
reference calls or forks.  Dense AO values (no screening): the sparse helpers `_dot_ao_ao_sparse`,
`_scale_ao_sparse` (pyscf/dft/numint.py) compute the same sums.

  * eval_rho1        -- `ni._gen_rho_evaluator(mol, dms, hermi=0)` / `eval_rho`
  * nr_uks_fxc       -- pyscf.dft.numint.nr_uks_fxc, called at xtddft/XTDA.py:514
  * nr_uks_fxc_sf    -- xtddft/SF_TDA.py:90-160 (ALDA0 spin-flip kernel, row a8)
  * nr_uks_fxc_sf_mc -- xtddft/SF_TDA.py:976-1047 (multicollinear kernel, row a9)
"""
import numpy as np


def eval_rho1(ao, dm, mgga=False):
    """rho[c,g] of a (non-symmetric) density matrix, hermi=0.
    LDA (ao[1,g,n]): rho_0 = sum phi_mu D_mu,nu phi_nu.
    GGA (ao[4,g,n]): rho_k = sum (d_k phi_mu) D phi_nu + phi_mu D (d_k phi_nu).
    meta-GGA (ao[4,g,n], MGGA_DENSITY_LAPL off): a fifth component tau = 1/2 sum_k (d_k phi_mu) D (d_k phi_nu)."""
    nvar = ao.shape[0]
    c0 = ao[0] @ dm                      # c0[g,nu] = sum_mu phi_mu D_mu,nu
    rho = np.empty((5 if mgga else nvar, ao.shape[1]))
    rho[0] = np.einsum("gn,gn->g", c0, ao[0])
    if nvar > 1:
        c1 = ao[0] @ dm.T
        for k in range(1, 4):
            rho[k] = np.einsum("gn,gn->g", c0, ao[k]) + np.einsum("gn,gn->g", c1, ao[k])
    if mgga:
        rho[4] = 0.5 * sum(np.einsum("gn,gn->g", ao[k] @ dm, ao[k]) for k in range(1, 4))
    return rho


def _integrate(ao, wv, lda):
    """V = ao^T diag(wv) ao (LDA)  or  sym( ao_0^T sum_c ao_c wv_c ) with wv_0 halved (GGA); a fifth row of wv is the tau
    potential: + 1/2 sum_k (d_k ao)^T wv_4 (d_k ao)  (SF_TDA.py:141-152: wv[4] *= .5, _tau_dot_sparse, added after the
    symmetrisation)."""
    if lda:
        return (ao[0] * wv[0][:, None]).T @ ao[0]
    w = wv[:4].copy()
    w[0] *= 0.5
    aow = np.einsum("cgn,cg->gn", ao, w)
    v = ao[0].T @ aow
    v = v + v.T
    if wv.shape[0] == 5:
        v = v + sum(ao[k].T @ (ao[k] * (0.5 * wv[4])[:, None]) for k in range(1, 4))
    return v


def nr_uks_fxc(ao, weights, fxc, dms):
    """dms[2,x,N,N] -> v[2,x,N,N]; fxc[2,nvar,2,nvar,ng] unweighted (numint.cache_xc_kernel layout)."""
    dms = np.asarray(dms)
    nvar = ao.shape[0]
    nset = dms.shape[1]
    out = np.zeros_like(dms)
    for i in range(nset):
        mg = fxc.shape[1] == 5
        rho1 = np.stack([eval_rho1(ao, dms[0, i], mg), eval_rho1(ao, dms[1, i], mg)])      # [2,nvar,g]
        wv = np.einsum("axg,axbyg->byg", rho1, fxc) * weights
        for s in range(2):
            out[s, i] = _integrate(ao, wv[s], lda=(nvar == 1))
    return out


def nr_uks_fxc_sf(ao, fxc_w, dms):
    """ALDA0 spin-flip kernel: only the density component enters, fxc_w[g] already carries the weights
    (SF_TDA.py:82-84,108-117); GGA route halves wv_0 and symmetrises (SF_TDA.py:133-139)."""
    dms = np.asarray(dms)
    out = np.zeros_like(dms)
    nvar = ao.shape[0]
    for i in range(dms.shape[0]):
        rho1 = eval_rho1(ao, dms[i])
        wv = np.zeros_like(rho1)
        wv[0] = rho1[0] * fxc_w
        out[i] = _integrate(ao, wv, lda=(nvar == 1))
    return out


def nr_uks_fxc_sf_mc(ao, weights, fxc_sf, dms):
    """multicollinear spin-flip kernel fxc_sf[nvar,nvar,g]: wv_a = sum_b rho_b * 2 f_ba * w (SF_TDA.py:998-1003)."""
    dms = np.asarray(dms)
    out = np.zeros_like(dms)
    nvar = ao.shape[0]
    for i in range(dms.shape[0]):
        rho1 = eval_rho1(ao, dms[i], fxc_sf.shape[0] == 5)
        if nvar == 1:
            wv = (rho1[0] * fxc_sf[0, 0] * 2.0 * weights)[None]
        else:
            wv = np.einsum("bg,bag->ag", rho1, fxc_sf * 2.0) * weights
        out[i] = _integrate(ao, wv, lda=(nvar == 1))
    return out
