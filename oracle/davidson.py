"""Block Davidson eigensolver, restating xtddft/utils/Davidson.py:21-298 (`davidson1`, a fork of
`pyscf.lib.linalg_helper.davidson1`) together with the upstream helpers it imports at
Davidson.py:7-9 (`_qr`, `_fill_heff_hermitian`, `_sort_elast`, `_outprod_to_subspace`, `_normalize_xt_`,
`make_diag_precond`; PySCF 2.11/2.12 lib/linalg_helper.py, not vendored -- semantics in SURVEY Appendix B).

Deviations from the shipped file: no CuPy `.get()` hops, and the cycle/sigma
counts are returned as the 4th value every caller unpacks (`Davidcyc`, XTDA.py:775, XTDA_GPU.py:393).
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
import numpy as np
import scipy.linalg


class LinearDependencyError(RuntimeError):
    pass


def make_diag_precond(diag, level_shift=1e-3):
    def precond(dx, e, *args):
        diagd = diag - (e - level_shift)
        diagd[abs(diagd) < 1e-8] = 1e-8
        return dx / diagd
    return precond


def qr(xs, lindep=1e-14):
    """Modified Gram-Schmidt; a vector is kept when its remaining squared norm exceeds `lindep`."""
    qs = []
    for x in xs:
        xi = np.array(x, dtype=float, copy=True)
        for q in qs:
            xi -= q * np.dot(q, xi)
        nrm2 = np.dot(xi, xi)
        if nrm2 > lindep:
            qs.append(xi / np.sqrt(nrm2))
    return qs


def fill_heff(heff, xs, ax, xt, axt):
    nrow = len(axt)
    row1 = len(ax)
    row0 = row1 - nrow
    for ip, i in enumerate(range(row0, row1)):
        for jp, j in enumerate(range(row0, i)):
            heff[i, j] = heff[j, i] = np.dot(xt[ip], axt[jp])
        heff[i, i] = np.dot(xt[ip], axt[ip])
    for i in range(row0):
        for jp, j in enumerate(range(row0, row1)):
            heff[j, i] = heff[i, j] = np.dot(xt[jp], ax[i])
    return heff


def sort_elast(elast, conv_last, vlast, v):
    head, nroots = vlast.shape
    ovlp = abs(np.dot(v[:head].conj().T, vlast))
    mapping = np.argmax(ovlp, axis=1)
    found = np.any(ovlp > .5, axis=1)
    conv = conv_last[mapping]
    e = elast[mapping]
    conv[~found] = False
    e[~found] = 0.
    return e, conv


def normalize_xt(xt, xs, threshold):
    norm_min = 1
    out = []
    for xi in xt:
        if xi is None:
            continue
        for xsi in xs:
            xi -= xsi * np.dot(xsi, xi)
        norm = np.dot(xi, xi) ** .5
        if norm ** 2 > threshold:
            xi *= 1 / norm
            norm_min = min(norm_min, norm)
            out.append(xi)
    return out, norm_min


def davidson1(aop, x0, precond, tol=1e-12, max_cycle=50, max_space=12, lindep=1e-14,
              nroots=1, pick=None, tol_residual=None, callback=None):
    toloose = np.sqrt(tol) if tol_residual is None else tol_residual
    if not callable(precond):
        precond = make_diag_precond(precond)
    if isinstance(x0, np.ndarray) and x0.ndim == 1:
        x0 = [x0]
    x0 = [np.asarray(x, dtype=float) for x in x0]
    max_space = max_space + (nroots - 1) * 4
    heff = None
    fresh_start = True
    e = None
    v = None
    conv = np.zeros(nroots, dtype=bool)
    nsigma = 0
    icyc = -1
    for icyc in range(max_cycle):
        if fresh_start:
            xs, ax = [], []
            space = 0
            xt = qr(x0, lindep)
            if len(xt) == 0:
                raise LinearDependencyError("Initial guess is empty or zero" if icyc == 0 else
                                            "No more linearly independent basis were found.")
            x0 = None
            max_dx_last = 1e9
        elif len(xt) > 1:
            xt = qr(xt, lindep)
            xt = xt[:40]
        axt = [np.asarray(a) for a in aop(np.asarray(xt))]
        nsigma += len(xt)
        for k in range(len(xt)):
            xs.append(xt[k])
            ax.append(axt[k])
        rnow = len(xt)
        head, space = space, space + rnow
        if rnow == 0:
            raise LinearDependencyError("No linearly independent basis found by the diagonalization solver.")
        if heff is None:
            heff = np.empty((max_space + nroots, max_space + nroots))
        elast, vlast, conv_last = e, v, conv
        fill_heff(heff, xs, ax, xt, axt)
        xt = axt = None
        w, v = scipy.linalg.eigh(heff[:space, :space])
        if callable(pick):
            w, v, idx = pick(w, v, nroots, locals())
            if len(w) == 0:
                raise RuntimeError(f"Not enough eigenvalues found by {pick}")
        e = w[:nroots]
        v = v[:, :nroots]
        conv = np.zeros(e.size, dtype=bool)
        if not fresh_start:
            elast, conv_last = sort_elast(elast, conv_last, vlast, v)
        if elast is None or elast.size != e.size:
            de = e
        else:
            de = e - elast
        x0 = list(np.dot(v.T, np.asarray(xs)))
        ax0 = np.dot(v.T, np.asarray(ax))
        dx_norm = np.zeros(e.size)
        xt = [None] * e.size
        for k, ek in enumerate(e):
            xt[k] = ax0[k] - ek * x0[k]
            dx_norm[k] = np.sqrt(np.dot(xt[k], xt[k]))
            conv[k] = abs(de[k]) < tol and dx_norm[k] < toloose
        ax0 = None
        max_dx_norm = max(dx_norm)
        if all(conv):
            break
        for k, ek in enumerate(e):
            if (not conv[k]) and dx_norm[k] ** 2 > lindep:
                xt[k] = precond(xt[k], e[0], x0[k])
                xt[k] *= np.dot(xt[k], xt[k]) ** -.5
            else:
                xt[k] = None
        xt, norm_min = normalize_xt(xt, xs, lindep)
        if len(xt) == 0:
            conv = dx_norm < toloose
            break
        max_dx_last = max_dx_norm
        fresh_start = space + nroots > max_space
        if callable(callback):
            callback(locals())
    return np.asarray(conv), e, x0, [icyc + 1, nsigma]
