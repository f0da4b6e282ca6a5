"""Trial-vector layouts and amplitude post-processing (integer permutations + tiny fp64 maps).

Follows xtddft/utils/utils.py:44-122 (`order_pyscf2my`, `so2st`, `st2so`), xtddft/SF_TDA.py:304-345 and
xtddft/XSF_TDA.py:1419-1453 (`deal_v_davidson`), xtddft/XSF_TDA.py:397-427,999-1027,1279-1290 (`get_vect`,
`remove`, OO expand / compress) and xtddft/XSF_TDA_GPU.py:426-439,485-500,705-720 (PySCF-order removed layout).
SURVEY Appendix A.5 describes the layouts.
"""
import numpy as np


def order_pyscf2my(nc, no, nv):
    """X-TDA: PySCF order [alpha (c,o)x v | beta c x (o,v)] -> [CV(aa) | OV(aa) | CO(bb) | CV(bb)]."""
    na = (nc + no) * nv
    beta = na + np.arange(nc * (no + nv)).reshape(nc, no + nv)
    return np.concatenate([np.arange(na), beta[:, :no].ravel(), beta[:, no:].ravel()])


def so2st(v, nc, no, nv):
    """spin-orbital [cva|ova|cob|cvb] rows -> spin-tensor [CV(0)|OV(0)|CO(0)|CV(1)] (utils.py:67-94)."""
    d1, d2, d3 = nc * nv, (nc + no) * nv, (nc + no) * nv + nc * no
    cva, ova, cob, cvb = v[:d1], v[d1:d2], v[d2:d3], v[d3:]
    r = np.sqrt(2.0) / 2.0
    return np.concatenate([r * (cva + cvb), ova, cob, r * (cva - cvb)], axis=0)


def st2so(v, nc, no, nv):
    d1, d2, d3 = nc * nv, (nc + no) * nv, (nc + no) * nv + nc * no
    cv0, ov0, co0, cv1 = v[:d1], v[d1:d2], v[d2:d3], v[d3:]
    return np.concatenate([(cv0 + cv1) / np.sqrt(2.0), ov0, co0, (cv0 - cv1) / np.sqrt(2.0)], axis=0)


def get_vect(no):
    """Orthonormal basis [no*no, no*no-1] of the OO space orthogonal to sum_u |u->u>/sqrt(no)
    (XSF_TDA.py:397-414): off-diagonal unit vectors plus no-1 traceless diagonal combinations."""
    vect = np.zeros((no, no - 1))
    for i in range(1, no):
        fac = 1.0 / np.sqrt((no - i + 1) * (no - i))
        vect[i - 1:, i - 1] = np.array([no - i] + [-1] * (no - i)) * fac
    vects = np.eye(no * no)[:, :-1].copy()
    diag = [i * (no + 1) for i in range(no)]
    for i in range(no - 1):
        vects[0::no + 1, diag[i]] = vect[:, i]
    return vects


def block_dims(nc, no, nv):
    d1 = nc * nv
    d2 = d1 + nc * no
    d3 = d2 + no * nv
    return d1, d2, d3


def split_blocks(data, nc, no, nv, vects=None):
    """block-order vectors [x, dim] -> (cv, co, ov, oo); `vects` expands a removed OO block (XSF_TDA.py:1011-1027)."""
    data = np.atleast_2d(np.asarray(data))
    d1, d2, d3 = block_dims(nc, no, nv)
    x = data.shape[0]
    oo = data[:, d3:]
    if vects is not None:
        oo = oo @ vects.T
    return (data[:, :d1].reshape(x, nc, nv), data[:, d1:d2].reshape(x, nc, no),
            data[:, d2:d3].reshape(x, no, nv), oo.reshape(x, no, no))


def join_blocks(cv, co, ov, oo, vects=None):
    """(cv, co, ov, oo) -> block-order [x, dim]; `vects` compresses OO (XSF_TDA.py:1279-1290)."""
    x = cv.shape[0]
    oo = oo.reshape(x, -1)
    if vects is not None:
        oo = oo @ vects
    return np.hstack([cv.reshape(x, -1), co.reshape(x, -1), ov.reshape(x, -1), oo])


def pyscf_to_block_index(nc, no, nv):
    """Permutation taking the spin-flip-down PySCF order ((c,o) x (o,v) row-major) to [cv|co|ov|oo]
    (`deal_v_davidson` without remove, SF_TDA.py:327-345)."""
    idx = np.arange((nc + no) * (no + nv)).reshape(nc + no, no + nv)
    return np.concatenate([idx[:nc, no:].ravel(), idx[:nc, :no].ravel(), idx[nc:, no:].ravel(), idx[nc:, :no].ravel()])


def deal_v_davidson(v, nc, no, nv, removed=False):
    """Columns of v in PySCF order (optionally with the last OO element dropped) -> block order.
    With removed=True the OO part stays the (no*no-1)-vector of reduced coordinates (XSF_TDA.py:1441-1446)."""
    v = np.asarray(v)
    if not removed:
        return v[pyscf_to_block_index(nc, no, nv)]
    full_idx = np.arange((nc + no) * (no + nv)).reshape(nc + no, no + nv)
    oo_pos = full_idx[nc:, :no].ravel()[:-1]                       # positions of the kept OO entries (full indexing)
    last = full_idx[nc + no - 1, no - 1]
    shift = lambda a: a - (a > last)                               # index in the shortened vector
    cv, co, ov = full_idx[:nc, no:].ravel(), full_idx[:nc, :no].ravel(), full_idx[nc:, no:].ravel()
    return v[np.concatenate([shift(cv), shift(co), shift(ov), shift(oo_pos)])]


def gpu_order_expand(zs, nc, no, nv, vects):
    """[x, dim-1] PySCF-order removed layout -> [x, dim] full PySCF order (XSF_TDA_GPU.py:485-500)."""
    zs = np.asarray(zs)
    nvir = no + nv
    full_idx = np.arange((nc + no) * nvir).reshape(nc + no, nvir)
    oo_pos = full_idx[nc:, :no].ravel()
    last = oo_pos[-1]
    other = np.setdiff1d(np.arange((nc + no) * nvir), oo_pos)
    out = np.zeros((zs.shape[0], (nc + no) * nvir))
    out[:, other] = zs[:, other - (other > last)]
    out[:, oo_pos] = zs[:, oo_pos[:-1]] @ vects.T
    return out


def gpu_order_compress(hx, nc, no, nv, vects):
    """[x, dim] full PySCF order -> [x, dim-1] removed layout (XSF_TDA_GPU.py:705-720)."""
    hx = np.asarray(hx)
    nvir = no + nv
    full_idx = np.arange((nc + no) * nvir).reshape(nc + no, nvir)
    oo_pos = full_idx[nc:, :no].ravel()
    last = oo_pos[-1]
    other = np.setdiff1d(np.arange((nc + no) * nvir), oo_pos)
    out = np.zeros((hx.shape[0], (nc + no) * nvir - 1))
    out[:, other - (other > last)] = hx[:, other]
    out[:, oo_pos[:-1]] = hx[:, oo_pos] @ vects
    return out


def delta_s2_xtda(v_my, nc, no, nv):
    """X-TDA: dS2 = |X_cv(aa) - X_cv(bb)|^2 per state, amplitudes in [cva|ova|cob|cvb] order (XTDA.py:831-836)."""
    d3 = (nc + no) * nv + nc * no
    cva, cvb = v_my[:nc * nv], v_my[d3:]
    return np.einsum("ik,ik->k", cva - cvb, cva - cvb)


def delta_s2_sf(v_block, nc, no, nv, vects=None):
    """spin-flip-down, ROKS reference: dS2 = -no + 1 + |cv|^2 - |oo|^2 + (tr oo)^2 (SF_TDA.py:819-825)."""
    out = []
    for k in range(v_block.shape[1]):
        cv, co, ov, oo = split_blocks(v_block[:, k][None], nc, no, nv, vects)
        out.append(-no + 1 + np.sum(cv * cv) - np.sum(oo * oo) + np.trace(oo[0]) ** 2)
    return np.array(out)
