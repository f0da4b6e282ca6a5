"""NumPy interpreter of a sigma-build `Plan` (TEST INFRASTRUCTURE).

Executes the same generic steps as the CUDA engine -- MO-resident DF blocks, block-weighted exchange,
J blocks with mixing, half-transformed grid contraction, local GEMM / rank-1 / diagonal terms, sparse layout
maps -- with dense NumPy, so plan construction (xtddft_b200/plan.py) can be checked against the oracle on a
machine without a GPU.  Never imported by the product.
"""
import numpy as np

es = lambda *a: np.einsum(*a, optimize=True)


def _gather_cols(c, idx):
    out = np.zeros((c.shape[0], len(idx)))
    m = idx >= 0
    out[:, m] = c[:, idx[m]]
    return out


class PlanInterpreter:
    def __init__(self, plan, p):
        self.plan, self.p = plan, p
        self.co, self.cv = [], []
        for ch in plan.channels:
            self.co.append(_gather_cols(p.mo_coeff[ch.spin_o], ch.occ_idx))
            self.cv.append(_gather_cols(p.mo_coeff[ch.spin_v], ch.vir_idx))
        self.tensors = [p.cderi, p.cderi_lr]
        self.loo, self.lvv = {}, {}
        for kt in plan.k_terms:
            key = (kt.tensor, kt.ch)
            if key not in self.loo:
                L = self.tensors[kt.tensor]
                self.loo[key] = es("Pmn,mi,nj->Pij", L, self.co[kt.ch], self.co[kt.ch])
                self.lvv[key] = es("Pmn,ma,nb->Pab", L, self.cv[kt.ch], self.cv[kt.ch])
        self.lov = {}
        for kt in plan.kt_terms:
            key = (kt.tensor, kt.ch)
            if key not in self.lov:
                self.lov[key] = es("Pmn,mi,na->Pia", self.tensors[kt.tensor], self.co[kt.ch], self.cv[kt.ch])
        self.ljb = []
        for jb in plan.j_blocks:
            co = self.co[jb.ch][:, jb.r0:jb.r0 + jb.nr]
            cv = self.cv[jb.ch][:, jb.c0:jb.c0 + jb.nc]
            self.ljb.append(es("Pmn,mi,na->Pia", p.cderi, co, cv))
        if plan.xc_kind != "none":
            self.phi = [es("cgm,mo->cgo", p.ao, co) for co in self.co]

    # ---- layout -------------------------------------------------------------------------------
    def pack(self, z_ext):
        plan = self.plan
        x = z_ext.shape[0]
        zs = [np.zeros((x, ch.no, ch.nv)) for ch in plan.channels]
        ent, coef = plan.layout_entries, plan.layout_coefs
        for k in range(ent.shape[0]):
            e, c, i, a = ent[k]
            zs[c][:, i, a] += coef[k] * z_ext[:, e]
        return zs

    def unpack(self, sig):
        plan = self.plan
        x = sig[0].shape[0]
        out = np.zeros((x, plan.ext_dim))
        ent, coef = plan.layout_entries, plan.layout_coefs
        for k in range(ent.shape[0]):
            e, c, i, a = ent[k]
            out[:, e] += coef[k] * sig[c][:, i, a]
        return out

    def jblock_diag(self, jbi):
        return es("Pia,Pia->ia", self.ljb[jbi], self.ljb[jbi])

    # ---- sigma --------------------------------------------------------------------------------
    def sigma(self, z_ext):
        z_ext = np.atleast_2d(np.asarray(z_ext, dtype=float))
        zs = self.pack(z_ext)
        sig = self.partial_blocks(zs)
        self.add_local_blocks(zs, sig)
        return self.unpack(sig)

    def partial_blocks(self, zs):
        """The part that is linear in this interpreter's aux functions and grid points (what `xtd_sigma_partial`
        leaves in the engine's buffer; summed over ranks by the all-reduce)."""
        plan, p = self.plan, self.p
        sig = [np.zeros_like(z) for z in zs]
        # exchange with block weights
        for kt in plan.k_terms:
            ch = plan.channels[kt.ch]
            loo, lvv = self.loo[(kt.tensor, kt.ch)], self.lvv[(kt.tensor, kt.ch)]
            z = zs[kt.ch]
            for ib, (i0, ni) in enumerate(ch.o_blocks):
                for ab, (a0, na) in enumerate(ch.v_blocks):
                    zt = np.zeros_like(z)
                    for jb, (j0, nj) in enumerate(ch.o_blocks):
                        for bb, (b0, nbb) in enumerate(ch.v_blocks):
                            zt[:, j0:j0 + nj, b0:b0 + nbb] = kt.weights[ib, ab, jb, bb] * z[:, j0:j0 + nj, b0:b0 + nbb]
                    u = es("Pij,xjb->Pxib", loo[:, i0:i0 + ni, :], zt)
                    sig[kt.ch][:, i0:i0 + ni, a0:a0 + na] += es("Pxib,Pba->xia", u, lvv[:, :, a0:a0 + na])
        # exchange of the transposed trial density (Z-vector plans)
        for kt in plan.kt_terms:
            lov = self.lov[(kt.tensor, kt.ch)]
            w = es("Pib,xjb->Pxij", lov, zs[kt.ch])
            sig[kt.ch] += kt.weight * es("Pxij,Pja->xia", w, lov)
        # Coulomb blocks
        if plan.j_blocks:
            rho = [es("Pia,xia->xP", self.ljb[k], zs[jb.ch][:, jb.r0:jb.r0 + jb.nr, jb.c0:jb.c0 + jb.nc])
                   for k, jb in enumerate(plan.j_blocks)]
            for t, jb in enumerate(plan.j_blocks):
                mixed = sum(plan.j_mix[t, s] * rho[s] for s in range(len(rho)))
                sig[jb.ch][:, jb.r0:jb.r0 + jb.nr, jb.c0:jb.c0 + jb.nc] += es("xP,Pia->xia", mixed, self.ljb[t])
        # grid kernel, half-transformed: Y = ao . (Cv z^T), rho via phi_o, A-buffers, R = ao^T A, project with Cv
        if plan.xc_kind != "none":
            nvar = p.ao.shape[0]
            tau = plan.xc_kind.endswith("_tau")                        # meta-GGA: fifth component of rho / wv
            kind = plan.xc_kind.replace("_tau", "")
            ys = [es("cgm,mv,xov->cgxo", p.ao, self.cv[c], zs[c]) for c in range(len(zs))]
            rhos = []
            for c in range(len(zs)):
                y, ph = ys[c], self.phi[c]
                r = np.zeros((nvar + int(tau),) + y.shape[1:3])        # [c, g, x]
                r[0] = es("gxo,go->gx", y[0], ph[0])
                for k in range(1, nvar):
                    r[k] = es("gxo,go->gx", y[k], ph[0]) + es("gxo,go->gx", y[0], ph[k])
                    if tau:
                        r[4] += 0.5 * es("gxo,go->gx", y[k], ph[k])
                rhos.append(r)
            wvs = []
            if kind == "uks":
                rho1 = np.stack(rhos)                                  # [s, c, g, x]
                wv = es("scgx,sctdg->tdgx", rho1, p.fxc_uks) * p.weights[None, None, :, None]
                wvs = [wv[0], wv[1]]
            elif kind == "alda0":
                wv = np.zeros_like(rhos[0])
                wv[0] = rhos[0][0] * p.fxc_alda0[:, None]
                wvs = [wv]
            elif kind == "mcol":
                wvs = [es("bgx,bag->agx", rhos[0], 2.0 * p.fxc_mcol) * p.weights[None, :, None]]
            for c in range(len(zs)):
                wv, ph = wvs[c], self.phi[c]
                a = np.zeros_like(ys[c])
                a[0] = es("gx,go->gxo", wv[0], ph[0])
                for k in range(1, nvar):
                    a[0] += es("gx,go->gxo", wv[k], ph[k])
                    a[k] = es("gx,go->gxo", wv[k], ph[0])
                    if tau:
                        a[k] += 0.5 * es("gx,go->gxo", wv[4], ph[k])
                rt = es("cgxo,cgm->xom", a, p.ao)
                sig[c] += plan.xc_scale * es("xom,mv->xov", rt, self.cv[c])
        return sig

    def add_local_blocks(self, zs, sig):
        """Replicated Fock / Delta-A local terms (what `xtd_sigma_finish` adds after the reduction)."""
        plan = self.plan
        for lg in plan.local_gemms:
            dc, r0, nr, c0, ncol = lg.dst
            sc, sr0, sc0 = lg.src
            if lg.side == "R":
                k = lg.mat.shape[0]
                sig[dc][:, r0:r0 + nr, c0:c0 + ncol] += lg.alpha * es("xrb,bc->xrc", zs[sc][:, sr0:sr0 + nr, sc0:sc0 + k], lg.mat)
            elif lg.side == "L":
                k = lg.mat.shape[1]
                sig[dc][:, r0:r0 + nr, c0:c0 + ncol] += lg.alpha * es("rj,xjc->xrc", lg.mat, zs[sc][:, sr0:sr0 + k, sc0:sc0 + ncol])
            elif lg.side == "LT":
                k = lg.mat.shape[1]
                sig[dc][:, r0:r0 + nr, c0:c0 + ncol] += lg.alpha * es("rk,xck->xrc", lg.mat, zs[sc][:, sr0:sr0 + ncol, sc0:sc0 + k])
            else:
                assert lg.side == "RT"
                k = lg.mat.shape[0]
                sig[dc][:, r0:r0 + nr, c0:c0 + ncol] += lg.alpha * es("xjr,jc->xrc", zs[sc][:, sr0:sr0 + k, sc0:sc0 + nr], lg.mat)
        for r1 in plan.rank1s:
            sig[r1.dst_ch] += es("x,ia->xia", es("ia,xia->x", r1.v, zs[r1.src_ch]), r1.u)
        for d in plan.diags:
            sig[d.ch] += d.d[None] * zs[d.ch]
        return sig
