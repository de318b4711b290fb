"""Test helper: the reference's Green's-function flow (ED_GF_NORMAL.f90:38-106,123-306,531-903) written
once over two interchangeable back-ends -- the product (CUDA through the C ABI) and the CPU oracle.
For a non-degenerate ground state in the sector (Ns/2, Ns/2), spin up, T=0 (zeta_function = 1)."""
import numpy as np

from cdmft_lanc_ed_b200 import models

LANC_NGFITER = 200  # ED_INPUT_VARS.f90:177


class _Product:
    def __init__(self, mdl, ed):
        self.E = ed
        ed.ed_set_model(mdl)

    def gs(self, isec):
        n = self.E.build_Hv_sector(isec)
        vec = np.zeros(n, dtype=np.complex128)
        e0, _, _, _ = self.E.sp_lanc_eigh(vec, 512, 1e-14)
        self.E.delete_Hv_sector()
        return e0, vec

    def apply(self, isec, iop, pos, coef, state):
        return self.E.apply_op(isec, iop, 1, pos, coef, state)

    def tridiag(self, jsec, v, nlanc):
        self.E.build_Hv_sector(jsec)
        _, a, b = self.E.sp_lanc_tridiag(v, nlanc)
        self.E.delete_Hv_sector()
        return a, b

    def add(self, w, e0, a, b, isign, wm, g):
        self.E.add_to_lanczos_gf_normal(w, e0, a, b, isign, 1.0, wm, g)


class _Oracle:
    def __init__(self, mdl, edo):
        self.edo = edo
        self.o = edo.Oracle(mdl)
        self.ns = mdl.ns

    def gs(self, isec):
        self.o.build_hv_sector(isec, self.edo.SPARSE_SERIAL)
        e0, vec, _, _, _ = self.o.lanc_eigh(512, 1e-14)
        self.o.delete_hv_sector()
        return e0, vec

    def apply(self, isec, iop, pos, coef, state):
        return self.edo.apply_op(self.ns, isec, iop, 1, pos, coef, state)

    def tridiag(self, jsec, v, nlanc):
        self.o.build_hv_sector(jsec, self.edo.SPARSE_SERIAL)
        _, a, b = self.o.lanc_tridiag(v, nlanc)
        self.o.delete_hv_sector()
        return a, b

    def add(self, w, e0, a, b, isign, wm, g):
        self.edo.add_to_lanczos_gf(w, e0, a, b, isign, 1.0, wm, g)


def _channel(be, isec, e0, gsvec, iop, pos, coef, weight_factor, isign, wm, g, getdim):
    jsec, vv = be.apply(isec, iop, pos, coef, gsvec)
    if jsec == 0:
        return
    norm2 = np.vdot(vv, vv).real
    vv = vv / np.sqrt(norm2)
    nlanc = min(getdim(jsec), LANC_NGFITER)
    a, b = be.tridiag(jsec, vv, nlanc)
    be.add(weight_factor * norm2, e0, a, b, isign, wm, g)


def gimp_element(backend, mdl, ia, ib, wm, ed=None, edo=None, gs=None):
    """G_{ia,ib}(i wm), ia/ib 1-based impurity orbital positions (imp_state_index)."""
    be = _Product(mdl, ed) if backend == "product" else _Oracle(mdl, edo)
    ns = mdl.ns
    isec = models.get_sector(ns, ns // 2, ns // 2)

    def getdim(jsec):
        from math import comb
        nup, ndw = (jsec - 1) // (ns + 1), (jsec - 1) % (ns + 1)
        return comb(ns, nup) * comb(ns, ndw)

    e0, vec = be.gs(isec) if gs is None else gs

    def diag(p):
        g = np.zeros(len(wm), dtype=np.complex128)
        _channel(be, isec, e0, vec, +1, [p], [1.0], 1.0, +1, wm, g, getdim)
        _channel(be, isec, e0, vec, -1, [p], [1.0], 1.0, -1, wm, g, getdim)
        return g

    if ia == ib:
        return diag(ia)
    g = np.zeros(len(wm), dtype=np.complex128)
    _channel(be, isec, e0, vec, +1, [ia, ib], [1.0, 1.0], 1.0, +1, wm, g, getdim)
    _channel(be, isec, e0, vec, -1, [ia, ib], [1.0, 1.0], 1.0, -1, wm, g, getdim)
    _channel(be, isec, e0, vec, +1, [ia, ib], [1.0, 1.0j], -1.0j, +1, wm, g, getdim)
    _channel(be, isec, e0, vec, -1, [ia, ib], [1.0, -1.0j], -1.0j, -1, wm, g, getdim)
    # non-diagonal trick, ED_GF_NORMAL.f90:91-103 (chan4 = 1)
    return 0.5 * (g - (1 - 1j) * diag(ia) - (1 - 1j) * diag(ib))
