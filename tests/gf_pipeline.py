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


# ---- the reference's own non-interacting Green's function (ED_BATH_FUNCTIONS.f90) ---------------------------------------
def _lso(m6):
    """nnn2lso_reshape (ED_AUX_FUNX.f90): [Nlat,Nlat,Nspin,Nspin,Norb,Norb] -> [Nlso,Nlso] with the index
    iorb + (ilat-1)*Norb + (ispin-1)*Norb*Nlat of index_stride_lso (ED_AUX_FUNX.f90:81-87)."""
    L, _, S, _, O, _ = m6.shape
    n = L * S * O
    out = np.zeros((n, n), dtype=np.complex128)
    for il in range(L):
        for jl in range(L):
            for isp in range(S):
                for jsp in range(S):
                    for io in range(O):
                        for jo in range(O):
                            out[io + il * O + isp * O * L, jo + jl * O + jsp * O * L] = m6[il, jl, isp, jsp, io, jo]
    return out


def g0and_bath(mdl, x):
    """g0and_bath(x) of the reference, restated from ED_BATH_FUNCTIONS.f90:39-155:
        Delta(x) = sum_ib diag(v_ib) (x 1 - Hbath_ib)^-1 diag(v_ib)            delta_bath_array   :69-99
        G0^-1(x) = (x + xmu) 1 - impHloc - Delta(x)                             invg0_bath_array   :140-155
        G0(x)    = inv(G0^-1(x))                                                 g0and_bath         :102-121
    in the lso basis.  Returns [Nlso, Nlso, len(x)]."""
    hloc = _lso(mdl.imphloc)
    n = hloc.shape[0]
    out = np.zeros((n, n, len(x)), dtype=np.complex128)
    for i, z in enumerate(x):
        delta = np.zeros((n, n), dtype=np.complex128)
        for ib in range(mdl.nbath):
            vk = np.diag(mdl.vbath[:, ib]).astype(np.complex128)
            delta += vk @ np.linalg.inv(z * np.eye(n) - _lso(mdl.hbath[..., ib])) @ vk
        out[:, :, i] = np.linalg.inv((z + mdl.xmu) * np.eye(n) - hloc - delta)
    return out


def noninteracting_half_filled(mdl, grid=None):
    """U = 0 copy of a model with xmu chosen so that the ground state is the non-degenerate Slater determinant of the
    sector (Ns/2, Ns/2): exactly Ns/2 single-particle levels of [[impHloc - xmu, V], [V, Hbath]] below zero.  Returns
    (model, gap) or (None, 0) when no xmu of the grid does it.  Nspin = 1."""
    import copy
    assert mdl.nspin == 1
    n = mdl.nimp
    ns = mdl.ns
    best, best_gap = None, 0.0
    for xmu in (np.linspace(-3.0, 3.0, 241) if grid is None else grid):
        h1 = np.zeros((ns, ns), dtype=np.complex128)
        h1[:n, :n] = _lso(mdl.imphloc) - xmu * np.eye(n)
        for ib in range(mdl.nbath):
            sl = slice(n * (ib + 1), n * (ib + 2))  # getBathStride: Nimp + imp index + (ibath-1)*Nimp, ED_SETUP.f90:367-375
            h1[sl, sl] = _lso(mdl.hbath[..., ib])
            h1[:n, sl] = np.diag(mdl.vbath[:, ib])
            h1[sl, :n] = np.diag(mdl.vbath[:, ib])
        e = np.linalg.eigvalsh(h1)
        if np.sum(e < 0) == ns // 2:
            gap = min(-e[ns // 2 - 1], e[ns // 2])
            if gap > best_gap:
                best, best_gap = xmu, gap
    if best is None:
        return None, 0.0
    m = copy.deepcopy(mdl)
    m.uloc = np.zeros(5)
    m.ust = m.jh = m.jx = m.jp = 0.0
    m.xmu = float(best)
    m.name = mdl.name + "_U0"
    return m, best_gap


def slater_reference(mdl):
    """U = 0 ground state of `noninteracting_half_filled` models as a Slater determinant of the single-particle matrix the
    reference's g0and_bath implies ([[impHloc - xmu, V], [V, Hbath]], checked against g0and_bath in the tests): returns
    (E0, rho) with rho[a, b] = <c^+_a c_b> over the impurity orbitals of one spin, E0 = 2 * sum of the occupied levels."""
    n, ns = mdl.nimp, mdl.ns
    h1 = np.zeros((ns, ns), dtype=np.complex128)
    h1[:n, :n] = _lso(mdl.imphloc) - mdl.xmu * np.eye(n)
    for ib in range(mdl.nbath):
        sl = slice(n * (ib + 1), n * (ib + 2))
        h1[sl, sl] = _lso(mdl.hbath[..., ib])
        h1[:n, sl] = np.diag(mdl.vbath[:, ib])
        h1[sl, :n] = np.diag(mdl.vbath[:, ib])
    e, psi = np.linalg.eigh(h1)
    occ = psi[:, : ns // 2]
    rho = occ[:n, :].conj() @ occ[:n, :].T  # sum_k conj(psi_k(a)) psi_k(b)
    return 2.0 * float(e[: ns // 2].sum()), rho


def slater_cluster_spectrum(rho):
    """Spectrum of the impurity reduced density matrix of a Slater determinant with the same correlation matrix `rho` for
    both spins (Peschel): all products of nu_i or 1 - nu_i over the eigenvalues nu of rho, spin up times spin down."""
    nu = np.clip(np.linalg.eigvalsh(rho), 0.0, 1.0)
    one = np.array([1.0])
    for x in nu:
        one = np.concatenate([one * x, one * (1.0 - x)])
    return np.sort(np.outer(one, one).ravel())


def hubbard_dimer(u, t, hfmode):
    """Two-site Hubbard model as an impurity model whose bath is switched off (V = 0, bath levels at +10): the textbook case
    with the closed-form singlet ground state E0 = U/2 - sqrt(U^2/4 + 4 t^2) (half filling, no chemical potential)."""
    m = models.hubbard_cluster(2, 1, 3, ts=t, uloc=u)  # Ns = 8: sector (1,1) has the 8 x 8 layout the other tiny-sector tests use
    m.vbath = np.zeros_like(m.vbath)
    m.hbath = np.asfortranarray(m.hbath * 0)
    for ib in range(3):
        for a in range(2):
            m.hbath[a, a, 0, 0, 0, 0, ib] = 10.0 + ib
    m.hfmode = hfmode
    m.xmu = 0.0
    m.name = f"dimer_U{u}_t{t}"
    return m
