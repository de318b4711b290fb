"""Independent full-Fock Jordan-Wigner exact diagonalisation (numpy, tiny systems).

TEST INFRASTRUCTURE ONLY -- the anchor that pins the oracle's reading of the
reference conventions (Fock ordering, fermionic signs, H = D + 1(x)Hup + Hdw(x)1,
orientation <i|H|j> = h(is,js) for c^+_is c_js, Kanamori terms) since the
reference ships no golden vectors.  It shares NO code with
ed_oracle.c: operators are dense 2^(2Ns) matrices built from Kronecker products.

Qubit order: spin-up orbitals 0..Ns-1, then spin-down orbitals Ns..2Ns-1; the
Fock index of |mup, mdw> is mup + (mdw << Ns).  Orbital numbering inside one spin
follows ED_SETUP.f90:367-375,563-568 (impurity iorb+(ilat-1)*Norb, then replicas).
"""
from __future__ import annotations

import numpy as np


def _ops(nq):
    """annihilators c_p (dense) with JW strings; c_p = (prod_{q<p} Z_q) sigma^-_p."""
    I2 = np.eye(2)
    Z = np.diag([1.0, -1.0])
    sm = np.array([[0.0, 1.0], [0.0, 0.0]])  # |0><1|
    out = []
    for p in range(nq):
        m = np.array([[1.0]])
        # kron order: qubit 0 is the least-significant bit -> rightmost factor
        for q in range(nq - 1, -1, -1):
            f = sm if q == p else (Z if q < p else I2)
            m = np.kron(m, f)
        out.append(m)
    return out


def one_body_matrix(model, spin_block):
    """[Ns,Ns] single-particle matrix h(a,b) of sum h(a,b) c^+_a c_b for one spin
    (spin_block = 0 for up, Nspin-1 for down), including diagonal entries."""
    L, O, B = model.nlat, model.norb, model.nbath
    nimp = L * O
    ns = nimp * (B + 1)
    h = np.zeros((ns, ns), dtype=np.complex128)
    s = spin_block
    for il in range(L):
        for jl in range(L):
            for io in range(O):
                for jo in range(O):
                    a, b = io + il * O, jo + jl * O
                    h[a, b] += model.imphloc[il, jl, s, s, io, jo]
                    for ib in range(B):
                        aa, bb = nimp + a + ib * nimp, nimp + b + ib * nimp
                        h[aa, bb] += model.hbath[il, jl, s, s, io, jo, ib]
    for il in range(L):
        for io in range(O):
            a = io + il * O
            for ib in range(B):
                v = model.vbath[io + il * O + s * O * L, ib]
                aa = nimp + a + ib * nimp
                h[a, aa] += v
                h[aa, a] += v
    return h


def full_hamiltonian(model):
    """Dense 4^Ns Hamiltonian. Hermitian iff the inputs are."""
    L, O, S = model.nlat, model.norb, model.nspin
    ns = model.ns
    nq = 2 * ns
    c = _ops(nq)
    cd = [m.T for m in c]
    n = [cd[p] @ c[p] for p in range(nq)]
    dim = 1 << nq
    H = np.zeros((dim, dim), dtype=np.complex128)
    hup = one_body_matrix(model, 0)
    hdw = one_body_matrix(model, S - 1)
    # bath diagonal uses only the REAL part (bath_diag = DREAL(...), ED_HAMILTONIAN_SPARSE_HxV.f90:70);
    # impurity diagonal uses the full (complex) value as the reference does (H_local.f90:24-25)
    nimp = L * O
    for a in range(ns):
        for b in range(ns):
            for (h, off) in ((hup, 0), (hdw, ns)):
                val = h[a, b]
                if a == b and a >= nimp:
                    val = val.real
                if val != 0:
                    H += val * (cd[a + off] @ c[b + off])
    Id = np.eye(dim)
    U, Ust, Jh, Jx, Jp, xmu = model.uloc, model.ust, model.jh, model.jx, model.jp, model.xmu
    for il in range(L):
        for io in range(O):
            a = io + il * O
            nu, nd = n[a], n[a + ns]
            H += -xmu * (nu + nd)
            H += U[io] * (nu @ nd)
            if model.hfmode:
                H += -0.5 * U[io] * (nu + nd) + 0.25 * U[io] * Id
        if O > 1:
            for io in range(O):
                for jo in range(io + 1, O):
                    a, b = io + il * O, jo + il * O
                    H += Ust * (n[a] @ n[b + ns] + n[b] @ n[a + ns])
                    H += (Ust - Jh) * (n[a] @ n[b] + n[a + ns] @ n[b + ns])
                    if model.hfmode:
                        tot = n[a] + n[a + ns] + n[b] + n[b + ns]
                        H += -0.5 * Ust * tot + 0.25 * Ust * Id
                        H += -0.5 * (Ust - Jh) * tot + 0.25 * (Ust - Jh) * Id
            if model.jhflag:
                for io in range(O):
                    for jo in range(O):
                        if io == jo:
                            continue
                        a, b = io + il * O, jo + il * O
                        # spin exchange: Jx (c^+_a,up c_b,up)(c^+_b,dw c_a,dw)
                        if Jx != 0:
                            H += Jx * (cd[a] @ c[b]) @ (cd[b + ns] @ c[a + ns])
                        # pair hopping: Jp (c^+_a,up c_b,up)(c^+_a,dw c_b,dw)
                        if Jp != 0:
                            H += Jp * (cd[a] @ c[b]) @ (cd[a + ns] @ c[b + ns])
    return H


def sector_indices(ns, nup, ndw):
    """Fock indices of the (nup,ndw) sector ordered like the reference: i = iup + idw*DimUp,
    both maps ascending in the integer value (ED_SETUP.f90:749-769)."""
    ups = [s for s in range(1 << ns) if bin(s).count("1") == nup]
    dws = [s for s in range(1 << ns) if bin(s).count("1") == ndw]
    return np.array([u + (d << ns) for d in dws for u in ups], dtype=np.int64)


def sector_hamiltonian(model, nup, ndw, H=None):
    if H is None:
        H = full_hamiltonian(model)
    idx = sector_indices(model.ns, nup, ndw)
    return H[np.ix_(idx, idx)]


def gimp_exact(model, H, wm, e0_tol=1e-9):
    """Exact T=0 impurity Matsubara Green's function G_ab(i wm) for spin up from the full
    spectrum, a,b over impurity orbitals; averages over degenerate ground states."""
    ns = model.ns
    nq = 2 * ns
    c = _ops(nq)
    w, U = np.linalg.eigh(H)
    gs = np.where(w - w[0] < e0_tol)[0]
    nimp = model.nimp
    G = np.zeros((nimp, nimp, len(wm)), dtype=np.complex128)
    for g in gs:
        psi = U[:, g]
        for a in range(nimp):
            for b in range(nimp):
                # <c_a (z-(H-E0))^-1 c^+_b> + <c^+_b (z+(H-E0))^-1 c_a>
                xb = U.conj().T @ (c[b].T @ psi)
                xa = U.conj().T @ (c[a].T @ psi)
                ya = U.conj().T @ (c[a] @ psi)
                yb = U.conj().T @ (c[b] @ psi)
                for k, om in enumerate(wm):
                    z = 1j * om
                    G[a, b, k] += np.sum(xa.conj() * xb / (z - (w - w[0])))
                    G[a, b, k] += np.sum(yb.conj() * ya / (z + (w - w[0])))
    return G / len(gs)
