"""CPU tests that PIN the oracle (the reference ships no golden vectors):
independent Jordan-Wigner ED, dense == sparse == direct == simulated MPI(P), scipy eigsh,
exact Green's function, and the committed golden fixtures."""
import json
import os

import numpy as np
import pytest

from cdmft_lanc_ed_b200 import models
from oracle import jw_ed

HERE = os.path.dirname(os.path.abspath(__file__))

SMALL = [
    lambda: models.random_model(2, 1, 1, seed=1),
    lambda: models.random_model(1, 2, 1, seed=2, kanamori=True),
    lambda: models.random_model(2, 1, 1, nspin=2, seed=3),
    lambda: models.random_model(1, 2, 1, seed=4, kanamori=True, complex_h=False),
    lambda: models.random_model(1, 1, 3, seed=5, hfmode=False),
]


@pytest.mark.parametrize("k", range(len(SMALL)))
def test_dense_hmat_equals_jordan_wigner_all_sectors(oracle_lib, k):
    mdl = SMALL[k]()
    H = jw_ed.full_hamiltonian(mdl)
    assert np.abs(H - H.conj().T).max() < 1e-14
    orc = oracle_lib.Oracle(mdl)
    ns = mdl.ns
    for nup in range(ns + 1):
        for ndw in range(ns + 1):
            isec = models.get_sector(ns, nup, ndw)
            Hd = orc.dense_hmat(isec)
            assert np.abs(jw_ed.sector_hamiltonian(mdl, nup, ndw, H) - Hd).max() < 1e-13
            v = np.random.default_rng(isec).normal(size=Hd.shape[0]) + 1j * np.random.default_rng(isec + 1).normal(size=Hd.shape[0])
            ref = Hd @ v
            for kind, P in [(0, 1), (1, 1), (1, 2), (1, 3), (2, 1), (3, 1), (3, 2), (3, 5)]:
                orc.build_hv_sector(isec, kind, P, 0)
                assert np.abs(orc.hxv(v) - ref).max() < 1e-12
                orc.delete_hv_sector()


def test_maps_operators_and_index_helpers(oracle_lib):
    e = oracle_lib
    m = e.sector_map(8, 4)
    assert m.size == 70 and np.all(np.diff(m) > 0) and all(bin(x).count("1") == 4 for x in m)
    assert e.binomial(16, 8) == 12870 and e.binomial(18, 9) == 48620 and e.binomial(5, -1) == 0
    # c / cdg sign rule and error stops (ED_SETUP.f90:807-833)
    rc, out, sg = e.c_op(3, 0b0111)
    assert (rc, out, sg) == (0, 0b0011, 1.0)
    rc, out, sg = e.c_op(2, 0b0111)
    assert (rc, out, sg) == (0, 0b0101, -1.0)
    assert e.c_op(4, 0b0111)[0] != 0 and e.cdg_op(1, 0b0001)[0] != 0
    rc, out, sg = e.cdg_op(4, 0b0111)
    assert (rc, out, sg) == (0, 0b1111, -1.0)
    for i, s in enumerate(m):
        assert e.binary_search(m, int(s)) == i + 1
    assert e.binary_search(m, 1) == 0
    # sharding arithmetic (ED_HAMILTONIAN.f90:92-105, 197-221)
    tot = 0
    for r in range(8):
        s = e.shard_of(12870, 12870, 8, r)
        assert s["qdw"] == (1609 if r < 6 else 1608) and s["istart"] == tot + 1
        assert e.vecdim(12870, 12870, 8, r) == s["q"] == 12870 * s["qdw"]
        tot += s["q"]
    assert tot == 12870 ** 2


def test_transpose_sim_matches_numpy(oracle_lib):
    rng = np.random.default_rng(0)
    for P, nrow, ncol in [(1, 5, 7), (2, 5, 7), (3, 10, 4), (4, 6, 6), (5, 7, 5)]:
        A = rng.normal(size=(nrow, ncol)) + 1j * rng.normal(size=(nrow, ncol))  # A[i,j]
        a = np.asfortranarray(A).ravel(order="F")  # column blocks concatenated == plain column-major
        b = oracle_lib.vector_transpose_sim(P, nrow, ncol, a)
        assert np.array_equal(b, np.asfortranarray(A.T).ravel(order="F"))


def test_tridiag_eigh_and_lanczos_vs_dense(oracle_lib):
    rng = np.random.default_rng(3)
    n = 40
    d, e = rng.normal(size=n), np.r_[0.0, rng.normal(size=n - 1)]
    w, z = oracle_lib.tridiag_eigh(d, e)
    T = np.diag(d) + np.diag(e[1:], 1) + np.diag(e[1:], -1)
    assert np.abs(w - np.linalg.eigvalsh(T)).max() < 1e-12
    assert np.abs(T @ z - z * w).max() < 1e-12
    mdl = models.hm2x2(1)
    orc = oracle_lib.Oracle(mdl)
    isec = models.get_sector(8, 4, 4)
    w = np.linalg.eigvalsh(orc.dense_hmat(isec))
    orc.build_hv_sector(isec, oracle_lib.SPARSE_SERIAL)
    e0, vec, nit, a, b = orc.lanc_eigh(512, 1e-14)
    assert abs(e0 - w[0]) < 1e-11
    assert np.linalg.norm(orc.hxv(vec) - e0 * vec) < 1e-6
    orc.delete_hv_sector()


def test_e0_vs_scipy_eigsh_K2(oracle_lib):
    """Ns=12 sector (6,6), Dim 853 776: sparse assembly from the oracle CSR + ARPACK."""
    import scipy.sparse as sp
    from scipy.sparse.linalg import eigsh
    mdl = models.hm2x2(2)
    orc = oracle_lib.Oracle(mdl)
    isec = models.get_sector(12, 6, 6)
    orc.build_hv_sector(isec, oracle_lib.SPARSE_SERIAL)
    n = orc.dimup
    rp, col, val = orc.get_csr(1)
    Hup = sp.csr_matrix((val.real, col - 1, rp), shape=(n, n))
    rp, col, val = orc.get_csr(2)
    Hdw = sp.csr_matrix((val.real, col - 1, rp), shape=(n, n))
    H = sp.diags(orc.get_diag()) + sp.kron(sp.identity(n), Hup) + sp.kron(Hdw, sp.identity(n))
    w = eigsh(H.tocsr(), k=1, which="SA", tol=1e-12)[0][0]
    e0, vec, nit, a, b = orc.lanc_eigh(300, 1e-13)
    assert abs(e0 - w) < 1e-9 * abs(w)
    orc.delete_hv_sector()


def test_gimp_vs_exact_full_fock(oracle_lib):
    """4-channel off-diagonal trick + pole/weight accumulation reproduce the exact Lehmann G (Ns=4)."""
    from tests.gf_pipeline import gimp_element
    mdl = models.random_model(2, 1, 1, seed=24)
    H = jw_ed.full_hamiltonian(mdl)
    w = np.linalg.eigvalsh(H)
    assert w[1] - w[0] > 1e-6, "need a non-degenerate ground state"
    Hs = jw_ed.sector_hamiltonian(mdl, 2, 2, H)
    assert abs(np.linalg.eigvalsh(Hs)[0] - w[0]) < 1e-12, "ground state must live in the (Ns/2,Ns/2) sector"
    wm = np.pi / 20.0 * (2 * np.arange(1, 33) - 1)
    G = jw_ed.gimp_exact(mdl, H, wm)
    for ia, ib in [(1, 1), (2, 2), (1, 2), (2, 1)]:
        g = gimp_element("oracle", mdl, ia, ib, wm, edo=oracle_lib)
        assert np.abs(g - G[ia - 1, ib - 1]).max() < 2e-9  # limited by the Lanczos GS vector, not by conventions


def test_golden_fixtures(oracle_lib):
    """tests/golden/*.npz were produced by tests/golden/make_golden.py: oracle outputs; the Ns = 4 cases (meta
    jw = true) were checked against the dense Jordan-Wigner ED when generated, and are re-checked here; the Ns = 8
    cases are plain oracle dumps.  They guard the oracle against drift and feed the GPU tests."""
    path = os.path.join(HERE, "golden", "golden_small.npz")
    g = np.load(path)
    meta = json.loads(str(g["meta"]))
    for case in meta["cases"]:
        mdl = getattr(models, case["builder"])(*case["args"], **case.get("kwargs", {}))
        orc = oracle_lib.Oracle(mdl)
        isec = case["isector"]
        key = case["key"]
        orc.build_hv_sector(isec, oracle_lib.SPARSE_SERIAL)
        assert np.array_equal(oracle_lib.sector_map(mdl.ns, case["nup"]), g[key + "_map_up"])
        rp, col, val = orc.get_csr(1)
        assert np.array_equal(rp, g[key + "_up_rowptr"]) and np.array_equal(col, g[key + "_up_col"])
        assert np.abs(val - g[key + "_up_val"]).max() < 1e-15
        v = g[key + "_v"]
        assert np.abs(orc.hxv(v) - g[key + "_hv"]).max() < 1e-13
        orc.delete_hv_sector()
        if case.get("jw"):
            Hs = jw_ed.sector_hamiltonian(mdl, case["nup"], case["ndw"])
            assert np.abs(Hs @ v - g[key + "_hv"]).max() < 1e-13, key
    assert sum(bool(c.get("jw")) for c in meta["cases"]) >= 4


def test_observables_oracle_against_jordan_wigner_operators(oracle_lib):
    """Pins the oracle's lanc_observables restatement (ED_OBSERVABLES.f90:120-192) with operators built
    independently in the full Fock space: <n_up>, <n_dw>, <n_up n_dw>, <(sum_orb Sz)^2>, <n_a n_b> from dense
    Jordan-Wigner number operators acting on the sector vector embedded in the 4^Ns-dimensional space."""
    from oracle import jw_ed as jw
    from cdmft_lanc_ed_b200 import models
    for mdl, (nup, ndw) in [(models.random_model(2, 1, 1, seed=2), (2, 1)), (models.random_model(2, 2, 0, seed=5), (2, 1)),
                            (models.random_model(1, 2, 1, seed=9, kanamori=True), (2, 2))]:
        ns, nlat, norb = mdl.ns, mdl.nlat, mdl.norb
        isec = models.get_sector(ns, nup, ndw)
        idx = jw.sector_indices(ns, nup, ndw)
        rng = np.random.default_rng(17)
        vec = rng.normal(size=len(idx)) + 1j * rng.normal(size=len(idx))
        vec /= np.linalg.norm(vec)
        full = np.zeros(1 << (2 * ns), dtype=np.complex128)
        full[idx] = vec
        c = jw._ops(2 * ns)
        num = [op.T.conj() @ op for op in c]  # n_p = c_p^+ c_p; spin-up orbitals 0..Ns-1, spin-down Ns..2Ns-1
        ev = lambda A: np.vdot(full, A @ full).real
        ref = oracle_lib.lanc_observables(ns, nlat, norb, isec, vec)
        for il in range(nlat):
            sz_site = 0
            for io in range(norb):
                p = io + il * norb
                assert abs(ref["dens_up"][il, io] - ev(num[p])) < 1e-12
                assert abs(ref["dens_dw"][il, io] - ev(num[ns + p])) < 1e-12
                assert abs(ref["docc"][il, io] - ev(num[p] @ num[ns + p])) < 1e-12
                assert abs(ref["magz"][il, io] - ev(num[p] - num[ns + p])) < 1e-12
                nt = num[p] + num[ns + p]
                assert abs(ref["n2"][il, il, io, io] - ev(nt @ nt)) < 1e-12
                sz_site = sz_site + (num[p] - num[ns + p]) / 2.0
                for jl in range(nlat):
                    for jo in range(io + 1, norb):
                        q = jo + jl * norb
                        assert abs(ref["n2"][il, jl, io, jo] - ev(nt @ (num[q] + num[ns + q]))) < 1e-12
                        szi, szj = (num[p] - num[ns + p]) / 2.0, (num[q] - num[ns + q]) / 2.0
                        assert abs(ref["sz2"][il, jl, io, jo] - ev(szi @ szj)) < 1e-12
            assert abs(ref["s2tot"][il] - ev(sz_site @ sz_site)) < 1e-12


def test_local_energy_oracle_against_jordan_wigner_operators(oracle_lib):
    """Pins the oracle's lanc_local_energy restatement (ED_OBSERVABLES.f90:246-460): <E0> = <sum impHloc c^+ c> over the
    impurity block of both spins, the interaction pieces and the Hartree terms from dense Jordan-Wigner operators on the
    embedded sector vector (complex hoppings, Nspin=2, Norb=2 with Ust/Jh, hfmode on and off)."""
    from oracle import jw_ed as jw
    for mdl, (nup, ndw) in [(models.random_model(2, 1, 1, seed=2), (2, 1)), (models.random_model(2, 1, 1, nspin=2, seed=3), (2, 2)),
                            (models.random_model(1, 2, 1, seed=9, kanamori=True), (2, 2)),
                            (models.random_model(1, 2, 1, seed=4, hfmode=False), (1, 2))]:
        ns, nlat, norb, nimp = mdl.ns, mdl.nlat, mdl.norb, mdl.nlat * mdl.norb
        isec = models.get_sector(ns, nup, ndw)
        idx = jw.sector_indices(ns, nup, ndw)
        rng = np.random.default_rng(23)
        vec = rng.normal(size=len(idx)) + 1j * rng.normal(size=len(idx))
        vec /= np.linalg.norm(vec)
        full = np.zeros(1 << (2 * ns), dtype=np.complex128)
        full[idx] = vec
        c = jw._ops(2 * ns)
        cd = [op.T.conj() for op in c]
        num = [cd[p] @ c[p] for p in range(2 * ns)]
        ev = lambda A: np.vdot(full, A @ full).real
        peso = 0.7
        ref = oracle_lib.Oracle(mdl).lanc_local_energy(isec, vec, peso)
        # <E0>: the impurity block of the one-body matrices of both spins (diagonal included)
        K = 0
        for s, off in ((0, 0), (mdl.nspin - 1, ns)):
            h = jw.one_body_matrix(mdl, s)
            for a in range(nimp):
                for b in range(nimp):
                    if a == b:
                        K = K + mdl.imphloc[a // norb, a // norb, s, s, a % norb, a % norb] * num[off + a]
                    elif h[a, b] != 0:
                        K = K + h[a, b] * cd[off + a] @ c[off + b]
        assert abs(ref["Eknot"] - peso * ev(K)) < 1e-12
        u, ust, jh = mdl.uloc, mdl.ust, mdl.jh
        V, dust, dund, eh = 0, 0, 0, 0
        for il in range(nlat):
            for io in range(norb):
                p = io + il * norb
                V = V + u[io] * num[p] @ num[ns + p]
                if mdl.hfmode:
                    eh = eh - 0.5 * u[io] * (num[p] + num[ns + p]) + 0.25 * (u[p] if p < 5 else 0.0) * np.eye(1 << (2 * ns))
                for jo in range(io + 1, norb):
                    q = jo + il * norb
                    a_ = num[p] @ num[ns + q] + num[q] @ num[ns + p]
                    b_ = num[p] @ num[q] + num[ns + p] @ num[ns + q]
                    V = V + ust * a_ + (ust - jh) * b_
                    dust, dund = dust + a_, dund + b_
                    if mdl.hfmode:
                        nn = num[p] + num[ns + p] + num[q] + num[ns + q]
                        eh = eh - 0.5 * (2 * ust - jh) * nn + 0.25 * (2 * ust - jh) * np.eye(1 << (2 * ns))
        assert abs(ref["Epot"] - peso * ev(V)) < 1e-12
        if norb > 1:
            assert abs(ref["Dust"] - peso * ev(dust)) < 1e-12 and abs(ref["Dund"] - peso * ev(dund)) < 1e-12
        if mdl.hfmode:
            assert abs(ref["Ehartree"] - peso * ev(eh)) < 1e-12
        else:
            assert ref["Ehartree"] == 0.0


def test_density_matrices_oracle_against_jordan_wigner(oracle_lib):
    """Pins the oracle's density_matrix_impurity restatement (ED_OBSERVABLES.f90:465-686): the cluster density matrix is the
    partial trace over the bath of |vec><vec| in the reference's labelling io = IimpUp + 2^Nimp IimpDw (einsum on the
    full-Fock vector reshaped to [bath_dw, imp_dw, bath_up, imp_up]); the single-particle one is <c^+_a c_b> from dense
    Jordan-Wigner operators."""
    from oracle import jw_ed as jw
    for mdl, (nup, ndw) in [(models.random_model(2, 1, 1, seed=2), (2, 1)), (models.random_model(2, 1, 1, nspin=2, seed=3), (2, 2)),
                            (models.random_model(1, 2, 1, seed=9, kanamori=True), (1, 3))]:
        ns, nlat, norb, nimp = mdl.ns, mdl.nlat, mdl.norb, mdl.nlat * mdl.norb
        nb = ns - nimp
        isec = models.get_sector(ns, nup, ndw)
        idx = jw.sector_indices(ns, nup, ndw)
        rng = np.random.default_rng(31)
        vec = rng.normal(size=len(idx)) + 1j * rng.normal(size=len(idx))
        vec /= np.linalg.norm(vec)
        full = np.zeros(1 << (2 * ns), dtype=np.complex128)
        full[idx] = vec
        peso = 0.8
        cdm, sp = oracle_lib.Oracle(mdl).density_matrix_impurity(isec, vec, peso)
        # Fock index = mup + (mdw << Ns), m = imp + (bath << Nimp): C-order reshape [bath_dw, imp_dw, bath_up, imp_up]
        T = full.reshape(1 << nb, 1 << nimp, 1 << nb, 1 << nimp)
        rho = peso * np.einsum("pdqu,pDqU->udUD", T, T.conj())  # [Iup, Idw, Jup, Jdw]
        ni = 1 << nimp
        R = np.zeros((ni * ni, ni * ni), dtype=np.complex128)
        for Iu in range(ni):
            for Id in range(ni):
                for Ju in range(ni):
                    for Jd in range(ni):
                        R[Iu + ni * Id, Ju + ni * Jd] = rho[Iu, Id, Ju, Jd]
        assert np.abs(cdm - R).max() < 1e-13
        assert abs(np.trace(cdm).real - peso) < 1e-12 and np.abs(cdm - cdm.conj().T).max() < 1e-13
        c = jw._ops(2 * ns)
        for s, off in ((0, 0), (mdl.nspin - 1, ns)):
            for a in range(nimp):
                for b in range(nimp):
                    ev = peso * np.vdot(full, c[off + a].T.conj() @ (c[off + b] @ full))
                    got = sp[a // norb, b // norb, s, s, a % norb, b % norb]
                    if mdl.nspin == 1 and off == ns:
                        continue  # Nspin = 1: the reference fills the spin-up block only
                    assert abs(got - ev) < 1e-12, (a, b, s)


@pytest.mark.parametrize("case", ["models.hm2x2(1)", "models.bhz2(1)", "models.random_model(2, 2, 1, seed=31)",
                                  "models.random_model(3, 1, 1, seed=32)", "models.hubbard_cluster(2, 1, 2)"])
def test_u0_gimp_equals_the_references_g0and_bath(oracle_lib, case):
    """An anchor that comes from the REFERENCE itself, not from this repository's reading of it: at U = 0 the impurity
    Green's function of the ED pipeline (sector Hamiltonian -> ground-state Lanczos -> c / c^+ start vectors ->
    tridiagonalisation -> pole sums, ED_GF_NORMAL.f90) must equal the reference's analytic non-interacting function
    g0and_bath (ED_BATH_FUNCTIONS.f90:102-155: G0^-1 = (iw + xmu) - impHloc - sum_ib V (iw - Hbath_ib)^-1 V).  This pins
    the bath-orbital layout (getBathStride), the hybridisation and replica-hopping conventions, the sign of xmu, the
    fermionic signs of the hops and the whole Green's-function flow, diagonal and off-diagonal (4-channel) elements."""
    from tests.gf_pipeline import g0and_bath, gimp_element, noninteracting_half_filled
    mdl, gap = noninteracting_half_filled(eval(case))
    assert mdl is not None and gap > 1e-2, "no half-filled non-degenerate ground state on the xmu grid"
    wm = np.pi / 40.0 * (2 * np.arange(1, 25) - 1)
    g0 = g0and_bath(mdl, 1j * wm)
    o = _Oracle_gs(mdl, oracle_lib)
    n = mdl.nimp
    pairs = [(a, a) for a in range(1, n + 1)] + [(1, 2), (2, 1)] + ([(1, n), (n, 2)] if n > 2 else [])
    for ia, ib in pairs:
        g = gimp_element("oracle", mdl, ia, ib, wm, edo=oracle_lib, gs=o)
        # limited by the Lanczos ground-state vector (gaps down to 0.14), not by conventions: a convention error is O(1)
        assert np.abs(g - g0[ia - 1, ib - 1]).max() < 5e-8 * max(1.0, np.abs(g0[ia - 1, ib - 1]).max()), (case, ia, ib)


def _Oracle_gs(mdl, edo):
    """ground state of the sector (Ns/2, Ns/2) once, shared by all elements"""
    o = edo.Oracle(mdl)
    o.build_hv_sector(models.get_sector(mdl.ns, mdl.ns // 2, mdl.ns // 2), edo.SPARSE_SERIAL)
    e0, vec, _, _, _ = o.lanc_eigh(512, 1e-14)  # (a threshold below what the recurrence resolves runs 512 steps and spoils the vector)
    o.delete_hv_sector()
    return e0, vec


@pytest.mark.parametrize("case", ["models.hm2x2(1)", "models.bhz2(1)", "models.random_model(3, 1, 1, seed=32)"])
def test_u0_observables_equal_the_slater_determinant(oracle_lib, case):
    """The N4 restatements (lanc_observables, lanc_local_energy, density_matrix_impurity) on the U = 0 ground state against
    the Slater determinant of the single-particle problem that the reference's g0and_bath fixes: E0 = 2 sum of occupied
    levels, <C^+_a C_b> = the correlation matrix, dens = 2 rho_aa, docc = rho_aa^2 (the spins factorise), the spectrum
    of the cluster density matrix = Peschel's products of nu_i / (1 - nu_i), <E0> = 2 Re sum impHloc(a,b) rho(a,b)."""
    from tests.gf_pipeline import noninteracting_half_filled, slater_cluster_spectrum, slater_reference
    mdl, gap = noninteracting_half_filled(eval(case))
    assert mdl is not None and gap > 1e-2
    e0_ref, rho = slater_reference(mdl)
    e0, vec = _Oracle_gs(mdl, oracle_lib)
    isec = models.get_sector(mdl.ns, mdl.ns // 2, mdl.ns // 2)
    # the Hamiltonian carries the Hartree-Fock constant of HFMODE only through U: none at U = 0
    assert abs(e0 - e0_ref) < 1e-10 * max(1.0, abs(e0_ref))
    cdm, sp = oracle_lib.Oracle(mdl).density_matrix_impurity(isec, vec, 1.0)
    n, norb = mdl.nimp, mdl.norb
    for a in range(n):
        for b in range(n):
            assert abs(sp[a // norb, b // norb, 0, 0, a % norb, b % norb] - rho[a, b]) < 2e-7, (a, b)
    obs = oracle_lib.lanc_observables(mdl.ns, mdl.nlat, mdl.norb, isec, vec)
    d = np.real(np.diag(rho)).reshape(mdl.nlat, mdl.norb)
    assert np.abs(np.asarray(obs["dens"]).reshape(d.shape) - 2 * d).max() < 2e-7
    assert np.abs(np.asarray(obs["docc"]).reshape(d.shape) - d * d).max() < 2e-7
    assert np.abs(np.sort(np.linalg.eigvalsh(cdm)) - slater_cluster_spectrum(rho)).max() < 2e-7
    # lanc_local_energy: <E0> = sum_ab impHloc(a,b) <c^+_a c_b> over both spins; no interaction energy at U = 0
    from tests.gf_pipeline import _lso
    en = oracle_lib.Oracle(mdl).lanc_local_energy(isec, vec, 1.0)
    assert abs(en["Eknot"] - 2.0 * np.real(np.sum(_lso(mdl.imphloc) * rho))) < 2e-7
    assert en["Epot"] == 0.0 and en["Ehartree"] == 0.0


def test_one_orbital_density_matrix_against_the_reference_drivers_benchmark(oracle_lib):
    """The reference's only in-tree self-check (drivers/cdn_hm_2dsquare.f90:443-464, `one_orb_benchmark`, Eq. 4 of
    Mod. Phys. Lett. B 27 (2013) 05): the reduced density matrix of ONE impurity orbital, obtained by tracing the cluster
    density matrix over the other orbitals, has the diagonal (1 - n_up - n_dw + docc, n_up - docc, n_dw - docc, docc).
    Ties the labelling of cluster_density_matrix (IimpUp + 2^Nimp IimpDw) to the observables of lanc_observables."""
    for mdl, (nup, ndw) in [(models.hm2x2(1), (4, 3)), (models.random_model(2, 2, 1, nspin=2, seed=12), (3, 5)),
                            (models.bhz2(1, kanamori=True), (4, 4))]:
        ns, nimp, norb = mdl.ns, mdl.nimp, mdl.norb
        isec = models.get_sector(ns, nup, ndw)
        orc = oracle_lib.Oracle(mdl)
        rng = np.random.default_rng(77)
        dim = orc.get_dim(isec)[0]
        vec = rng.normal(size=dim) + 1j * rng.normal(size=dim)
        vec /= np.linalg.norm(vec)
        cdm, _ = orc.density_matrix_impurity(isec, vec, 1.0)
        obs = oracle_lib.lanc_observables(ns, mdl.nlat, norb, isec, vec)
        p = np.real(np.diag(cdm))
        ni = 1 << nimp
        iu, idw = np.arange(ni * ni) % ni, np.arange(ni * ni) // ni
        for a in range(nimp):
            il, io = a // norb, a % norb
            nu, nd, d = obs["dens_up"][il][io], obs["dens_dw"][il][io], obs["docc"][il][io]
            bu, bd = (iu >> a) & 1, (idw >> a) & 1
            one_orb = [p[(bu == x) & (bd == y)].sum() for (x, y) in ((0, 0), (1, 0), (0, 1), (1, 1))]
            assert np.abs(np.array(one_orb) - np.array([1 - nu - nd + d, nu - d, nd - d, d])).max() < 1e-13, (mdl.name, a)


@pytest.mark.parametrize("u,t", [(2.0, 0.25), (4.0, 1.0), (0.5, 1.3)])
def test_hubbard_dimer_closed_form(oracle_lib, u, t):
    """A literature anchor that owes nothing to this repository: the two-site Hubbard model (bath switched off) has the
    singlet ground-state energy U/2 - sqrt(U^2/4 + 4t^2) in the sector (1,1); with HFMODE the Hamiltonian carries
    -U/2 (n_up + n_dw) + U/4 per site (sparse/H_local.f90:61-77), i.e. -U/2 in that sector.  The triplet sits at 0 (-U/2)."""
    from tests.gf_pipeline import hubbard_dimer
    for hf in (False, True):
        mdl = hubbard_dimer(u, t, hf)
        orc = oracle_lib.Oracle(mdl)
        w = np.linalg.eigvalsh(orc.dense_hmat(models.get_sector(mdl.ns, 1, 1)))
        shift = -u / 2 if hf else 0.0
        assert abs(w[0] - (u / 2 - np.sqrt(u * u / 4 + 4 * t * t) + shift)) < 1e-12
        wt = np.linalg.eigvalsh(orc.dense_hmat(models.get_sector(mdl.ns, 2, 0)))  # both electrons up: the triplet
        assert abs(wt[0] - shift) < 1e-12
