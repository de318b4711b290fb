"""CPU check of the Green's-function ORCHESTRATION of cdmft_lanc_ed_b200/gf_normal.py over a state list (finite temperature,
degenerate ground states, both spins, ed_gf_symmetric) -- the flow of ED_GF_NORMAL.f90:38-106 with `do istate=1,state_list%size`
inside every channel routine and the Boltzmann weights of add_to_lanczos_gf_normal (:930-936).  The product flow is run with
an adapter over the CPU oracle in place of the device entry points (sector build, c / c^+, tridiagonalisation); the pole sums
are the product's own host routine (cdmft_b200_add_to_lanczos_gf_full needs no device).  Reference: the EXACT Lehmann
representation from a dense full-Fock Jordan-Wigner diagonalisation (oracle/jw_ed.py), every eigenstate in the list."""
from math import comb

import numpy as np
import pytest

from cdmft_lanc_ed_b200 import ed_hamiltonian as E
from cdmft_lanc_ed_b200 import gf_normal, models
from oracle import edo, jw_ed


class OracleBackend:
    """the six calls build_gf_normal_states makes, served by the checker (tests only)"""

    def __init__(self, mdl):
        self.ns = mdl.ns
        self.o = edo.Oracle(mdl)

    def getDim(self, isec):
        nup, ndw = (isec - 1) // (self.ns + 1), (isec - 1) % (self.ns + 1)
        return comb(self.ns, nup) * comb(self.ns, ndw), comb(self.ns, nup), comb(self.ns, ndw)

    def apply_op(self, isec, iop, ispin, pos, coef, state):
        return edo.apply_op(self.ns, isec, iop, ispin, pos, coef, state)

    def build_Hv_sector(self, jsec):
        self.o.build_hv_sector(jsec, edo.SPARSE_SERIAL)
        return self.o.dim

    def delete_Hv_sector(self):
        self.o.delete_hv_sector()

    def sp_lanc_tridiag(self, v, nlanc):
        return self.o.lanc_tridiag(v, nlanc)

    add_to_lanczos_gf_normal_full = staticmethod(E.add_to_lanczos_gf_normal_full)


def _all_states(mdl):
    """every eigenstate of every sector: (isector, energy, vector) -- a complete state_list"""
    o = edo.Oracle(mdl)
    out = []
    for nup in range(mdl.ns + 1):
        for ndw in range(mdl.ns + 1):
            isec = models.get_sector(mdl.ns, nup, ndw)
            w, u = np.linalg.eigh(o.dense_hmat(isec))
            out += [(isec, float(w[k]), np.ascontiguousarray(u[:, k])) for k in range(len(w))]
    return out


def _exact(mdl, wm, beta, spin_off, states_e0_only=False, zs=None):
    """Lehmann sum over the full Fock space: (1/Z) sum_n exp(-beta (E_n - E0)) [ <n|c_a (z - (H - E_n))^-1 c_b^+|n> +
    <n|c_b^+ (z + (H - E_n))^-1 c_a|n> ]; states_e0_only: average over the (degenerate) ground states instead."""
    H = jw_ed.full_hamiltonian(mdl)
    w, U = np.linalg.eigh(H)
    c = jw_ed._ops(2 * mdl.ns)
    nimp = mdl.nimp
    if states_e0_only:
        p = (w - w[0] < 1e-9).astype(float)
    else:
        p = np.exp(-beta * (w - w[0]))
    p /= p.sum()
    C = [U.conj().T @ (c[spin_off + a] @ U) for a in range(nimp)]  # C[a][m, n] = <m| c_a |n>
    de = w[:, None] - w[None, :]  # E_m - E_n
    zs = 1j * np.asarray(wm) if zs is None else np.asarray(zs)  # Matsubara points unless complex frequencies are given
    G = np.zeros((nimp, nimp, len(zs)), dtype=np.complex128)
    for a in range(nimp):
        for b in range(nimp):
            part = C[a] * C[b].conj()       # [n, m]: <n|c_a|m> <m|c_b^+|n> = C_a[n, m] conj(C_b[n, m])
            hole = C[b].conj() * C[a]       # [m, n]: <n|c_b^+|m> <m|c_a|n>
            for k, z in enumerate(zs):
                G[a, b, k] = np.sum(p[:, None] * part / (z - de.T)) + np.sum(p[None, :] * hole / (z + de))
    return G


@pytest.mark.parametrize("case,nspin", [("models.random_model(2, 1, 1, seed=24)", 1), ("models.random_model(2, 1, 1, nspin=2, seed=3)", 2)])
def test_finite_temperature_gf_over_the_whole_state_list(case, nspin):
    mdl = eval(case)
    beta = 1.5
    wm = np.pi / beta * (2 * np.arange(1, 13) - 1)
    states = _all_states(mdl)
    assert len(states) == 4 ** mdl.ns
    G, _ = gf_normal.build_gf_normal_states(mdl.nimp, states, wm, nspin=nspin, finite_t=True, beta=beta, backend=OracleBackend(mdl))
    for s in range(nspin):
        ref = _exact(mdl, wm, beta, spin_off=s * mdl.ns)
        assert np.abs(G[s] - ref).max() < 1e-10 * np.abs(ref).max(), (case, s)
    if nspin == 2:
        assert np.abs(G[0] - G[1]).max() > 1e-3  # the two spin blocks really differ in this model


def test_degenerate_ground_states_at_zero_temperature():
    """T = 0 with a degenerate ground state: zeta_function = number of ground states (ED_DIAG.f90:365), each contributes 1/zeta."""
    mdl = models.random_model(2, 1, 1, seed=5)  # Ns = 4; the ground state of this model is a doublet (odd particle number)
    allst = _all_states(mdl)
    emin = min(e for _, e, _ in allst)
    gs = [s for s in allst if s[1] - emin < 1e-9]
    assert len(gs) == 2 and gs[0][0] != gs[1][0]  # two sectors
    wm = np.pi / 20.0 * (2 * np.arange(1, 9) - 1)
    G, _ = gf_normal.build_gf_normal_states(mdl.nimp, gs, wm, finite_t=False, backend=OracleBackend(mdl))
    ref = _exact(mdl, wm, 0.0, 0, states_e0_only=True)
    assert np.abs(G[0] - ref).max() < 1e-10 * np.abs(ref).max()
    assert gf_normal.zeta_function([e for _, e, _ in gs], False, 0.0) == len(gs)


def test_symmetric_two_channel_variant_and_real_axis():
    """ed_gf_symmetric (lanc_build_gf_normal_mix_chan2): two mixed channels per pair and chan4 = 0 -- exact when G_ab = G_ba
    (real symmetric hoppings); the real-axis function accumulated next to the Matsubara one."""
    mdl = models.hubbard_cluster(2, 1, 1)  # Ns = 4, real
    beta = 2.0
    wm = np.pi / beta * (2 * np.arange(1, 9) - 1)
    states = _all_states(mdl)
    be = OracleBackend(mdl)
    wr, eps = np.linspace(-2.5, 2.5, 11), 0.07
    G2, G2r = gf_normal.build_gf_normal_states(mdl.nimp, states, wm, wr=wr, eps=eps, finite_t=True, beta=beta, symmetric=True, backend=be)
    n2 = gf_normal.build_gf_normal_states.last_sector_builds
    G4, _ = gf_normal.build_gf_normal_states(mdl.nimp, states, wm, finite_t=True, beta=beta, symmetric=False, backend=be)
    ref = _exact(mdl, wm, beta, 0)
    assert np.abs(ref[0, 1] - ref[1, 0]).max() < 1e-13
    assert np.abs(G2[0] - ref).max() < 1e-10 * np.abs(ref).max() and np.abs(G4[0] - ref).max() < 1e-10 * np.abs(ref).max()
    refr = _exact(mdl, None, beta, 0, zs=wr + 1j * eps)  # impGreal: the same poles and weights at w + i eps (:968-971)
    assert np.abs(G2r[0] - refr).max() < 1e-10 * np.abs(refr).max()
    assert n2 <= 2 * len(states)  # channels batched by target sector: at most two sector builds per state


# ---- observables over a state list (cdmft_lanc_ed_b200/observables.py) against exact thermal averages -------------------
class OracleObsBackend:
    def __init__(self, mdl):
        self.m = mdl
        self.o = edo.Oracle(mdl)

    def build(self, isector):
        pass

    def delete(self):
        pass

    def observables(self, isector, vec, peso):
        return edo.lanc_observables(self.m.ns, self.m.nlat, self.m.norb, isector, vec, peso)

    def local_energy(self, isector, vec, peso):
        return self.o.lanc_local_energy(isector, vec, peso)

    def density_matrices(self, isector, vec, peso):
        return self.o.density_matrix_impurity(isector, vec, peso)


@pytest.mark.parametrize("case", ["models.random_model(2, 1, 1, seed=24)", "models.random_model(1, 2, 1, seed=9, kanamori=True)"])
def test_observables_over_the_state_list_against_thermal_averages(case):
    """dens, docc, <C^+_a C_b> and the trace of the cluster density matrix summed over a complete state list with Boltzmann
    weights = Tr(rho O) with rho = exp(-beta H)/Z from the dense Jordan-Wigner ED."""
    from cdmft_lanc_ed_b200 import observables
    mdl = eval(case)
    beta = 1.3
    states = _all_states(mdl)
    res = observables.observables_states(mdl, states, finite_t=True, beta=beta, backend=OracleObsBackend(mdl))
    H = jw_ed.full_hamiltonian(mdl)
    w, U = np.linalg.eigh(H)
    p = np.exp(-beta * (w - w[0]))
    assert abs(res["zeta_function"] - p.sum()) < 1e-10 * p.sum()
    rho = (U * (p / p.sum())) @ U.conj().T
    c = jw_ed._ops(2 * mdl.ns)
    ns, nimp, norb = mdl.ns, mdl.nimp, mdl.norb
    for a in range(nimp):
        nu, nd = c[a].T.conj() @ c[a], c[ns + a].T.conj() @ c[ns + a]
        il, io = a // norb, a % norb
        assert abs(res["dens"][il, io] - np.trace(rho @ (nu + nd)).real) < 1e-11
        assert abs(res["docc"][il, io] - np.trace(rho @ nu @ nd).real) < 1e-11
        for b in range(nimp):
            ev = np.trace(rho @ c[a].T.conj() @ c[b])
            assert abs(res["single_particle_density_matrix"][il, b // norb, 0, 0, io, b % norb] - ev) < 1e-11
    assert abs(np.trace(res["cluster_density_matrix"]).real - 1.0) < 1e-11


# ---- the diagonalisation step (cdmft_lanc_ed_b200/ed_diag.py): sector loop + state-list rules of ED_DIAG.f90 -------------
class OracleDiagBackend:
    def __init__(self, mdl):
        self.m, self.o, self.isec = mdl, edo.Oracle(mdl), None

    def build(self, isector):
        self.o.build_hv_sector(isector, edo.SPARSE_SERIAL)
        self.isec = isector
        return self.o.dim

    def delete(self):
        self.o.delete_hv_sector()

    def dense_hmat(self):
        return edo.Oracle(self.m).dense_hmat(self.isec)

    def sp_eigh(self, neigen, nblock, nitermax, tol):
        w, z, _ = E.eigh_logic_host(self.o.hxv, self.o.dim, neigen, nblock=nblock, nitermax=nitermax, tol=tol)
        return w, z

    def sp_lanc_eigh(self, n, nitermax, tol):
        e0, vec, _, _, _ = self.o.lanc_eigh(nitermax, max(tol, 1e-14))
        return np.array([e0]), vec.reshape(-1, 1)


def test_ed_diag_state_list_rules():
    """ed_diag_c + ed_post_diag: T = 0 keeps exactly the (degenerate) ground states of the whole Fock space; at finite
    temperature the list converges, through the reference's own adaptation of neigen_sector / lanc_nstates_total, to the
    lowest states of the exact spectrum up to the Boltzmann cut-off.  Krylov branch (sp_eigh logic / sp_lanc_eigh with the
    oracle's mat-vec) and LAPACK branch (dense Hmat) both exercised via lanc_dim_threshold."""
    from cdmft_lanc_ed_b200 import ed_diag
    mdl = models.random_model(3, 1, 1, seed=5)  # Ns = 6, sectors up to 400 states
    exact = np.sort(np.concatenate([np.linalg.eigvalsh(edo.Oracle(mdl).dense_hmat(i)) for i in range(1, (mdl.ns + 1) ** 2 + 1)]))
    be = OracleDiagBackend(mdl)
    # T = 0, default method, Krylov solver for every sector above 30 states
    p0 = ed_diag.DiagParams(finite_t=False, lanc_nstates_total=1, lanc_dim_threshold=30, lanc_tolerance=1e-13)
    st = ed_diag.ed_diag(mdl.ns, p0, backend=be)
    st, zeta, numgs = ed_diag.ed_post_diag(mdl.ns, st, p0)
    ngs = int(np.sum(exact - exact[0] <= 1e-9))
    assert len(st) == ngs == numgs and zeta == float(ngs)
    assert all(abs(e - exact[0]) < 1e-9 for _, e, _ in st)
    for isec, e, v in st:  # eigenpairs of their sectors
        h = edo.Oracle(mdl).dense_hmat(isec)
        assert np.linalg.norm(h @ v - e * v) < 1e-7
    # T = 0 with the plain Lanczos method: same ground-state energy
    pl = ed_diag.DiagParams(finite_t=False, lanc_method="lanczos", lanc_nstates_total=1, lanc_dim_threshold=30)
    sl = ed_diag.ed_diag(mdl.ns, pl, backend=be)
    assert abs(sl[0][1] - exact[0]) < 1e-9
    # finite temperature: iterate diag / post_diag as the DMFT loop does; the list must settle on the exact low-energy spectrum
    beta = 30.0  # cut-off 1e-4 -> levels within 0.307 of the ground state: 7 states; the list grows by lanc_nstates_step per call
    pf = ed_diag.DiagParams(finite_t=True, beta=beta, lanc_nstates_sector=2, lanc_nstates_total=4, lanc_nstates_step=2,
                            lanc_dim_threshold=30, lanc_tolerance=1e-13, cutoff=1e-4)
    for _ in range(8):
        st = ed_diag.ed_diag(mdl.ns, pf, backend=be)
        st, zeta, numgs = ed_diag.ed_post_diag(mdl.ns, st, pf)
    want = exact[np.exp(-beta * (exact - exact[0])) > pf.cutoff]
    got = np.array([e for _, e, _ in st])
    assert len(got) == len(want) and np.abs(got - want).max() < 1e-8, (got, want)
    # the reference sums zeta_function BEFORE it trims the list (:355-366 vs :455-469): the trimmed states are still in it
    assert 0.0 <= zeta - np.exp(-beta * (want - want[0])).sum() < 10 * pf.cutoff
