"""CPU checks of the local-observables path (ED_OBSERVABLES.f90:94-236): the host mirror evaluates
the reference's master loop on the impurity-configuration weights W[mu, md]; here W is built with numpy from the
oracle's sector maps and the result is compared with the oracle's state-by-state restatement of the loop."""
import numpy as np
import pytest

from cdmft_lanc_ed_b200 import ed_hamiltonian as E
from cdmft_lanc_ed_b200 import models
from oracle import edo


def weights_numpy(mdl, nup, ndw, vec):
    mu_map, md_map = edo.sector_map(mdl.ns, nup), edo.sector_map(mdl.ns, ndw)
    nmu = 1 << mdl.nimp
    w2 = (np.abs(vec) ** 2).reshape(len(md_map), len(mu_map))  # [idw, iup], iup fastest in memory
    W = np.zeros((nmu, nmu))
    np.add.at(W, (np.asarray(mu_map)[None, :] & (nmu - 1), np.asarray(md_map)[:, None] & (nmu - 1)), w2)
    return W


@pytest.mark.parametrize("case", [("hm2x2(1)", (4, 4)), ("hm2x2(1)", (3, 5)), ("bhz2(1)", (3, 2)), ("bhz2(1)", (4, 4)),
                                  ("random_model(2, 2, 1, seed=3)", (2, 5)), ("random_model(3, 1, 1, seed=4)", (3, 3))])
def test_observables_from_weights_match_the_reference_loop(case):
    mdl = eval("models." + case[0])
    nup, ndw = case[1]
    isec = models.get_sector(mdl.ns, nup, ndw)
    dim = edo.binomial(mdl.ns, nup) * edo.binomial(mdl.ns, ndw)
    rng = np.random.default_rng(7)
    vec = rng.normal(size=dim) + 1j * rng.normal(size=dim)
    vec /= np.linalg.norm(vec)
    ref = edo.lanc_observables(mdl.ns, mdl.nlat, mdl.norb, isec, vec, peso=0.7)
    got = E.observables_from_weights(weights_numpy(mdl, nup, ndw, vec), mdl.nlat, mdl.norb, peso=0.7)
    for k in ("dens", "dens_up", "dens_dw", "docc", "magz", "s2tot", "sz2", "n2"):
        assert np.abs(got[k] - ref[k]).max() < 1e-13, k
    # sum rules: particle numbers on the impurity never exceed the sector's, weights sum to peso
    assert abs(weights_numpy(mdl, nup, ndw, vec).sum() - 1.0) < 1e-12
    assert ref["dens_up"].sum() <= nup * 0.7 + 1e-12 and ref["dens_dw"].sum() <= ndw * 0.7 + 1e-12


def test_add_to_lanczos_gf_all_branches_host_logic(oracle_lib):
    """add_to_lanczos_gf_normal (ED_GF_NORMAL.f90:915-975) is host arithmetic on an nlanc x nlanc tridiagonal problem:
    the product's version (T=0 / finite-T weight, Matsubara and real axis) against the oracle's restatement."""
    from cdmft_lanc_ed_b200 import ed_hamiltonian as E
    rng = np.random.default_rng(3)
    n = 25
    a, b = rng.normal(size=n), np.abs(rng.normal(size=n)) + 0.1
    b[0] = 0.0
    wm = np.pi / 30.0 * (2 * np.arange(1, 33) - 1)
    wr = np.linspace(-4, 4, 41)
    for finite_t, beta, ei, egs in [(0, 0.0, -1.3, -1.3), (1, 30.0, -1.1, -1.3), (1, 30.0, 9.0, -1.3)]:
        for isign in (1, -1):
            gm, gr = np.zeros(wm.size, dtype=np.complex128), np.zeros(wr.size, dtype=np.complex128)
            om, orr = gm.copy(), gr.copy()
            p, w = E.add_to_lanczos_gf_normal_full(0.7 + 0.1j, ei, egs, finite_t, beta, a, b, isign, 1.3, wm, gm, wr, 0.01, gr)
            op, ow = oracle_lib.add_to_lanczos_gf_full(0.7 + 0.1j, ei, egs, finite_t, beta, a, b, isign, 1.3, wm, om, wr, 0.01, orr)
            assert np.abs(p - op).max() < 1e-12 and np.abs(w - ow).max() < 1e-13
            assert np.abs(gm - om).max() <= 1e-12 * max(np.abs(om).max(), 1e-30)
            assert np.abs(gr - orr).max() <= 1e-12 * max(np.abs(orr).max(), 1e-30)
            if finite_t and ei > 5:
                assert np.all(gm == 0) and np.all(gr == 0)  # beta*(Ei-Egs) >= 200: the state carries no weight
    # the T=0 entry is the same routine
    g1, g2 = np.zeros(wm.size, dtype=np.complex128), np.zeros(wm.size, dtype=np.complex128)
    E.add_to_lanczos_gf_normal(0.5, -1.0, a, b, 1, 1.0, wm, g1)
    E.add_to_lanczos_gf_normal_full(0.5, -1.0, -1.0, 0, 0.0, a, b, 1, 1.0, wm, g2, np.zeros(0), 0.0, np.zeros(0, dtype=np.complex128))
    assert np.array_equal(g1, g2)


def test_local_energy_host_formulas_against_oracle(oracle_lib):
    """The occupation-dependent pieces of lanc_local_energy as contractions of the impurity-configuration weights
    (host logic of the product; the weight table is built with numpy here) against the oracle's state loop."""
    from cdmft_lanc_ed_b200 import ed_hamiltonian as E
    from cdmft_lanc_ed_b200 import models
    for mdl, (nup, ndw) in [(models.hm2x2(1), (4, 3)), (models.bhz2(1, kanamori=True), (4, 4)),
                            (models.random_model(1, 3, 1, seed=14, hfmode=False), (2, 3)), (models.random_model(2, 2, 1, nspin=2, seed=12), (3, 3))]:
        ns, nimp = mdl.ns, mdl.nlat * mdl.norb
        isec = models.get_sector(ns, nup, ndw)
        mu_of = oracle_lib.sector_map(ns, nup) & ((1 << nimp) - 1)
        md_of = oracle_lib.sector_map(ns, ndw) & ((1 << nimp) - 1)
        rng = np.random.default_rng(5)
        vec = rng.normal(size=mu_of.size * md_of.size) + 1j * rng.normal(size=mu_of.size * md_of.size)
        vec /= np.linalg.norm(vec)
        W = np.zeros((1 << nimp, 1 << nimp))
        np.add.at(W, (np.tile(mu_of, md_of.size), np.repeat(md_of, mu_of.size)), np.abs(vec) ** 2)
        got = E.local_energy_from_weights(W, mdl, 0.3)
        ref = oracle_lib.Oracle(mdl).lanc_local_energy(isec, vec, 0.3)
        for k in ("Epot", "Ehartree", "Dust", "Dund"):
            assert abs(got[k] - ref[k]) < 1e-12 * max(1.0, abs(ref[k])), (mdl.name, k)
