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
