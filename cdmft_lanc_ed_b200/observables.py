"""Host orchestration of the impurity observables over a STATE LIST -- the `do istate=1,state_list%size` loops of
lanc_observables (ED_OBSERVABLES.f90:94-236), lanc_local_energy (:246-460) and density_matrix_impurity (:465-686):
every state enters with peso = exp(-beta (E_i - Egs)) / zeta_function at finite temperature, 1 / zeta_function at T = 0
(:133-135, :282-284, :501-503).  Per state one sector build; the O(Dim) work runs on the device behind the C ABI
(cdmft_b200_imp_weights, cdmft_b200_imp_kinetic, cdmft_b200_density_matrices).  Nothing here imports the test oracle."""
from __future__ import annotations

import numpy as np

from . import ed_hamiltonian as E
from .gf_normal import zeta_function


class _Device:
    """the calls this module makes, served by the C-ABI mirror"""

    def __init__(self, model):
        self.m = model

    def build(self, isector):
        return E.build_Hv_sector(isector)

    def delete(self):
        E.delete_Hv_sector()

    def observables(self, isector, vec, peso):
        return E.lanc_observables(vec, self.m.nlat, self.m.norb, peso)

    def local_energy(self, isector, vec, peso):
        return E.lanc_local_energy(vec, self.m, peso)

    def density_matrices(self, isector, vec, peso):
        return E.density_matrix_impurity(vec, self.m.nlat, self.m.norb, self.m.nspin, peso)


def observables_states(model, states, finite_t: bool = False, beta: float = 1000.0, density_matrices: bool = True, backend=None):
    """states: list of (isector, energy, vector) = the rows of `state_list`.  Returns a dict with the sums over the list of
    the local observables (dens, dens_up, dens_dw, docc, magz, s2tot, sz2, n2), the local energy pieces (Eknot, Epot, Ehartree,
    Dust, Dund) and -- density_matrices -- `cluster_density_matrix` [4^Nimp, 4^Nimp] and `single_particle_density_matrix`.
    The model must be set (ed_set_model), no sector may be active.  `backend`: object with build / delete / observables /
    local_energy / density_matrices (the device by default; the CPU tests pass an adapter over their checker)."""
    B = _Device(model) if backend is None else backend
    states = list(states)
    if not states:
        raise ValueError("observables_states: empty state list")
    egs = min(float(e) for _, e, _ in states)
    zeta = zeta_function([e for _, e, _ in states], finite_t, beta)
    out = {}

    def add(d):
        for k, v in d.items():
            out[k] = (out[k] + v) if k in out else (np.array(v, copy=True) if isinstance(v, np.ndarray) else v)

    for isector, e_i, vec in states:
        peso = (np.exp(-beta * (float(e_i) - egs)) if finite_t else 1.0) / zeta
        B.build(isector)
        try:
            add(B.observables(isector, vec, peso))
            add(B.local_energy(isector, vec, peso))
            if density_matrices:
                cdm, spdm = B.density_matrices(isector, vec, peso)
                add({"cluster_density_matrix": cdm, "single_particle_density_matrix": spdm})
        finally:
            B.delete()
    out["zeta_function"] = zeta
    return out
