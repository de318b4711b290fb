"""One impurity solve on the GPU through the host mirror, in the order the reference runs it (ED_MAIN.f90 ed_solve:
ed_diag -> observables -> Green's functions), for a 2x2 Hubbard cluster with a replica bath:

  1. ED_DIAG.f90:139-190   every sector of interest: build_Hv_sector -> sp_eigh (device-resident) -> delete_Hv_sector,
                           the lowest states form the state list (finite temperature: Boltzmann cut-off)
  2. ED_OBSERVABLES.f90    dens / docc / local energy / density matrices summed over the list
  3. ED_GF_NORMAL.f90      impurity Green's function (Matsubara + real axis) summed over the list

  python examples/impurity_solve.py [nbath=1] [beta=20]        (needs a B200; there is no CPU fallback)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cdmft_lanc_ed_b200 import ed_hamiltonian as E  # noqa: E402
from cdmft_lanc_ed_b200 import gf_normal, models, observables  # noqa: E402


def main(nbath=1, beta=20.0, neigen=2, cutoff=1e-9):
    mdl = models.hm2x2(nbath)
    ns, nimp = mdl.ns, mdl.nlat * mdl.norb
    E.ed_init(0)
    try:
        E.ed_set_model(mdl)
        # --- 1. ed_diag: a window of sectors around half filling
        states = []
        for nup in range(ns // 2 - 1, ns // 2 + 2):
            for ndw in range(ns // 2 - 1, ns // 2 + 2):
                isec = models.get_sector(ns, nup, ndw)
                E.build_Hv_sector(isec, True)
                w, z, info = E.sp_eigh_device(neigen, tol=1e-12)
                E.delete_Hv_sector()
                states += [(isec, float(w[k]), np.ascontiguousarray(z[:, k])) for k in range(neigen)]
        egs = min(e for _, e, _ in states)
        states = [s for s in states if np.exp(-beta * (s[1] - egs)) > cutoff]  # the reference's Boltzmann cut-off of the list
        print(f"Egs = {egs:.12f}, {len(states)} states kept, Z = {gf_normal.zeta_function([e for _, e, _ in states], True, beta):.6f}")
        # --- 2. observables
        obs = observables.observables_states(mdl, states, finite_t=True, beta=beta)
        print("dens =", np.asarray(obs["dens"]).ravel(), " docc =", np.asarray(obs["docc"]).ravel())
        print(f"Eknot = {obs['Eknot']:.10f}  Epot = {obs['Epot']:.10f}  Tr rho_imp = {np.trace(obs['cluster_density_matrix']).real:.12f}")
        # --- 3. Green's functions
        wm = np.pi / beta * (2 * np.arange(1, 65) - 1)
        wr = np.linspace(-4, 4, 201)
        G, Gr = gf_normal.build_gf_normal_states(nimp, states, wm, wr=wr, eps=0.05, finite_t=True, beta=beta)
        print("G_11(i w_0) =", G[0, 0, 0, 0], " G_12(i w_0) =", G[0, 0, 1, 0])
        print("spectral weight of A_11 on the grid =", float(-np.trapezoid(Gr[0, 0, 0].imag, wr) / np.pi))
    finally:
        E.ed_finalize()


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 1, float(sys.argv[2]) if len(sys.argv) > 2 else 20.0)
