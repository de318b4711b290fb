"""Host orchestration of the diagonalisation step -- the CALLER of the hot path: ed_diag_c / ed_post_diag of ED_DIAG.f90
(:52-260, :337-470) restated over the C-ABI mirror.  Per sector: build_Hv_sector -> the Krylov solver (sp_eigh, the default
LANC_METHOD, device-resident; or sp_lanc_eigh) or, for small sectors, the dense Hmat + LAPACK (:194-219) -> delete_Hv_sector;
the eigenpairs enter the STATE LIST by the reference's rules (finite temperature: the lanc_nstates_total lowest states, then
the Boltzmann cut-off and the per-sector Neigen adaptation of ed_post_diag; T = 0: the ground states within gs_threshold).
The list is what observables.observables_states and gf_normal.build_gf_normal_states consume.
Twin sectors (ed_twin) and the restart files are not mirrored.  Nothing here imports the test oracle."""
from __future__ import annotations

from dataclasses import dataclass, field
from math import comb

import numpy as np

from . import ed_hamiltonian as E


@dataclass
class DiagParams:
    """the input variables the step reads (ED_INPUT_VARS.f90:160-182), with the reference's defaults"""
    finite_t: bool = False
    beta: float = 1000.0
    lanc_method: str = "arpack"          # "arpack" (default: sp_eigh) or "lanczos" (sp_lanc_eigh, Neigen = 1)
    lanc_nstates_sector: int = 2
    lanc_nstates_total: int = 2
    lanc_nstates_step: int = 2
    lanc_ncv_factor: int = 10
    lanc_ncv_add: int = 0
    lanc_niter: int = 512
    lanc_tolerance: float = 1e-18
    lanc_dim_threshold: int = 1024
    gs_threshold: float = 1e-9
    cutoff: float = 1e-9
    neigen_sector: dict = field(default_factory=dict)   # isector -> Neigen; missing: lanc_nstates_sector (ED_SETUP.f90 init)


class _Device:
    def build(self, isector):
        return E.build_Hv_sector(isector)

    def delete(self):
        E.delete_Hv_sector()

    def dense_hmat(self):
        return E.build_Hmat()

    def sp_eigh(self, neigen, nblock, nitermax, tol):
        w, z, info = E.sp_eigh_device(neigen, nblock=nblock, nitermax=nitermax, tol=tol)
        return w, z

    def sp_lanc_eigh(self, n, nitermax, tol):
        vec = np.zeros(n, dtype=np.complex128)
        e0, _, _, _ = E.sp_lanc_eigh(vec, nitermax, tol)
        return np.array([e0]), vec.reshape(-1, 1)


def sector_dim(ns: int, isector: int) -> int:
    nup, ndw = (isector - 1) // (ns + 1), (isector - 1) % (ns + 1)
    return comb(ns, nup) * comb(ns, ndw)


def ed_diag(ns: int, params: DiagParams, sectors=None, nranks: int = 1, backend=None):
    """The sector loop of ed_diag_c.  sectors: iterable of sector ids (default: all (Ns+1)^2, `sectors_mask` all true).
    Returns the state list [(isector, energy, vector)] sorted by energy (es_add_state keeps it sorted)."""
    B = _Device() if backend is None else backend
    p = params
    states = []
    oldzero = 1000.0
    for isector in (range(1, (ns + 1) ** 2 + 1) if sectors is None else sectors):
        dim = sector_dim(ns, isector)
        if p.lanc_method == "lanczos":
            neigen, nitermax, nblock = 1, min(dim, p.lanc_niter), 1
        else:  # ED_DIAG.f90:94-97
            neigen = min(dim, p.neigen_sector.get(isector, p.lanc_nstates_sector))
            nitermax = min(dim, p.lanc_niter)
            nblock = min(dim, p.lanc_ncv_factor * max(neigen, p.lanc_nstates_sector) + p.lanc_ncv_add)
        lanc_solve = not (neigen == dim or dim <= max(p.lanc_dim_threshold, nranks))  # :104-106
        n = B.build(isector)
        try:
            if lanc_solve:
                w, z = B.sp_lanc_eigh(n, nitermax, p.lanc_tolerance) if p.lanc_method == "lanczos" else B.sp_eigh(neigen, nblock, nitermax, p.lanc_tolerance)
            else:  # LAPACK branch (:194-219): build_Hv_sector(isector, Hmat); eigh; the Neigen lowest columns
                wall, zall = np.linalg.eigh(B.dense_hmat())
                w, z = wall[:neigen], zall[:, :neigen]
        finally:
            B.delete()
        for i in range(neigen):
            e_i, v_i = float(w[i]), np.ascontiguousarray(z[:, i])
            if p.finite_t:  # es_add_state(..., size = lanc_nstates_total): sorted insert, the list never grows past `size`
                states.append((isector, e_i, v_i))
                states.sort(key=lambda s: s[1])
                del states[p.lanc_nstates_total:]
            elif e_i < oldzero - 10.0 * p.gs_threshold:  # a new ground state: the list starts over (:233-237)
                oldzero = e_i
                states = [(isector, e_i, v_i)]
            elif abs(e_i - oldzero) <= p.gs_threshold:   # degenerate with it (:238-241)
                oldzero = min(oldzero, e_i)
                states.append((isector, e_i, v_i))
                states.sort(key=lambda s: s[1])
    return states


def ed_post_diag(ns: int, states, params: DiagParams):
    """ed_post_diag (:337-470): zeta_function, the ground-state degeneracy, and -- finite temperature -- the adaptation of
    neigen_sector / lanc_nstates_total for the NEXT call and the trimming of the list at exp(-beta (E - Egs)) <= cutoff.
    Returns (states, zeta_function, numgs); `params` is updated in place like the reference's module variables."""
    p = params
    states = list(states)
    egs = states[0][1]
    zeta = float(sum(np.exp(-p.beta * (e - egs)) for _, e, _ in states)) if p.finite_t else float(len(states))
    numgs = sum(1 for _, e, _ in states if abs(e - egs) <= p.gs_threshold)  # es_return_gs_degeneracy
    if p.finite_t:
        listed = [s[0] for s in states]
        for isector in range(1, (ns + 1) ** 2 + 1):  # :432-441
            cnt = listed.count(isector)
            ne = p.neigen_sector.get(isector, p.lanc_nstates_sector) + (1 if cnt else -1)
            if ne > cnt:
                ne = cnt + 1
            if ne <= 0:
                ne = 1
            p.neigen_sector[isector] = ne
        ec = states[-1][1]
        if np.exp(-p.beta * (ec - egs)) > p.cutoff:   # not enough states for the requested accuracy (:451-453)
            p.lanc_nstates_total += p.lanc_nstates_step
        else:                                          # trim (:455-469)
            while np.exp(-p.beta * (states[-1][1] - egs)) <= p.cutoff:
                states.pop()
            p.lanc_nstates_total = max(len(states), p.lanc_nstates_step) + p.lanc_nstates_step
    return states, zeta, numgs
