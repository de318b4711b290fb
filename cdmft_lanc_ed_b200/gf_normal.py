"""Host orchestration of the normal Green's function at T = 0 for one spin -- the flow of build_gf_normal /
lanc_build_gf_normal_diag / lanc_build_gf_normal_mix_chan4 (ED_GF_NORMAL.f90:38-106,123-306,531-903) on top of the C ABI,
with the channels BATCHED BY TARGET SECTOR: every c^+ channel (diagonal and mixed) lives in getCDGsector(isector), every c
channel in getCsector(isector), so each of the two sectors is built ONCE and serves all its channels (the reference
builds and deletes the sector once per channel: 2*Nimp + 4*Nimp*(Nimp-1) times).

    G_aa(z)  = <gs| c_a (z - (H - E0))^-1 c_a^+ |gs>  +  <gs| c_a^+ (z + (H - E0))^-1 c_a |gs>
    G_ab(z)  = 1/2 [ G_{a+b} - i G_{a+ib} - (1 - i)(G_aa + G_bb) ]            (the reference's "chan4" combination, :91-103)

The start vectors are made by cdmft_b200_apply_op, the continued fractions by cdmft_b200_lanczos_tridiag (device), the
pole / weight accumulation by cdmft_b200_add_to_lanczos_gf_full (host).  Nothing here imports the test oracle."""
from __future__ import annotations

import numpy as np

from . import ed_hamiltonian as E

LANC_NGFITER = 200  # ED_INPUT_VARS.f90:177


def build_gf_normal(nimp: int, isector: int, e0: float, gsvec: np.ndarray, wm, wr=None, eps: float = 1e-2, ispin: int = 1,
                    nlanc_max: int = LANC_NGFITER, zeta: float = 1.0, pairs=None):
    """Returns (Gmats [Nimp, Nimp, Lmats], Greal [Nimp, Nimp, Lreal] or None) for the non-degenerate ground state
    (e0, gsvec) of sector `isector`; a, b are imp_state_index positions (1-based in the reference, 0-based here in the
    array indices).  pairs: ordered off-diagonal pairs to compute (default: all); the model must be set (ed_set_model),
    no sector may be active."""
    wm = np.ascontiguousarray(wm, dtype=np.float64)
    wr = np.zeros(0) if wr is None else np.ascontiguousarray(wr, dtype=np.float64)
    if pairs is None:
        pairs = [(a, b) for a in range(1, nimp + 1) for b in range(1, nimp + 1) if a != b]
    # channel = (accumulator key, iop, positions, coefficients, prefactor of the weight)
    chans = []
    for a in range(1, nimp + 1):
        chans.append((("d", a), +1, [a], [1.0], 1.0))
        chans.append((("d", a), -1, [a], [1.0], 1.0))
    for (a, b) in pairs:
        chans.append((("m", a, b), +1, [a, b], [1.0, 1.0], 1.0))
        chans.append((("m", a, b), -1, [a, b], [1.0, 1.0], 1.0))
        chans.append((("m", a, b), +1, [a, b], [1.0, 1.0j], -1.0j))
        chans.append((("m", a, b), -1, [a, b], [1.0, -1.0j], -1.0j))
    acc_m = {c[0]: np.zeros(wm.size, dtype=np.complex128) for c in chans}
    acc_r = {c[0]: np.zeros(wr.size, dtype=np.complex128) for c in chans}
    builds = 0
    for iop in (+1, -1):  # one target sector per operator kind
        todo = []
        jsec = 0
        for key, op, pos, coef, fac in chans:
            if op != iop:
                continue
            js, vv = E.apply_op(isector, iop, ispin, pos, coef, gsvec)
            if js == 0:
                continue
            jsec = js
            todo.append((key, vv, fac))
        if not todo:
            continue
        n = E.build_Hv_sector(jsec)
        builds += 1
        try:
            jdim = E.getDim(jsec)[0]
            for key, vv, fac in todo:
                norm2 = float(np.vdot(vv, vv).real)
                if norm2 == 0.0:
                    continue
                assert vv.size == n
                nlanc = int(min(jdim, nlanc_max))
                nd, a_, b_ = E.sp_lanc_tridiag(vv / np.sqrt(norm2), nlanc)
                E.add_to_lanczos_gf_normal_full(fac * norm2, e0, e0, False, 0.0, a_[:nd], b_[:nd], iop, zeta, wm, acc_m[key], wr, eps,
                                                acc_r[key])
        finally:
            E.delete_Hv_sector()

    def combine(acc, nw):
        G = np.zeros((nimp, nimp, nw), dtype=np.complex128)
        for a in range(1, nimp + 1):
            G[a - 1, a - 1] = acc[("d", a)]
        for (a, b) in pairs:  # non-diagonal trick, ED_GF_NORMAL.f90:91-103
            G[a - 1, b - 1] = 0.5 * (acc[("m", a, b)] - (1 - 1j) * acc[("d", a)] - (1 - 1j) * acc[("d", b)])
        return G

    out = combine(acc_m, wm.size), (combine(acc_r, wr.size) if wr.size else None)
    build_gf_normal.last_sector_builds = builds
    return out


def zeta_function(energies, finite_t: bool, beta: float) -> float:
    """The partition sum over the state list (ED_DIAG.f90:355-366): sum_i exp(-beta (E_i - Egs)) at finite temperature, the
    number of states (= ground-state degeneracy) at T = 0."""
    e = np.asarray(energies, dtype=float)
    return float(np.exp(-beta * (e - e.min())).sum()) if finite_t else float(e.size)


def build_gf_normal_states(nimp: int, states, wm, wr=None, eps: float = 1e-2, nspin: int = 1, finite_t: bool = False, beta: float = 1000.0,
                           symmetric: bool = False, nlanc_max: int = LANC_NGFITER, backend=None):
    """build_gf_normal over a whole STATE LIST (ED_GF_NORMAL.f90:38-106: `do istate=1,state_list%size` inside every
    lanc_build_gf_normal_* routine) -- what the reference runs after `sp_eigh` has filled the list at finite temperature, or
    with a degenerate ground state at T = 0.

    states: list of (isector, energy, vector) -- the rows of `state_list` (es_return_sector / _energy / _cvector); every
    state contributes its channels with the weight exp(-beta (E_i - Egs)) / zeta (finite_t) or 1 / zeta (T = 0), `Egs` the
    list minimum and `zeta` = zeta_function (add_to_lanczos_gf_normal :930-936).  nspin: the spin-diagonal blocks computed
    (ispin = 1..nspin; spin-down operators for ispin = 2).  symmetric = ed_gf_symmetric: two mixed channels per pair
    (lanc_build_gf_normal_mix_chan2, :314-530) and the combination with chan4 = 0 instead of four.
    Channels are batched by target sector PER STATE: two sector builds per (state, spin).
    Returns (Gmats [nspin, Nimp, Nimp, Lmats], Greal [nspin, Nimp, Nimp, Lreal] or None).
    `backend`: the module providing apply_op / build_Hv_sector / getDim / sp_lanc_tridiag / add_to_lanczos_gf_normal_full /
    delete_Hv_sector -- the C-ABI mirror by default; the CPU tests pass an adapter over their checker to exercise this flow."""
    B = E if backend is None else backend
    wm = np.ascontiguousarray(wm, dtype=np.float64)
    wr = np.zeros(0) if wr is None else np.ascontiguousarray(wr, dtype=np.float64)
    states = list(states)
    if not states:
        raise ValueError("build_gf_normal_states: empty state list")
    egs = min(float(e) for _, e, _ in states)
    zeta = zeta_function([e for _, e, _ in states], finite_t, beta)
    pairs = [(a, b) for a in range(1, nimp + 1) for b in range(1, nimp + 1) if a != b]
    chans = []
    for a in range(1, nimp + 1):
        chans.append((("d", a), +1, [a], [1.0], 1.0))
        chans.append((("d", a), -1, [a], [1.0], 1.0))
    for (a, b) in pairs:
        chans.append((("m", a, b), +1, [a, b], [1.0, 1.0], 1.0))
        chans.append((("m", a, b), -1, [a, b], [1.0, 1.0], 1.0))
        if not symmetric:
            chans.append((("m", a, b), +1, [a, b], [1.0, 1.0j], -1.0j))
            chans.append((("m", a, b), -1, [a, b], [1.0, -1.0j], -1.0j))
    gm = np.zeros((nspin, nimp, nimp, wm.size), dtype=np.complex128)
    gr = np.zeros((nspin, nimp, nimp, wr.size), dtype=np.complex128)
    builds = 0
    for ispin in range(1, nspin + 1):
        acc_m = {c[0]: np.zeros(wm.size, dtype=np.complex128) for c in chans}
        acc_r = {c[0]: np.zeros(wr.size, dtype=np.complex128) for c in chans}
        for isector, e_i, vec in states:
            if finite_t and beta * (float(e_i) - egs) >= 200.0:
                continue  # the reference's weight is exactly zero there (:932-938)
            for iop in (+1, -1):
                todo, jsec = [], 0
                for key, op, pos, coef, fac in chans:
                    if op != iop:
                        continue
                    js, vv = B.apply_op(isector, iop, ispin, pos, coef, vec)
                    if js == 0:
                        continue
                    jsec = js
                    todo.append((key, vv, fac))
                if not todo:
                    continue
                n = B.build_Hv_sector(jsec)
                builds += 1
                try:
                    jdim = B.getDim(jsec)[0]
                    for key, vv, fac in todo:
                        norm2 = float(np.vdot(vv, vv).real)
                        if norm2 == 0.0:
                            continue
                        assert vv.size == n
                        nd, a_, b_ = B.sp_lanc_tridiag(vv / np.sqrt(norm2), int(min(jdim, nlanc_max)))
                        B.add_to_lanczos_gf_normal_full(fac * norm2, float(e_i), egs, finite_t, beta, a_[:nd], b_[:nd], iop, zeta, wm,
                                                        acc_m[key], wr, eps, acc_r[key])
                finally:
                    B.delete_Hv_sector()
        chan4 = 0.0 if symmetric else 1.0
        for acc, G in ((acc_m, gm), (acc_r, gr)):
            for a in range(1, nimp + 1):
                G[ispin - 1, a - 1, a - 1] = acc[("d", a)]
            for (a, b) in pairs:  # non-diagonal trick, ED_GF_NORMAL.f90:91-103
                G[ispin - 1, a - 1, b - 1] = 0.5 * (acc[("m", a, b)] - (1 - chan4 * 1j) * acc[("d", a)] - (1 - chan4 * 1j) * acc[("d", b)])
    build_gf_normal_states.last_sector_builds = builds
    return gm, (gr if wr.size else None)
