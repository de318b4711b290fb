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
