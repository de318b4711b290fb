"""CPU check of the host mirror's sp_eigh (the reference's default LANC_METHOD: ARPACK in reverse communication around
the procedure pointer, ED_DIAG.f90:150-170): the same wrapper the GPU path uses, driven here by the ORACLE's mat-vec,
against a dense diagonalisation of the oracle's Hmat.  Several eigenpairs (the finite-temperature use), complex H."""
import numpy as np
import pytest

from cdmft_lanc_ed_b200 import ed_hamiltonian as E
from cdmft_lanc_ed_b200 import models
from oracle import edo


@pytest.mark.parametrize("case", [("hm2x2(1)", (4, 4), 3), ("bhz2(1)", (3, 3), 4), ("random_model(2, 2, 1, seed=8)", (3, 4), 2)])
def test_sp_eigh_wrapper_with_oracle_matvec(case):
    mdl = eval("models." + case[0])
    nup, ndw = case[1]
    neigen = case[2]
    isec = models.get_sector(mdl.ns, nup, ndw)
    orc = edo.Oracle(mdl)
    ref = np.linalg.eigvalsh(orc.dense_hmat(isec))  # build_Hv_sector(isector, Hmat): dense assembly, no active sector
    orc.build_hv_sector(isec, edo.SPARSE_SERIAL)
    n = orc.dim
    w, z = E.sp_eigh(neigen, nitermax=512, tol=1e-18, matvec=orc.hxv, n=n)
    assert np.abs(w - ref[:neigen]).max() < 1e-10 * max(1.0, np.abs(ref[:neigen]).max())
    for k in range(neigen):
        assert np.linalg.norm(orc.hxv(z[:, k]) - w[k] * z[:, k]) < 1e-8
    assert np.abs(z.conj().T @ z - np.eye(neigen)).max() < 1e-10
    orc.delete_hv_sector()


def test_sp_eigh_rejects_full_spectrum_requests():
    with pytest.raises(E.EdB200Error):
        E.sp_eigh(5, matvec=lambda x: x, n=5)


# ---- restart logic of the DEVICE-RESIDENT sp_eigh (cdmft_b200_eigh, csrc/trlan.h) through its host test hook ----------
def _check_pairs(hxv, w, z, ref, neigen, tol_e=1e-10):
    assert np.abs(w - ref[:neigen]).max() < tol_e * max(1.0, np.abs(ref[:neigen]).max())
    for k in range(neigen):
        assert np.linalg.norm(hxv(z[:, k]) - w[k] * z[:, k]) < 1e-8
    assert np.abs(z.conj().T @ z - np.eye(neigen)).max() < 1e-10


@pytest.mark.parametrize("case", [("hm2x2(1)", (4, 4), 1, None), ("hm2x2(1)", (3, 2), 3, 12), ("bhz2(1)", (3, 3), 4, None),
                                  ("random_model(2, 2, 1, seed=8)", (3, 4), 2, 3), ("bhz2(1, kanamori=True)", (3, 3), 3, 20)])
def test_thick_restart_logic_with_oracle_matvec(case):
    """nev lowest eigenpairs of real, complex and Kanamori sectors vs a dense diagonalisation of the oracle's Hmat;
    default Nblock (the reference's 10*max(Neigen,2)), a small one (many restarts) and the minimum nev + 1."""
    mdl = eval("models." + case[0])
    (nup, ndw), neigen, nblock = case[1], case[2], case[3]
    isec = models.get_sector(mdl.ns, nup, ndw)
    orc = edo.Oracle(mdl)
    ref = np.linalg.eigvalsh(orc.dense_hmat(isec))
    orc.build_hv_sector(isec, edo.SPARSE_SERIAL)
    # pairs inside a degenerate multiplet cannot be told apart by one Krylov sequence (ARPACK neither): only ask for
    # as many pairs as are separated from the next level
    while neigen > 1 and abs(ref[neigen] - ref[neigen - 1]) < 1e-8:
        neigen -= 1
    w, z, info = E.eigh_logic_host(orc.hxv, orc.dim, neigen, nblock=nblock, nitermax=2000, tol=1e-12)
    assert info["nconv"] == neigen, info
    mult = [np.sum(np.abs(ref - e) < 1e-8) for e in ref[:neigen]]
    if max(mult) == 1:
        _check_pairs(orc.hxv, w, z, ref, neigen)
    else:  # a degenerate level below the cut: every returned value must be an eigenvalue, every vector an eigenvector
        for k in range(neigen):
            assert np.abs(ref - w[k]).min() < 1e-9
            assert np.linalg.norm(orc.hxv(z[:, k]) - w[k] * z[:, k]) < 1e-8
    orc.delete_hv_sector()


def test_thick_restart_default_tolerance_terminates():
    """The reference passes lanc_tolerance = 1e-18 (ED_INPUT_VARS.f90:178): it must act as machine precision, not run forever."""
    mdl = models.hm2x2(1)
    isec = models.get_sector(mdl.ns, 3, 3)
    orc = edo.Oracle(mdl)
    ref = np.linalg.eigvalsh(orc.dense_hmat(isec))
    orc.build_hv_sector(isec, edo.SPARSE_SERIAL)
    w, z, info = E.eigh_logic_host(orc.hxv, orc.dim, 2, nitermax=512, tol=1e-18)
    assert info["nconv"] == 2 and info["nmatvec"] < 2000, info
    _check_pairs(orc.hxv, w, z, ref, 2, tol_e=1e-12)
    orc.delete_hv_sector()


def test_thick_restart_invariant_subspace_and_restart_cap():
    """A diagonal operator with 5 distinct levels: the Krylov space closes after 5 steps (beta = 0) and the driver must go
    on with fresh orthogonal directions; then Nitermax = 0 on a generic operator: one sweep, nconv < neigen, no failure."""
    n = 200
    d = np.repeat(np.array([-2.0, -1.0, 0.5, 1.0, 3.0]), n // 5)
    w, z, info = E.eigh_logic_host(lambda v: d * v, n, 3, nblock=12, nitermax=50, tol=1e-12)
    assert np.allclose(w, [-2.0, -2.0, -2.0], atol=1e-10) or np.allclose(w, [-2.0, -1.0, 0.5], atol=1e-10) or \
        all(np.abs(d - x).min() < 1e-10 for x in w), w
    for k in range(3):
        assert np.linalg.norm(d * z[:, k] - w[k] * z[:, k]) < 1e-8
    rng = np.random.default_rng(3)
    a = rng.normal(size=(300, 300)) + 1j * rng.normal(size=(300, 300))
    a = a + a.conj().T
    w, z, info = E.eigh_logic_host(lambda v: a @ v, 300, 2, nblock=6, nitermax=0, tol=1e-12)
    assert info["nconv"] < 2 and info["nmatvec"] == 6
    w, z, info = E.eigh_logic_host(lambda v: a @ v, 300, 2, nblock=24, nitermax=500, tol=1e-12)
    _check_pairs(lambda v: a @ v, w, z, np.linalg.eigvalsh(a), 2)


def test_thick_restart_rejects_tiny_spaces():
    with pytest.raises(E.EdB200Error):
        E.eigh_logic_host(lambda v: v, 3, 2)


def test_thick_restart_one_sided_spectrum():
    """A sector whose whole spectrum lies on one side of zero (|alfa| >> beta in every Lanczos step): the case in which one
    classical Gram-Schmidt pass per step lets the orthogonality error compound (an earlier version lost the basis after the
    first restart here and returned -128 for a spectrum in [-10.1, -3.0]).  Found by the ed_diag test; kept as a regression."""
    mdl = models.random_model(3, 1, 1, seed=5)
    isec = models.get_sector(mdl.ns, 5, 5)  # Dim 36
    orc = edo.Oracle(mdl)
    ref = np.linalg.eigvalsh(orc.dense_hmat(isec))
    assert ref[-1] < 0
    orc.build_hv_sector(isec, edo.SPARSE_SERIAL)
    for nblock in (20, 10, 35, 3):
        w, z, info = E.eigh_logic_host(orc.hxv, orc.dim, 2, nblock=nblock, nitermax=2000, tol=1e-13)
        assert info["nconv"] == 2, (nblock, info)
        _check_pairs(orc.hxv, w, z, ref, 2, tol_e=1e-11)
    orc.delete_hv_sector()
    # the same with a shifted dense operator far from zero: H - 1000
    rng = np.random.default_rng(11)
    a = rng.normal(size=(120, 120))
    a = (a + a.T) / 2 - 1000.0 * np.eye(120)
    w, z, info = E.eigh_logic_host(lambda v: a @ v, 120, 3, nblock=16, nitermax=2000, tol=1e-14)
    _check_pairs(lambda v: a @ v, w, z, np.linalg.eigvalsh(a), 3, tol_e=1e-12)


def test_thick_restart_randomised_campaign():
    """Random Hermitian operators with prescribed spectra (generic, one-sided with a tiny spread, heavily degenerate,
    near-degenerate lowest pair, quadratic ladder), random Neigen and Nblock: every converged pair is an eigenpair, the
    vectors are orthonormal, and with separated lowest levels the values are THE lowest."""
    rng = np.random.default_rng(2026)
    for trial in range(60):
        n = int(rng.choice([5, 8, 13, 36, 64, 100]))
        kind = int(rng.integers(0, 5))
        q, _ = np.linalg.qr(rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n)))
        d = [rng.normal(size=n), -1000 + rng.random(n), np.round(rng.normal(size=n) * 2) / 2,
             np.concatenate([[-5.0, -5.0 + 1e-7], rng.random(n - 2)]), np.arange(n, dtype=float) ** 2 * 1e-3 + 50][kind]
        a = (q * d) @ q.conj().T
        a = (a + a.conj().T) / 2
        ref = np.linalg.eigvalsh(a)
        nev = int(rng.integers(1, min(6, n - 2) + 1))
        ncv = int(rng.integers(min(nev + 2, n - 1), min(n - 1, 40) + 1))
        w, z, info = E.eigh_logic_host(lambda v: a @ v, n, nev, nblock=ncv, nitermax=3000, tol=1e-12)
        if info["nconv"] < nev:
            continue  # slow cases (Nblock = Neigen + 2 on a near-degenerate pair) end at the restart cap, like ARPACK's info = 1
        scale = max(1.0, np.abs(ref).max())
        for k in range(nev):
            assert np.linalg.norm(a @ z[:, k] - w[k] * z[:, k]) < 1e-7 * scale, (trial, n, kind, nev, ncv)
            assert np.abs(ref - w[k]).min() < 1e-8 * scale
        assert np.abs(z.conj().T @ z - np.eye(nev)).max() < 1e-8
        if np.min(np.diff(ref[: nev + 1])) > 1e-6 * scale:
            assert np.abs(w - ref[:nev]).max() < 1e-8 * scale, (trial, n, kind, nev, ncv)
