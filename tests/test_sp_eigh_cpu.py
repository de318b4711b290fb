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
