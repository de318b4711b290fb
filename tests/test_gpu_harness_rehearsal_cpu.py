"""Harness rehearsal, NOT a parity test: the bodies of a few `-m gpu` tests are run here against a FAKE device whose entry
points are served by the CPU oracle, so that mistakes in the test code itself (argument order, shapes, tolerances that
two runs of one implementation cannot meet) surface in the CPU suite instead of on the GPU box.  Oracle against oracle
says nothing about the product; the real comparison happens when the same bodies run with the CUDA library."""
from math import comb

import numpy as np

from cdmft_lanc_ed_b200 import ed_hamiltonian as E
from cdmft_lanc_ed_b200 import ed_diag, gf_normal, observables
from oracle import edo
import tests.test_gpu_parity as T


class FakeDevice:
    """the subset of cdmft_lanc_ed_b200.ed_hamiltonian the rehearsed tests call"""
    EdB200Error = E.EdB200Error
    add_to_lanczos_gf_normal = staticmethod(E.add_to_lanczos_gf_normal)            # host arithmetic of the product
    add_to_lanczos_gf_normal_full = staticmethod(E.add_to_lanczos_gf_normal_full)

    def ed_set_model(self, mdl):
        self.m, self.o, self.isec = mdl, edo.Oracle(mdl), None

    def build_Hv_sector(self, isec, sparse=True):
        self.o.build_hv_sector(isec, edo.SPARSE_SERIAL)
        self.isec = isec
        return self.o.dim

    def delete_Hv_sector(self):
        self.o.delete_hv_sector()
        self.isec = None

    def getDim(self, isec):
        ns = self.m.ns
        nup, ndw = (isec - 1) // (ns + 1), (isec - 1) % (ns + 1)
        return comb(ns, nup) * comb(ns, ndw), comb(ns, nup), comb(ns, ndw)

    def sp_lanc_eigh(self, vec, nitermax=512, threshold=1e-18, ncheck=10):
        e0, v, nit, a, b = self.o.lanc_eigh(nitermax, threshold, ncheck, v0=vec if np.any(vec) else None)
        vec[:] = v
        return e0, nit, a, b

    def sp_lanc_tridiag(self, v, nlanc, threshold=1e-12):
        return self.o.lanc_tridiag(v, nlanc)

    def apply_op(self, isec, iop, ispin, pos, coef, state):
        return edo.apply_op(self.m.ns, isec, iop, ispin, pos, coef, state)

    def sp_eigh_device(self, neig, nblock=None, nitermax=512, tol=1e-18, basis=None):
        return E.eigh_logic_host(self.o.hxv, self.o.dim, neig, nblock=nblock, nitermax=nitermax, tol=tol)

    def density_matrix_impurity(self, vec, nlat, norb, nspin, peso=1.0):
        return self.o.density_matrix_impurity(self.isec, vec, peso)

    def lanc_observables(self, vec, nlat, norb, peso=1.0):
        return edo.lanc_observables(self.m.ns, nlat, norb, self.isec, vec, peso)

    def lanc_local_energy(self, vec, mdl, peso=1.0):
        return self.o.lanc_local_energy(self.isec, vec, peso)

    def build_Hmat(self):
        isec = self.isec
        self.o.delete_hv_sector()  # the oracle assembles the dense matrix without an active sector
        h = self.o.dense_hmat(isec)
        self.o.build_hv_sector(isec, edo.SPARSE_SERIAL)
        return h


def test_rehearse_u0_and_state_list_bodies(monkeypatch, oracle_lib):
    fake = FakeDevice()
    T.test_u0_gimp_equals_the_references_g0and_bath(fake, "models.hm2x2(1)")
    T.test_u0_observables_equal_the_slater_determinant(fake, "models.bhz2(1)")
    T.test_hubbard_dimer_closed_form(fake, 2.0, 0.25)
    monkeypatch.setattr(gf_normal, "E", fake)
    monkeypatch.setattr(observables, "E", fake)
    T.test_gf_over_a_state_list_finite_temperature(fake, oracle_lib)
    monkeypatch.setattr(ed_diag, "E", fake)
    T.test_ed_diag_sector_loop_on_the_device(fake, oracle_lib)
