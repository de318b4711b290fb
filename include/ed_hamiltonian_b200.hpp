// ed_hamiltonian_b200.hpp -- header-only C++ host mirror of the reference's ED_HAMILTONIAN interface on
// top of the C ABI (cdmft_b200.h).  The reference is compiled (Fortran) code whose toolchain is absent
// from the build image; this is the compiled-language host layer a C++ caller links against, with the
// reference's names, argument meaning and error behaviour (`stop "msg"` -> std::runtime_error):
//   build_Hv_sector / delete_Hv_sector / vecDim_Hv_sector     ED_HAMILTONIAN.f90:39-221
//   spHtimesV_p (procedure pointer, cc_sparse_HxV)            ED_VARS_GLOBAL.f90:72-78,146
//   sp_lanc_eigh / sp_lanc_tridiag call sites                 ED_DIAG.f90:176-184, ED_GF_NORMAL.f90:215
//   sp_eigh (default LANC_METHOD, (P)ARPACK)                  ED_DIAG.f90:150-170
#pragma once
#include <complex>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "cdmft_b200.h"

namespace ed_b200 {

using cplx = std::complex<double>;
using cc_sparse_HxV = void (*)(int Nloc, const cplx *v, cplx *Hv);  // ED_VARS_GLOBAL.f90:72-78

inline void check(int rc, const char *where) {
  if (rc != 0) throw std::runtime_error(std::string(where) + ": " + cdmft_b200_last_error());
}

// impHloc / Hbath_build(lambda) / bath V in Fortran order, plus the interaction scalars
struct Model {
  int Nlat = 1, Norb = 1, Nspin = 1, Nbath = 1;
  double Uloc[5] = {2, 0, 0, 0, 0};
  double Ust = 0, Jh = 0, Jx = 0, Jp = 0, xmu = 0;
  bool hfmode = true;
  std::vector<cplx> impHloc;  // [Nlat,Nlat,Nspin,Nspin,Norb,Norb]
  std::vector<cplx> Hbath;    // [Nlat,Nlat,Nspin,Nspin,Norb,Norb,Nbath]
  std::vector<double> Vbath;  // [Nlat*Nspin*Norb, Nbath]
};

inline cc_sparse_HxV &spHtimesV_p() {  // null when no sector is built (ED_MAIN.f90:231,280)
  static cc_sparse_HxV p = nullptr;
  return p;
}
inline void b200_HxV(int Nloc, const cplx *v, cplx *Hv) { check(cdmft_b200_hxv(Nloc, v, Hv), "spHtimesV_p"); }

inline void ed_init(int device = 0) { check(cdmft_b200_init(device), "ed_init"); }
inline void ed_finalize() { check(cdmft_b200_finalize(), "ed_finalize"); spHtimesV_p() = nullptr; }

inline void ed_set_model(const Model &m) {
  cdmft_b200_model c{};
  c.nlat = m.Nlat; c.norb = m.Norb; c.nspin = m.Nspin; c.nbath = m.Nbath;
  for (int i = 0; i < 5; i++) c.uloc[i] = m.Uloc[i];
  c.ust = m.Ust; c.jh = m.Jh; c.jx = m.Jx; c.jp = m.Jp; c.xmu = m.xmu;
  c.hfmode = m.hfmode ? 1 : 0;
  c.quirk_direct_bathdiag = 0;
  c.imphloc = reinterpret_cast<const double *>(m.impHloc.data());
  c.hbath = reinterpret_cast<const double *>(m.Hbath.data());
  c.vbath = m.Vbath.data();
  check(cdmft_b200_set_model(&c), "ed_set_model");
}

inline int get_Sector(int Nup, int Ndw) {  // ED_SETUP.f90:446-457
  int32_t ns = 0;
  check(cdmft_b200_get_ns(&ns), "get_Sector");
  return 1 + Nup * (ns + 1) + Ndw;
}

inline int64_t build_Hv_sector(int isector, bool ed_sparse_H = true) {
  int64_t nloc = 0;
  check(cdmft_b200_build_hv_sector(isector, ed_sparse_H ? CDMFT_B200_SPARSE : CDMFT_B200_DIRECT, &nloc), "build_Hv_sector");
  spHtimesV_p() = b200_HxV;
  return nloc;
}
// build_Hv_sector(isector, Hmat): the reference's optional dense matrix (ED_HAMILTONIAN.f90:123-127, caller ED_DIAG.f90:199),
// column-major [Dim, Dim]
inline int64_t build_Hv_sector(int isector, std::vector<cplx> &Hmat, bool ed_sparse_H = true) {
  const int64_t nloc = build_Hv_sector(isector, ed_sparse_H);
  int64_t du = 0, dd = 0, dim = 0;
  check(cdmft_b200_get_sector_dims(isector, &du, &dd, &dim), "build_Hv_sector(Hmat)");
  Hmat.assign((size_t)dim * (size_t)dim, cplx(0.0, 0.0));
  check(cdmft_b200_build_hmat(Hmat.data()), "build_Hv_sector(Hmat)");
  return nloc;
}
// scatter_vector_MPI / gather_vector_MPI (ED_SETUP.f90:575-668), root = master
inline void scatter_vector_MPI(const cplx *v, cplx *vloc) { check(cdmft_b200_scatter_vector(v, vloc, 0), "scatter_vector_MPI"); }
inline void gather_vector_MPI(const cplx *vloc, cplx *v) { check(cdmft_b200_gather_vector(vloc, v, 0), "gather_vector_MPI"); }
// hopping part of ed_Eknot in lanc_local_energy (ED_OBSERVABLES.f90:305-345)
inline double imp_kinetic(const std::vector<cplx> &vec) {
  double out[2] = {0, 0};
  check(cdmft_b200_imp_kinetic((int64_t)vec.size(), vec.data(), out), "lanc_local_energy");
  return out[0];
}
inline void delete_Hv_sector() {
  check(cdmft_b200_delete_hv_sector(), "delete_Hv_sector");
  spHtimesV_p() = nullptr;
}
inline int64_t vecDim_Hv_sector(int isector) {
  int64_t n = 0;
  check(cdmft_b200_vecdim_hv_sector(isector, &n), "vecDim_Hv_sector");
  return n;
}

// sp_lanc_eigh(spHtimesV_p, egs, vect, Nitermax, threshold): vect = start vector in / eigenvector out
inline int sp_lanc_eigh(double &egs, std::vector<cplx> &vect, int Nitermax, double threshold = 1e-18, int ncheck = 10) {
  int32_t niter = 0;
  check(cdmft_b200_lanczos_gs((int64_t)vect.size(), vect.data(), Nitermax, threshold, ncheck, &egs, &niter, nullptr, nullptr),
        "sp_lanc_eigh");
  return niter;
}
// sp_eigh(spHtimesV_p, eig_values, eig_basis, Nblock, Nitermax, tol) -- the default LANC_METHOD (ED_DIAG.f90:150-170):
// eig_values.size() = Neigen on entry; eig_basis receives [vecDim, Neigen] column-major.  Device-resident thick-restart
// Lanczos (cdmft_b200_eigh); returns the number of converged pairs (ARPACK's nconv).
inline int sp_eigh(std::vector<double> &eig_values, std::vector<cplx> &eig_basis, int64_t vecDim, int Nblock, int Nitermax,
                   double tol = 1e-18) {
  int32_t nconv = 0, nmv = 0;
  eig_basis.assign((size_t)vecDim * eig_values.size(), cplx(0.0, 0.0));
  check(cdmft_b200_eigh(vecDim, (int32_t)eig_values.size(), Nblock, Nitermax, tol, eig_values.data(), eig_basis.data(), &nconv, &nmv),
        "sp_eigh");
  return nconv;
}
// sp_lanc_tridiag(spHtimesV_p, vin, alanc, blanc)
inline int sp_lanc_tridiag(const std::vector<cplx> &vin, std::vector<double> &alanc, std::vector<double> &blanc,
                           double threshold = 1e-12) {
  int32_t nd = 0;
  blanc.resize(alanc.size());
  check(cdmft_b200_lanczos_tridiag((int64_t)vin.size(), vin.data(), (int32_t)alanc.size(), threshold, alanc.data(),
                                   blanc.data(), &nd),
        "sp_lanc_tridiag");
  return nd;
}

}  // namespace ed_b200
