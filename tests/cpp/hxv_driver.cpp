// C++ driver over the header-only host mirror: 2x2 Hubbard plaquette, Nbath=1 (BASELINE config K1),
// sector (4,4): one H x v on a deterministic vector, the Lanczos ground-state energy and 10 tridiagonal
// coefficients are printed as JSON; tests/test_gpu_cpp_host.py compares them with the oracle.
#include <cmath>
#include <cstdio>

#include "ed_hamiltonian_b200.hpp"

using namespace ed_b200;

int main() {
  try {
    Model m;
    m.Nlat = 4; m.Norb = 1; m.Nspin = 1; m.Nbath = 1;
    const double ts = 0.25;
    m.impHloc.assign(16, cplx(0, 0));
    m.Hbath.assign(16, cplx(0, 0));
    const int bonds[4][2] = {{0, 1}, {2, 3}, {0, 2}, {1, 3}};  // drivers/cdn_hm_2dsquare.f90:221-259
    for (auto &b : bonds) {
      m.impHloc[b[0] + 4 * b[1]] = m.impHloc[b[1] + 4 * b[0]] = -ts;
      m.Hbath[b[0] + 4 * b[1]] = m.Hbath[b[1] + 4 * b[0]] = ts;  // lambda2 * abs(Hloc), onsite 0
    }
    m.Vbath.assign(4, 1.0);
    ed_init(0);
    ed_set_model(m);
    const int isector = get_Sector(4, 4);
    if (spHtimesV_p() != nullptr) return 2;
    const int64_t n = build_Hv_sector(isector);
    if (n != vecDim_Hv_sector(isector) || n != 4900) return 3;
    std::vector<cplx> v(n), hv(n);
    for (int64_t i = 0; i < n; i++) v[i] = cplx(std::sin(0.37 * (i + 1)), std::cos(0.11 * (i + 1)));
    spHtimesV_p()((int)n, v.data(), hv.data());
    std::vector<double> a(10), b;
    sp_lanc_tridiag(v, a, b);
    std::vector<cplx> gs(n, cplx(0, 0));
    double e0 = 0;
    int nit = sp_lanc_eigh(e0, gs, 512, 1e-14);
    // the default LANC_METHOD of ED_DIAG.f90:150-170: sp_eigh with Neigen = 2, Nblock = 20
    std::vector<double> evals(2, 0.0);
    std::vector<cplx> ebasis;
    const int nconv = sp_eigh(evals, ebasis, n, 20, 512, 1e-13);
    double eres = 0;
    for (int k = 0; k < 2; k++) {
      spHtimesV_p()((int)n, ebasis.data() + (size_t)k * n, hv.data());
      double r2 = 0;
      for (int64_t i = 0; i < n; i++) r2 += std::norm(hv[i] - evals[k] * ebasis[(size_t)k * n + i]);
      eres = std::fmax(eres, std::sqrt(r2));
    }
    spHtimesV_p()((int)n, v.data(), hv.data());
    delete_Hv_sector();
    // the LAPACK branch of ED_DIAG.f90:199: build_Hv_sector(isector, Hmat) on a small sector; <E0> hopping part
    const int ismall = get_Sector(2, 3);
    std::vector<cplx> Hmat;
    const int64_t ns_ = build_Hv_sector(ismall, Hmat);
    std::vector<cplx> w(ns_), hw(ns_), sw(ns_), gw(ns_);
    for (int64_t i = 0; i < ns_; i++) w[i] = cplx(std::cos(0.23 * (i + 1)), std::sin(0.31 * (i + 1)));
    spHtimesV_p()((int)ns_, w.data(), hw.data());
    double hdiff = 0, htrace = 0;
    for (int64_t i = 0; i < ns_; i++) {
      cplx acc(0, 0);
      for (int64_t j = 0; j < ns_; j++) acc += Hmat[i + j * ns_] * w[j];
      hdiff = std::fmax(hdiff, std::abs(acc - hw[i]));
      htrace += Hmat[i + i * ns_].real();
    }
    const double ekin = imp_kinetic(w);
    scatter_vector_MPI(w.data(), sw.data());
    gather_vector_MPI(sw.data(), gw.data());
    bool moved = true;
    for (int64_t i = 0; i < ns_; i++) moved = moved && gw[i] == w[i];
    delete_Hv_sector();
    bool threw = false;
    try { b200_HxV((int)n, v.data(), hv.data()); } catch (const std::runtime_error &) { threw = true; }
    ed_finalize();
    std::printf("{\"sp_eigh\": [%.15e, %.15e], \"sp_eigh_nconv\": %d, \"sp_eigh_residual\": %.3e, ", evals[0], evals[1], nconv, eres);
    std::printf("\"n\": %lld, \"e0\": %.15e, \"niter\": %d, \"threw_after_delete\": %s, \"hmat_n\": %lld, \"hmat_maxdiff\": %.3e, "
                "\"hmat_trace\": %.15e, \"imp_kinetic\": %.15e, \"scatter_gather_ok\": %s, \"hv\": [",
                (long long)n, e0, nit, threw ? "true" : "false", (long long)ns_, hdiff, htrace, ekin, moved ? "true" : "false");
    for (int i = 0; i < 8; i++) std::printf("%s[%.17e, %.17e]", i ? ", " : "", hv[i * 601].real(), hv[i * 601].imag());
    std::printf("], \"alanc\": [");
    for (int i = 0; i < 10; i++) std::printf("%s%.17e", i ? ", " : "", a[i]);
    std::printf("], \"blanc\": [");
    for (int i = 0; i < 10; i++) std::printf("%s%.17e", i ? ", " : "", b[i]);
    std::printf("]}\n");
    return 0;
  } catch (const std::exception &e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
}
