// trlan.h -- restart logic of cdmft_b200_eigh: the device-resident replacement of SciFortran's
//   sp_eigh(MatVec, eig_values, eig_basis, Nblock, Nitermax, tol)          caller ED_DIAG.f90:150-170
// (reverse-communication (P)ARPACK, which = 'SA', nev = Neigen, ncv = Nblock; SciFortran is an un-vendored dependency,
// SURVEY.md App. B).  ARPACK's implicitly restarted Lanczos and the thick-restart Lanczos used here (Wu & Simon, SIAM J.
// Matrix Anal. Appl. 22 (2000): "Thick-restart Lanczos method for large symmetric eigenvalue problems") span the same
// Krylov subspaces for a Hermitian operator; thick restart only needs basis rotations (one streaming pass over the
// basis) instead of the QR sweeps on the basis, which is what a device-resident basis wants.
//
// Host-only C++ (no CUDA): the vectors live behind a backend -- device slots (lanczos.cu) or host arrays (the CPU test
// hook cdmft_b200_eigh_logic_host).  Backend concept:
//   int rand_w(uint64_t seed)                          w = counter-based pseudo-random vector (function of the GLOBAL index)
//   int orth_w(int nv, complex *h, double *nrm2)       h = V[0..nv)^H w (all-reduced);  w -= V h;  nrm2 = |w|^2 afterwards
//   int store_w(int j, double scale)                   V_j = scale * w
//   int matvec(int j)                                  w = H V_j
//   int rotate(int m, int k, const double *Y)          V[0..k) = V[0..m) Y,  Y column-major m x k (in place)
//   int move(int dst, int src)                         V_dst = V_src
// every call returns 0 or an error code, which is passed through.
#pragma once
#include <stdint.h>
#include <algorithm>
#include <cmath>
#include <complex>
#include <vector>

namespace cb {

constexpr int kTrlMaxNcv = 64;  // basis vectors the rotation kernel keeps per thread

// counter-based uniform number in (-1, 1): splitmix64 of (seed, global index) -- the same start vector whatever the sharding
#if defined(__CUDACC__)
__host__ __device__
#endif
inline double trl_rand(uint64_t seed, uint64_t idx) {
  uint64_t z = (idx + 1) * 0x9E3779B97F4A7C15ull + seed * 0xD1B54A32D192ED03ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return ((double)(z >> 11) + 0.5) * (2.0 / 9007199254740992.0) - 1.0;
}

// real symmetric eigenproblem by cyclic Jacobi rotations (the projected matrix is at most 64 x 64: diagonal + arrow +
// tridiagonal tail).  A row-major n x n (a copy is worked on); w ascending, Z row-major: Z[r*n + c] = component r of vector c.
inline void jacobi_eigh(int n, std::vector<double> A, std::vector<double> &w, std::vector<double> &Z) {
  Z.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; i++) Z[(size_t)i * n + i] = 1.0;
  auto a = [&](int r, int c) -> double & { return A[(size_t)r * n + c]; };
  for (int sweep = 0; sweep < 100; sweep++) {
    bool rotated = false;
    for (int p = 0; p < n - 1; p++)
      for (int q = p + 1; q < n; q++) {
        const double apq = a(p, q);
        if (apq == 0.0) continue;
        // negligible against both diagonal entries: drop it (relative criterion -> small eigenvalues stay accurate)
        if (std::fabs(apq) <= 1e-3 * 2.220446049250313e-16 * std::sqrt(std::fabs(a(p, p) * a(q, q))) && sweep > 3) {
          a(p, q) = a(q, p) = 0.0;
          continue;
        }
        rotated = true;
        const double theta = (a(q, q) - a(p, p)) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double cs = 1.0 / std::sqrt(t * t + 1.0), sn = t * cs;
        a(p, p) -= t * apq;
        a(q, q) += t * apq;
        a(p, q) = a(q, p) = 0.0;
        for (int r = 0; r < n; r++) {
          if (r != p && r != q) {
            const double arp = a(r, p), arq = a(r, q);
            a(r, p) = a(p, r) = cs * arp - sn * arq;
            a(r, q) = a(q, r) = sn * arp + cs * arq;
          }
          const double zrp = Z[(size_t)r * n + p], zrq = Z[(size_t)r * n + q];
          Z[(size_t)r * n + p] = cs * zrp - sn * zrq;
          Z[(size_t)r * n + q] = sn * zrp + cs * zrq;
        }
      }
    if (!rotated) break;
  }
  std::vector<int> idx(n);
  for (int i = 0; i < n; i++) idx[i] = i;
  std::stable_sort(idx.begin(), idx.end(), [&](int x, int y) { return a(x, x) < a(y, y); });
  w.resize(n);
  std::vector<double> Z2((size_t)n * n);
  for (int c = 0; c < n; c++) {
    w[c] = a(idx[c], idx[c]);
    for (int r = 0; r < n; r++) Z2[(size_t)r * n + c] = Z[(size_t)r * n + idx[c]];
  }
  Z.swap(Z2);
}

struct TrlStats {
  int nconv = 0;     // wanted Ritz pairs whose error bound met the tolerance
  int nmatvec = 0;   // H x v products
  int nrestart = 0;  // restarts performed
};

// nev lowest eigenpairs; on return V[0..nev) hold the Ritz vectors and theta[0..nev) the Ritz values (ascending).
// ncv = size of the Krylov basis between restarts (nev < ncv <= kTrlMaxNcv, ncv < dimension of the space),
// maxrestart = ARPACK's mxiter, tol = ARPACK's tol: a pair counts as converged when its error bound
// |beta_m y_{m,i}| <= max(tol, eps) * max(eps^(2/3), |theta_i|)  (dsconv / znaup2's test; bounds below eps cannot be
// resolved from a dense solve of the projected problem, so the reference's default 1e-18 acts as eps), or <= eps * |H|.
template <class Backend>
int trl_solve(Backend &be, int nev, int ncv, int maxrestart, double tol, std::vector<double> &theta, TrlStats &st) {
  typedef std::complex<double> cplx;
  const double eps = 2.220446049250313e-16, eps23 = std::pow(eps, 2.0 / 3.0);
  const double tol_eff = std::max(tol, eps);
  // Second Gram-Schmidt pass when the first one removed more than half of |w|^2 (Daniel-Gragg-Kaufman-Stewart, eta = 1/sqrt 2,
  // ARPACK's rule).  In a Lanczos step w = H v_j is dominated by alfa_j v_j, so this fires almost always -- and it has to: with
  // ONE classical pass the orthogonality error of v_{j+1} is that of the basis times |h| / beta, which compounds step after
  // step whenever |alfa| > beta (a sector whose spectrum lies on one side of zero loses 1e-5 in 20 steps and everything
  // after the first restart; tests/test_sp_eigh_cpu.py::test_thick_restart_one_sided_spectrum).  "Twice is enough."
  const double eta2 = 0.5;
  const int m = ncv;
  std::vector<double> T((size_t)m * m, 0.0), w, Y;
  std::vector<cplx> h(m + 1), h2(m + 1);
  uint64_t seed = 1;
  double n2 = 0.0;
  int rc;
  // V_0 = normalised pseudo-random vector (ARPACK: info = 0 -> random start)
  if ((rc = be.rand_w(seed++))) return rc;
  if ((rc = be.orth_w(0, h.data(), &n2))) return rc;
  if (!(n2 > 0.0)) return -1;
  if ((rc = be.store_w(0, 1.0 / std::sqrt(n2)))) return rc;
  int k = 0;
  double beta_last = 0.0, anorm = 0.0;
  st = TrlStats();
  for (int restart = 0;; restart++) {
    for (int j = k; j < m; j++) {
      if ((rc = be.matvec(j))) return rc;
      st.nmatvec++;
      if ((rc = be.orth_w(j + 1, h.data(), &n2))) return rc;
      double alpha = h[j].real(), before2 = n2;
      for (int i = 0; i <= j; i++) before2 += std::norm(h[i]);
      if (n2 < eta2 * before2) {  // cancellation: orthogonalise once more ("twice is enough")
        if ((rc = be.orth_w(j + 1, h2.data(), &n2))) return rc;
        alpha += h2[j].real();
      }
      T[(size_t)j * m + j] = alpha;
      double beta = std::sqrt(std::max(n2, 0.0));
      anorm = std::max(anorm, std::fabs(alpha) + beta);
      if (beta <= 1e3 * eps * anorm) {
        // invariant subspace: continue with a fresh direction orthogonal to the basis, coupled by beta = 0
        bool ok = false;
        for (int attempt = 0; attempt < 5 && !ok; attempt++) {
          double r2 = 0.0, o2 = 0.0;
          if ((rc = be.rand_w(seed++))) return rc;
          if ((rc = be.orth_w(0, h2.data(), &r2))) return rc;
          if ((rc = be.orth_w(j + 1, h2.data(), &o2))) return rc;
          if ((rc = be.orth_w(j + 1, h2.data(), &o2))) return rc;
          if (o2 > 1e-20 * r2) { n2 = o2; ok = true; }
        }
        if (!ok) return -2;
        beta = 0.0;
        if ((rc = be.store_w(j + 1, 1.0 / std::sqrt(n2)))) return rc;
      } else {
        if ((rc = be.store_w(j + 1, 1.0 / beta))) return rc;
      }
      if (j + 1 < m) T[(size_t)j * m + j + 1] = T[(size_t)(j + 1) * m + j] = beta;
      else beta_last = beta;
    }
    jacobi_eigh(m, T, w, Y);
    int nconv = 0;
    for (int i = 0; i < nev; i++)
      // ARPACK's test, with the floor eps * |H| (anorm): a level that happens to sit at |theta| << |H| cannot be resolved
      // below the rounding of the products however long the iteration runs
      if (std::fabs(beta_last * Y[(size_t)(m - 1) * m + i]) <= std::max(tol_eff * std::max(eps23, std::fabs(w[i])), eps * anorm)) nconv++;
    st.nconv = nconv;
    st.nrestart = restart;
    const bool done = nconv >= nev || restart >= maxrestart;
    // vectors kept across the restart: the wanted ones + half of the rest (at least one more, at most m - 1)
    const int keep = done ? nev : std::min(m - 1, nev + std::max(1, (m - nev) / 2));
    std::vector<double> Yk((size_t)m * keep);
    for (int c = 0; c < keep; c++)
      for (int r = 0; r < m; r++) Yk[(size_t)c * m + r] = Y[(size_t)r * m + c];
    if ((rc = be.rotate(m, keep, Yk.data()))) return rc;
    if (done) {
      theta.assign(w.begin(), w.begin() + nev);
      return 0;
    }
    if ((rc = be.move(keep, m))) return rc;  // the residual direction becomes V_k
    std::fill(T.begin(), T.end(), 0.0);
    for (int i = 0; i < keep; i++) {
      T[(size_t)i * m + i] = w[i];
      T[(size_t)i * m + keep] = T[(size_t)keep * m + i] = beta_last * Y[(size_t)(m - 1) * m + i];
    }
    k = keep;
  }
}

// host backend of the CPU test hook: vectors are host arrays, the mat-vec is the caller's
struct TrlHostBackend {
  typedef std::complex<double> cplx;
  typedef void (*matvec_fn)(int64_t n, const double *v, double *hv, void *user);
  typedef void (*allreduce_fn)(double *buf, int64_t count, void *user);  // in-place sum over the ranks (NULL: one rank)
  int64_t n = 0, goff = 0;  // local length, global index of the first local element (sharded runs)
  matvec_fn mv = nullptr;
  allreduce_fn red = nullptr;
  void *user = nullptr;
  std::vector<std::vector<cplx>> V;
  std::vector<cplx> w;
  void init(int64_t n_, int64_t goff_, int nvec, matvec_fn f, allreduce_fn r, void *u) {
    n = n_; goff = goff_; mv = f; red = r; user = u;
    V.assign(nvec, std::vector<cplx>((size_t)n_));
    w.assign((size_t)n_, cplx(0.0, 0.0));
  }
  int rand_w(uint64_t seed) {
    for (int64_t i = 0; i < n; i++) w[i] = cplx(trl_rand(seed, (uint64_t)(goff + i)), 0.0);
    return 0;
  }
  int orth_w(int nv, cplx *h, double *nrm2) {
    for (int k = 0; k < nv; k++) {
      cplx s(0.0, 0.0);
      for (int64_t i = 0; i < n; i++) s += std::conj(V[k][i]) * w[i];
      h[k] = s;
    }
    if (red && nv > 0) red((double *)h, 2 * (int64_t)nv, user);
    for (int k = 0; k < nv; k++)
      for (int64_t i = 0; i < n; i++) w[i] -= h[k] * V[k][i];
    double s = 0.0;
    for (int64_t i = 0; i < n; i++) s += std::norm(w[i]);
    if (red) red(&s, 1, user);
    *nrm2 = s;
    return 0;
  }
  int store_w(int j, double scale) {
    for (int64_t i = 0; i < n; i++) V[j][i] = w[i] * scale;
    return 0;
  }
  int matvec(int j) {
    mv(n, (const double *)V[j].data(), (double *)w.data(), user);
    return 0;
  }
  int rotate(int m, int k, const double *Y) {
    std::vector<cplx> x(m);
    for (int64_t i = 0; i < n; i++) {
      for (int r = 0; r < m; r++) x[r] = V[r][i];
      for (int c = 0; c < k; c++) {
        cplx acc(0.0, 0.0);
        for (int r = 0; r < m; r++) acc += Y[(size_t)c * m + r] * x[r];
        V[c][i] = acc;
      }
    }
    return 0;
  }
  int move(int dst, int src) {
    V[dst] = V[src];
    return 0;
  }
};

}  // namespace cb
