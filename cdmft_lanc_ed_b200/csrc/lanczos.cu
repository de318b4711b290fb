// lanczos.cu -- device-resident Krylov drivers around hxv_device and the Green's-function helpers.
//
// Restates the SciFortran SF_SP_LINALG routines the reference wraps around spHtimesV_p
// (un-vendored dependency, no version pin -- SURVEY.md App. B):
//   lanczos_iteration / sp_lanc_tridiag   caller ED_GF_NORMAL.f90:215 (and 7 more blocks)
//   sp_lanc_eigh                          caller ED_DIAG.f90:176-184
// and from the reference itself
//   vvinit = c^+_is|gs>, c_is|gs>, mixed channels       ED_GF_NORMAL.f90:180-199,244-266,590-620
//   add_to_lanczos_gf_normal                              ED_GF_NORMAL.f90:915-975
// The vectors never leave HBM between iterations and the scalars of the recurrence stay on the device; the
// drivers read (alfa, beta) back once per batch of steps (option lanczos_batch).
// Dot products use a fixed two-stage reduction tree -> bitwise reproducible run to run.
#include <algorithm>
#include <cmath>
#include <complex>
#include <vector>

#include "ctx.h"
#include "trlan.h"

namespace cb {

int hxv_device_real(const double *v, double *hv);  // hxv_real.cu

static const int kRedBlocks = 1024;  // stage-1 partials (fixed => deterministic)

__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
  return x;
}
// block reduction of two doubles; result valid in thread 0
__device__ __forceinline__ void block_sum2(double &a, double &b) {
  __shared__ double sa[32], sb[32];
  a = warp_sum(a);
  b = warp_sum(b);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { sa[w] = a; sb[w] = b; }
  __syncthreads();
  if (w == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    a = lane < nw ? sa[lane] : 0.0;
    b = lane < nw ? sb[lane] : 0.0;
    a = warp_sum(a);
    b = warp_sum(b);
  }
}

// element helpers: T = double2 (complex(8), the reference's vector type) or double (real mode)
__device__ __forceinline__ void dot_acc(double &re, double &im, double2 x, double2 y) {  // conj(x)*y
  re += x.x * y.x + x.y * y.y;
  im += x.x * y.y - x.y * y.x;
}
__device__ __forceinline__ void dot_acc(double &re, double &, double x, double y) { re += x * y; }
__device__ __forceinline__ double2 lin3(double a, double2 x, double b, double2 y, double c, double2 z) {
  return make_double2(a * x.x - b * y.x - c * z.x, a * x.y - b * y.y - c * z.y);
}
__device__ __forceinline__ double lin3(double a, double x, double b, double y, double c, double z) { return a * x - b * y - c * z; }
__device__ __forceinline__ double norm2(double2 x) { return x.x * x.x + x.y * x.y; }
__device__ __forceinline__ double norm2(double x) { return x * x; }
__device__ __forceinline__ double2 scaled(double2 x, double s) { return make_double2(x.x * s, x.y * s); }
__device__ __forceinline__ double scaled(double x, double s) { return x * s; }
__device__ __forceinline__ double2 axpy1(double2 a, double z, double2 x) { return make_double2(fma(z, x.x, a.x), fma(z, x.y, a.y)); }
__device__ __forceinline__ double axpy1(double a, double z, double x) { return fma(z, x, a); }
__device__ __forceinline__ void set_real(double2 &y, double re) { y = make_double2(re, 0.0); }
__device__ __forceinline__ void set_real(double &y, double re) { y = re; }

#define GRID_STRIDE(i, n) \
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += (int64_t)gridDim.x * blockDim.x)

// partial[blk] = sum conj(a)*b over the block's grid-stride share
template <typename T>
__global__ void __launch_bounds__(256) k_dot(int64_t n, const T *__restrict__ a, const T *__restrict__ b, double2 *__restrict__ partial) {
  double re = 0, im = 0;
  GRID_STRIDE(i, n) dot_acc(re, im, a[i], b[i]);
  block_sum2(re, im);
  if (threadIdx.x == 0) partial[blockIdx.x] = make_double2(re, im);
}
__global__ void k_reduce_final(int nparts, const double2 *__restrict__ partial, double *__restrict__ out) {
  double re = 0, im = 0;
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) { re += partial[i].x; im += partial[i].y; }
  block_sum2(re, im);
  if (threadIdx.x == 0) { out[0] = re; out[1] = im; }
}
// stage 1 of the fixed-tree sum of the row pass's per-CTA dot partials
__global__ void __launch_bounds__(256) k_sum_partials(int64_t n, const double *__restrict__ p, double2 *__restrict__ partial) {
  double re = 0, im = 0;
  GRID_STRIDE(i, n) re += p[i];
  block_sum2(re, im);
  if (threadIdx.x == 0) partial[blockIdx.x] = make_double2(re, 0.0);
}
template <typename T>
__global__ void __launch_bounds__(256) k_scale(int64_t n, T *__restrict__ v, double s) {
  GRID_STRIDE(i, n) v[i] = scaled(v[i], s);
}
template <typename T>
__global__ void __launch_bounds__(256) k_fill(int64_t n, T *__restrict__ v, double re) {
  GRID_STRIDE(i, n) set_real(v[i], re);
}
// acc += z * v
template <typename T>
__global__ void __launch_bounds__(256) k_axpy_real(int64_t n, T *__restrict__ acc, const T *__restrict__ v, double z) {
  GRID_STRIDE(i, n) acc[i] = axpy1(acc[i], z, v[i]);
}
// real <-> complex conversion; c2r also reduces sum |Im|^2 (real mode is only entered when it is 0)
__global__ void __launch_bounds__(256) k_c2r(int64_t n, const double2 *__restrict__ z, double *__restrict__ r, double2 *__restrict__ partial) {
  double im2 = 0, dummy = 0;
  GRID_STRIDE(i, n) {
    const double2 x = z[i];
    if (r) r[i] = x.x;
    im2 += x.y * x.y;
  }
  block_sum2(im2, dummy);
  if (threadIdx.x == 0) partial[blockIdx.x] = make_double2(im2, 0.0);
}
__global__ void __launch_bounds__(256) k_r2c(int64_t n, const double *__restrict__ r, double2 *__restrict__ z) {
  GRID_STRIDE(i, n) z[i] = make_double2(r[i], 0.0);
}

static unsigned vec_grid(int64_t n) {
  Ctx &c = ctx();
  int64_t nb = (n + 255) / 256;
  return (unsigned)std::max<int64_t>(1, std::min<int64_t>(nb, (int64_t)c.sm_count * 8));
}

// finish a reduction whose stage-1 partials sit in c.red (as double2[kRedBlocks]); all-reduce over
// ranks in SPMD mode; returns the complex sum on the host.
static int finish_reduce(std::complex<double> *out) {
  Ctx &c = ctx();
  double2 *partial = (double2 *)c.red;
  double *res = c.red + 2 * kRedBlocks;
  k_reduce_final<<<1, 256, 0, c.stream>>>(kRedBlocks, partial, res);
  c.launches++;
  CB_CHECK(nccl_allreduce_sum(res, 2));
  CB_CUDA(cudaMemcpyAsync(c.red_host, res, 16, cudaMemcpyDeviceToHost, c.stream));
  CB_CUDA(cudaStreamSynchronize(c.stream));
  *out = std::complex<double>(c.red_host[0], c.red_host[1]);
  return 0;
}
static int zero_partials() {
  Ctx &c = ctx();
  CB_CUDA(cudaMemsetAsync(c.red, 0, 2 * kRedBlocks * sizeof(double), c.stream));
  return 0;
}
template <typename T>
static int dot(int64_t n, const T *a, const T *b, std::complex<double> *out) {
  Ctx &c = ctx();
  CB_CHECK(zero_partials());
  if (n > 0) {
    k_dot<T><<<std::min<unsigned>(vec_grid(n), kRedBlocks), 256, 0, c.stream>>>(n, a, b, (double2 *)c.red);
    c.launches++;
  }
  return finish_reduce(out);
}

static int ensure_kv(int64_t n, int count) {
  Ctx &c = ctx();
  if (c.kv_n < n) {
    for (auto &k : c.kv) dev_free(k);
    c.kv_n = n;
  }
  for (int i = 0; i < count; i++)
    if (!c.kv[i]) CB_CHECK(dev_alloc(&c.kv[i], c.kv_n));
  return 0;
}

static int64_t local_n() {
  int64_t n = 0;
  for (auto &r : ctx().rk) n += r.nloc;
  return n;
}

static inline int hxv_t(const double2 *v, double2 *hv) { return hxv_device(v, hv); }
static inline int hxv_t(const double *v, double *hv) { return hxv_device_real(v, hv); }

// ------------------------------------------------------------------------------------
// The 3-term recurrence of SciFortran's lanczos_iteration, restated on UNNORMALISED
// vectors u_j = beta_j v_j so that no separate normalise / swap sweeps are needed:
//     t        = H u_j
//     alfa_j   = Re<u_j, t> / beta_j^2
//     u_{j+1}  = t/beta_j - (alfa_j/beta_j) u_j - (beta_j/beta_{j-1}) u_{j-1},   beta_{j+1} = |u_{j+1}|
// (identical to vout = H vin - alfa vin - beta vin_prev with vin = u_j/beta_j).  Per iteration the
// vector kernels move 96 B/state (one dot, one fused 3-term update + norm) instead of the textbook
// 176 B/state (swap/scale, add+dot, axpy+norm); half of that in real mode.
// ------------------------------------------------------------------------------------
// The scalars of the recurrence never visit the host inside a step: beta_{j-1}, beta_j live in scal[0..1], the two
// reductions of a step land in dotp[0] (Re<u_j, H u_j>) and nrm[0] (|u_{j+1}|^2) after their all-reduce, the update
// kernel derives its three coefficients from them, and k_lz_advance appends (alfa_j, beta_{j+1}) to a device
// history that the drivers read back once per batch of steps.
template <typename T>
__global__ void __launch_bounds__(256) k_lanczos_update(int64_t n, T *out /* may alias um */, const T *um, const T *__restrict__ u,
                                                        const T *__restrict__ t, const double *__restrict__ scal,
                                                        const double *__restrict__ dotp, double2 *__restrict__ partial) {
  const double bp = scal[0], bc = scal[1];
  const double a = dotp[0] / (bc * bc);
  const double ct = 1.0 / bc, cu = a / bc, cum = bc / bp;
  double re = 0, im = 0;
  GRID_STRIDE(i, n) {
    const T y = lin3(ct, t[i], cu, u[i], cum, um[i]);
    out[i] = y;
    re += norm2(y);
  }
  block_sum2(re, im);
  if (threadIdx.x == 0) partial[blockIdx.x] = make_double2(re, 0.0);
}
// eigenvector assembly from stored Krylov vectors: acc (=|+=) sum_k z[k] v[k], terms added in ascending k with
// one fma each -- the same operations, in the same order, as the step-by-step axpy of the two-pass scheme
constexpr int kAsmVecs = 8;
template <typename T>
struct AsmArgs {
  const T *v[kAsmVecs];
  double z[kAsmVecs];
  int nv;
  int first;  // 1: acc starts from zero
};
template <typename T>
__global__ void __launch_bounds__(256) k_assemble(int64_t n, T *__restrict__ acc, AsmArgs<T> a) {
  GRID_STRIDE(i, n) {
    T x;
    if (a.first) set_real(x, 0.0); else x = acc[i];
#pragma unroll
    for (int k = 0; k < kAsmVecs; k++)
      if (k < a.nv) x = axpy1(x, a.z[k], a.v[k][i]);
    acc[i] = x;
  }
}
__global__ void k_lz_init(double *scal) { scal[0] = scal[1] = 1.0; }
__global__ void k_lz_set(double *scal, double bp, double bc) { scal[0] = bp; scal[1] = bc; }
__global__ void k_lz_advance(double *scal, const double *__restrict__ dotp, const double *__restrict__ nrm, double *__restrict__ hist,
                             int iter /*0-based*/) {
  const double bc = scal[1];
  const double bn = sqrt(nrm[0]);
  hist[2 * iter] = dotp[0] / (bc * bc);
  hist[2 * iter + 1] = bn;
  scal[0] = bc;
  scal[1] = bn;
}

template <typename T>
struct LanczosRun {
  int64_t n = 0;
  T *um = nullptr, *u = nullptr, *t = nullptr;  // u_{j-1}, u_j, H u_j
  T *next = nullptr;                             // != nullptr: u_{j+1} is written here (Krylov vectors are kept) instead of over u_{j-1}
  int steps = 0;                                 // steps enqueued since lanczos_start
};

static double *lz_scal() { return ctx().red + 2 * kRedBlocks + 8; }  // [2]
static double *lz_dot() { return ctx().red + 2 * kRedBlocks; }       // [2]
static double *lz_nrm() { return ctx().red + 2 * kRedBlocks + 2; }   // [2]
static int lz_hist_reserve(int nsteps) {
  Ctx &c = ctx();
  if (c.lz_hist_cap >= nsteps) return 0;
  dev_free(c.lz_hist);
  CB_CHECK(dev_alloc(&c.lz_hist, (int64_t)2 * nsteps));
  c.lz_hist_cap = nsteps;
  return 0;
}
// stage-2 reduction of the partials in c.red into out[0..1] + all-reduce over ranks; stays on the device
static int finish_reduce_dev(double *out) {
  Ctx &c = ctx();
  k_reduce_final<<<1, 256, 0, c.stream>>>(kRedBlocks, (const double2 *)c.red, out);
  c.launches++;
  return nccl_allreduce_sum(out, 2);
}

// u holds the start vector on entry; normalises it (iter == 1 branch of lanczos_iteration)
template <typename T>
static int lanczos_start(LanczosRun<T> &L) {
  Ctx &c = ctx();
  std::complex<double> z;
  CB_CHECK(dot(L.n, L.u, L.u, &z));
  const double norm = std::sqrt(z.real());
  if (norm == 0.0) return fail("LANCZOS_ITERATION: norm(vin)=0");
  if (L.n > 0) {
    prof_begin(4);
    k_scale<T><<<vec_grid(L.n), 256, 0, c.stream>>>(L.n, L.u, 1.0 / norm);
    c.launches++;
    CB_CUDA(cudaMemsetAsync(L.um, 0, (size_t)L.n * sizeof(T), c.stream));
    prof_end();
  }
  k_lz_init<<<1, 1, 0, c.stream>>>(lz_scal());
  c.launches++;
  L.steps = 0;
  return 0;
}

// enqueue one Lanczos step (no host synchronisation): (alfa_j, beta_{j+1}) go to the device history at position
// L.steps; afterwards the normalised vector of this step is v_j = L.um / beta_j (buffers are rotated)
template <typename T>
static int lanczos_step(LanczosRun<T> &L) {
  Ctx &c = ctx();
  // alpha = Re<u, H u>: the last pass of H x v reduces it on the fly when it can (single rank)
  c.dot_request = c.opt.fuse_dot != 0;
  c.dot_done = false;
  int rc = hxv_t(L.u, L.t);
  c.dot_request = false;
  CB_CHECK(rc);
  prof_begin(4);
  CB_CHECK(zero_partials());
  if (c.dot_done) {
    k_sum_partials<<<kRedBlocks, 256, 0, c.stream>>>(c.dot_npartial, c.dot_partial, (double2 *)c.red);
    c.launches++;
  } else if (L.n > 0) {
    k_dot<T><<<std::min<unsigned>(vec_grid(L.n), kRedBlocks), 256, 0, c.stream>>>(L.n, L.u, L.t, (double2 *)c.red);
    c.launches++;
  }
  CB_CHECK(finish_reduce_dev(lz_dot()));
  CB_CHECK(zero_partials());
  if (L.n > 0) {
    k_lanczos_update<T><<<std::min<unsigned>(vec_grid(L.n), kRedBlocks), 256, 0, c.stream>>>(
        L.n, L.next ? L.next : L.um, L.um, L.u, L.t, lz_scal(), lz_dot(), (double2 *)c.red);
    c.launches++;
  }
  CB_CHECK(finish_reduce_dev(lz_nrm()));
  k_lz_advance<<<1, 1, 0, c.stream>>>(lz_scal(), lz_dot(), lz_nrm(), c.lz_hist, L.steps);
  c.launches++;
  prof_end();
  if (L.next) {
    L.um = L.u;
    L.u = L.next;
    L.next = nullptr;
  } else {
    std::swap(L.um, L.u);
  }
  L.steps++;
  return 0;
}
// read steps [from, to) of the device history (one copy + one synchronisation per batch)
static int lanczos_fetch(int from, int to, std::vector<double> &al, std::vector<double> &be) {
  Ctx &c = ctx();
  if (to <= from) return 0;
  std::vector<double> h((size_t)2 * (to - from));
  CB_CUDA(cudaMemcpyAsync(h.data(), c.lz_hist + 2 * from, h.size() * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  CB_CUDA(cudaStreamSynchronize(c.stream));
  CB_CUDA(cudaGetLastError());
  for (int k = 0; k < to - from; k++) {
    al.push_back(h[2 * k]);
    be.push_back(h[2 * k + 1]);
  }
  return 0;
}

// symmetric tridiagonal eigenproblem, implicit-shift QL (what SF_LINALG eigh(diag,subdiag,Ev) gets
// from LAPACK).  d[n] diagonal, e[n] with e[i] coupling i-1,i (e[0] unused); Z row-major n x n.
static int tridiag_eigh(int n, std::vector<double> &d, const std::vector<double> &e_in, std::vector<double> *Z) {
  std::vector<double> e(n + 1, 0.0);
  for (int i = 1; i < n; i++) e[i - 1] = e_in[i];
  if (Z) { Z->assign((size_t)n * n, 0.0); for (int i = 0; i < n; i++) (*Z)[(size_t)i * n + i] = 1.0; }
  auto z = [&](int r, int col) -> double & { return (*Z)[(size_t)r * n + col]; };
  for (int l = 0; l < n; l++) {
    for (int iter = 0;; iter++) {
      int m = l;
      for (; m < n - 1; m++) {
        double dd = std::fabs(d[m]) + std::fabs(d[m + 1]);
        if (std::fabs(e[m]) <= 2.220446049250313e-16 * dd) break;
      }
      if (m == l) break;
      if (iter == 300) return fail("tridiag_eigh: no convergence");
      double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
      double r = std::hypot(g, 1.0);
      g = d[m] - d[l] + e[l] / (g + std::copysign(r, g));
      double s = 1.0, cs = 1.0, p = 0.0;
      int i = m - 1;
      bool underflow = false;
      for (; i >= l; i--) {
        double f = s * e[i], b = cs * e[i];
        r = std::hypot(f, g);
        e[i + 1] = r;
        if (r == 0.0) { d[i + 1] -= p; e[m] = 0.0; underflow = true; break; }
        s = f / r; cs = g / r;
        g = d[i + 1] - p;
        r = (d[i] - g) * s + 2.0 * cs * b;
        p = s * r;
        d[i + 1] = g + p;
        g = cs * r - b;
        if (Z)
          for (int k = 0; k < n; k++) {
            double t = z(k, i + 1);
            z(k, i + 1) = s * z(k, i) + cs * t;
            z(k, i) = cs * z(k, i) - s * t;
          }
      }
      if (underflow) continue;
      d[l] -= p; e[l] = g; e[m] = 0.0;
    }
  }
  // ascending order
  std::vector<int> idx(n);
  for (int i = 0; i < n; i++) idx[i] = i;
  std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return d[a] < d[b]; });
  std::vector<double> d2(n);
  for (int i = 0; i < n; i++) d2[i] = d[idx[i]];
  if (Z) {
    std::vector<double> Z2((size_t)n * n);
    for (int k = 0; k < n; k++)
      for (int i = 0; i < n; i++) Z2[(size_t)k * n + i] = (*Z)[(size_t)k * n + idx[i]];
    Z->swap(Z2);
  }
  d.swap(d2);
  return 0;
}

// c^+_pos / c_pos applied to one spin index of every basis state (ED_GF_NORMAL.f90:180-194)
__global__ void __launch_bounds__(256) k_apply_op(int64_t idim, int64_t idimup, int64_t jdimup, int ispin, int iop, int pos0,
                                                  const int32_t *__restrict__ imap, const int32_t *__restrict__ jlo,
                                                  const int32_t *__restrict__ jhi, int lbits, double2 coef,
                                                  const double2 *__restrict__ state, double2 *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= idim) return;
  const int64_t iu = i % idimup, id = i / idimup;
  const uint32_t s = (uint32_t)__ldg(imap + (ispin == 1 ? iu : id));
  const uint32_t bit = 1u << pos0;
  const bool occ = (s & bit) != 0;
  if (iop > 0 ? occ : !occ) return;
  const double sg = (__popc(s & (bit - 1u)) & 1) ? -1.0 : 1.0;  // c / cdg sign rule, ED_SETUP.f90:807-833
  const uint32_t r = iop > 0 ? (s | bit) : (s & ~bit);
  const int64_t jr = __ldg(jhi + (r >> lbits)) + __ldg(jlo + (r & ((1u << lbits) - 1u)));
  const int64_t j = ispin == 1 ? jr + id * jdimup : iu + jr * jdimup;
  const double2 x = state[i];
  double2 o = out[j];
  o.x += sg * (coef.x * x.x - coef.y * x.y);
  o.y += sg * (coef.x * x.y + coef.y * x.x);
  out[j] = o;
}


// real mode = real Hamiltonian, one rank, no Jx/Jp, and a start vector without imaginary part
static bool real_mode_possible() {
  Ctx &c = ctx();
  if (!c.opt.real_lanczos || !c.real_h || c.jhflag) return false;
  // sharded layouts use the paired-row view (hxv.cu): DimUp must be even
  if (c.spmd || c.sim || c.opt.force_sharded) return (c.dimup & 1) == 0;
  return true;
}
// sum |Im z|^2 (and optionally r = Re z)
static int split_real(int64_t n, const double2 *z, double *r, double *im2) {
  Ctx &c = ctx();
  CB_CHECK(zero_partials());
  if (n > 0) {
    k_c2r<<<std::min<unsigned>(vec_grid(n), kRedBlocks), 256, 0, c.stream>>>(n, z, r, (double2 *)c.red);
    c.launches++;
  }
  std::complex<double> s;
  CB_CHECK(finish_reduce(&s));
  *im2 = s.real();
  return 0;
}

// Steps are enqueued in batches and (alfa, beta) read back once per batch; the stopping rules are then applied
// step by step exactly as SciFortran applies them, so the number of steps REPORTED and every returned coefficient
// are those of the step-by-step loop (steps enqueued past the stopping point are discarded).
template <typename T>
static int tridiag_run(LanczosRun<T> &L, int32_t nitermax, double threshold, double *alanc, double *blanc, int32_t *ndone) {
  for (int i = 0; i < nitermax; i++) { alanc[i] = 0; blanc[i] = 0; }
  CB_CHECK(lz_hist_reserve(std::max(nitermax, 1)));
  CB_CHECK(lanczos_start(L));
  std::vector<double> al, be;
  int done = 0;
  bool stop = false;
  const int batch = (int)std::max<int64_t>(1, ctx().opt.lanczos_batch);
  while (!stop && L.steps < nitermax) {
    const int from = L.steps, to = std::min(nitermax, from + batch);
    for (int k = from; k < to; k++) CB_CHECK(lanczos_step(L));
    CB_CHECK(lanczos_fetch(from, to, al, be));
    for (int iter = from + 1; iter <= to; iter++) {
      alanc[iter - 1] = al[iter - 1];
      done = iter;
      if (std::fabs(be[iter - 1]) < threshold) { stop = true; break; }
      if (iter < nitermax) blanc[iter] = be[iter - 1];
    }
  }
  if (ndone) *ndone = done;
  return 0;
}

// Pool of Krylov-vector slots (ground-state driver, "store" mode): HBM is large enough to keep every Lanczos
// vector of a sector that fits one GPU many times over (K3, real vectors: 69 x 1.3 GB), so the eigenvector is
// assembled from the stored vectors in one streaming pass instead of re-running the whole recurrence.
static int lz_slot(int j, size_t bytes, void **out) {
  Ctx &c = ctx();
  if (c.lz_slot_bytes < bytes) {  // slots of another sector / element type: start over
    for (void *p : c.lz_slots) cudaFree(p);
    c.lz_slots.clear();
    c.lz_slot_bytes = bytes;
  }
  while ((int)c.lz_slots.size() <= j) {
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, std::max<size_t>(c.lz_slot_bytes, 16));
    if (e != cudaSuccess) return fail("lanczos: cudaMalloc of a Krylov-vector slot failed: %s", cudaGetErrorString(e));
    c.lz_slots.push_back(p);
  }
  *out = c.lz_slots[j];
  return 0;
}
void lz_free_slots() {
  Ctx &c = ctx();
  for (void *p : c.lz_slots) cudaFree(p);
  c.lz_slots.clear();
  c.lz_slot_bytes = 0;
}
// how many slots every rank can afford (collective: all ranks get the same number)
static int lz_slot_budget(size_t bytes, int want, int *out) {
  Ctx &c = ctx();
  size_t fr = 0, tot = 0;
  CB_CUDA(cudaMemGetInfo(&fr, &tot));
  const size_t have = c.lz_slot_bytes >= bytes ? c.lz_slots.size() : 0;
  if (c.lz_slot_bytes < bytes) fr += c.lz_slots.size() * c.lz_slot_bytes;  // would be released first
  const size_t reserve = std::max<size_t>((size_t)6 << 30, tot / 16);
  size_t m = have + (fr > reserve ? (fr - reserve) / std::max<size_t>(bytes, 16) : 0);
  m = std::min<size_t>(m, (size_t)want);
  if (c.spmd && c.nranks > 1) {  // min over ranks = -max(-m): the library only loads the sum reduction
    // sum of one-hot? keep it simple: gather every rank's value through a sum into its own cell
    double *w = c.red + 3008;
    if (2 * c.nranks + 3008 > 4096) return fail("lanczos: too many ranks for the scratch");
    std::vector<double> mine(c.nranks, 0.0), all(c.nranks, 0.0);
    mine[c.rank] = (double)m;
    CB_CUDA(cudaMemcpyAsync(w, mine.data(), c.nranks * sizeof(double), cudaMemcpyHostToDevice, c.stream));
    CB_CHECK(nccl_allreduce_sum(w, c.nranks));
    CB_CUDA(cudaMemcpyAsync(all.data(), w, c.nranks * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    CB_CUDA(cudaStreamSynchronize(c.stream));
    for (int r = 0; r < c.nranks; r++) m = std::min<size_t>(m, (size_t)all[r]);
  }
  *out = (int)m;
  return 0;
}

// gs: start vector in (already on device, type T), eigenvector out (same buffer)
template <typename T>
static int gs_run(LanczosRun<T> &L, T *gs, int32_t nitermax, double threshold, int32_t ncheck, double *egs, int32_t *niter,
                  double *alanc_out, double *blanc_out) {
  Ctx &c = ctx();
  const int64_t nloc = L.n;
  T *const kv_u = L.u, *const kv_um = L.um;
  const size_t vbytes = (size_t)std::max<int64_t>(nloc, 1) * sizeof(T);
  // store mode: slot[j] = u_{j+1} (slot[0] = normalised start vector); budget agreed between the ranks
  int nslots = 0;
  if (c.opt.lanczos_store) CB_CHECK(lz_slot_budget(vbytes, nitermax + 1, &nslots));
  if (c.opt.lanczos_store > 1) nslots = (int)std::min<int64_t>(nslots, c.opt.lanczos_store);  // > 1: cap on the slots (tests)
  const bool storing = nslots >= 3;
  auto slot = [&](int j, T **p) -> int { void *q = nullptr; CB_CHECK(lz_slot(j, vbytes, &q)); *p = (T *)q; return 0; };
  if (storing) CB_CHECK(slot(0, &L.u));
  if (nloc > 0) CB_CUDA(cudaMemcpyAsync(L.u, gs, (size_t)nloc * sizeof(T), cudaMemcpyDeviceToDevice, c.stream));
  CB_CHECK(lz_hist_reserve(std::max(nitermax, 1)));
  CB_CHECK(lanczos_start(L));
  std::vector<double> al, bl;  // bl[i] couples i-1,i ; bl[0]=0
  std::vector<double> ha, hb;  // history as fetched: alfa_j, beta_{j+1}
  std::vector<double> d, Z;
  double esave = 0, e0 = 0;
  bool stop = false;
  const int batch = (int)std::max<int64_t>(1, c.opt.lanczos_batch);
  // where step k (0-based; input u_{k+1}, output u_{k+2}) writes: slot k+1 while slots last, then the two work
  // buffers (never over a stored vector), then in place between the work buffers
  auto set_next = [&](int k) -> int {
    L.next = nullptr;
    if (!storing) return 0;
    if (k + 1 < nslots) return slot(k + 1, &L.next);
    if (k + 1 == nslots) L.next = kv_um;
    else if (k + 1 == nslots + 1) L.next = kv_u;
    return 0;
  };
  while (!stop && L.steps < nitermax) {
    // nothing can stop before ncheck steps except an invariant subspace: first batch = ncheck steps
    const int from = L.steps;
    const int to = std::min<int>(nitermax, from + (from == 0 ? std::max<int>(batch, std::min<int>(ncheck, 64)) : batch));
    for (int k = from; k < to; k++) {
      CB_CHECK(set_next(k));
      CB_CHECK(lanczos_step(L));
    }
    CB_CHECK(lanczos_fetch(from, to, ha, hb));
    for (int iter = from + 1; iter <= to; iter++) {
      const double a = ha[iter - 1], b = hb[iter - 1];
      al.push_back(a);
      if ((int)bl.size() < (int)al.size()) bl.push_back(0.0);
      if (std::fabs(b) < threshold) { stop = true; break; }  // invariant subspace
      if (iter < nitermax) bl.push_back(b);
      d = al;
      std::vector<double> e(bl.begin(), bl.begin() + al.size());
      CB_CHECK(tridiag_eigh((int)al.size(), d, e, nullptr));
      e0 = d[0];
      if ((int)al.size() >= ncheck && std::fabs(e0 - esave) <= threshold) { stop = true; break; }
      esave = e0;
    }
  }
  const int nlanc = (int)al.size();
  d = al;
  {
    std::vector<double> e(bl.begin(), bl.begin() + nlanc);
    CB_CHECK(tridiag_eigh(nlanc, d, e, &Z));
  }
  e0 = d[0];
  // vect = sum_iter v_iter * Z(iter,1), v_iter = u_iter / beta_iter, beta_1 = 1, beta_iter = the value step iter-1 returned
  auto beta_of = [&](int iter) { return iter <= 1 ? 1.0 : hb[iter - 2]; };
  auto zcoef = [&](int iter) { return Z[(size_t)(iter - 1) * nlanc + 0] / beta_of(iter); };
  auto axpy_gs = [&](const T *vec, int iter) {
    if (nloc <= 0) return;
    prof_begin(4);
    k_axpy_real<T><<<vec_grid(nloc), 256, 0, c.stream>>>(nloc, gs, vec, zcoef(iter));
    c.launches++;
    prof_end();
  };
  if (storing) {
    // stored vectors u_1..u_m (slots 0..m-1): one streaming pass, terms added in ascending order
    const int m = std::min(nslots, nlanc);
    prof_begin(4);
    for (int i0 = 1; i0 <= m; i0 += kAsmVecs) {
      AsmArgs<T> aa{};
      aa.nv = std::min(kAsmVecs, m - i0 + 1);
      aa.first = i0 == 1;
      for (int k = 0; k < aa.nv; k++) {
        aa.v[k] = (const T *)c.lz_slots[i0 - 1 + k];
        aa.z[k] = zcoef(i0 + k);
      }
      if (nloc > 0) {
        k_assemble<T><<<vec_grid(nloc), 256, 0, c.stream>>>(nloc, gs, aa);
        c.launches++;
      }
    }
    prof_end();
    if (nlanc > m) {
      // the vectors past the last slot: re-run the recurrence from (u_{m-1}, u_m) = (slot m-2, slot m-1) with the
      // betas the first pass returned -- bitwise the same vectors as a second pass from the start would produce
      L.um = (T *)c.lz_slots[m - 2];
      L.u = (T *)c.lz_slots[m - 1];
      k_lz_set<<<1, 1, 0, c.stream>>>(lz_scal(), beta_of(m - 1), beta_of(m));
      c.launches++;
      L.steps = m - 1;
      for (int k = m - 1; k + 1 < nlanc; k++) {  // step k: u_{k+1} -> u_{k+2}
        CB_CHECK(set_next(k));
        CB_CHECK(lanczos_step(L));
        axpy_gs(L.u, k + 2);
      }
    }
  } else {
    // second pass: same recurrence, same start: bitwise the same vectors; no synchronisation inside
    L.u = kv_u; L.um = kv_um; L.next = nullptr;
    if (nloc > 0) {
      CB_CUDA(cudaMemcpyAsync(L.u, gs, (size_t)nloc * sizeof(T), cudaMemcpyDeviceToDevice, c.stream));
      CB_CUDA(cudaMemsetAsync(gs, 0, (size_t)nloc * sizeof(T), c.stream));
    }
    CB_CHECK(lanczos_start(L));
    for (int iter = 1; iter <= nlanc; iter++) {
      CB_CHECK(lanczos_step(L));
      axpy_gs(L.um, iter);
    }
  }
  std::complex<double> z;
  CB_CHECK(dot(nloc, gs, gs, &z));
  if (nloc > 0) { k_scale<T><<<vec_grid(nloc), 256, 0, c.stream>>>(nloc, gs, 1.0 / std::sqrt(z.real())); c.launches++; }
  *egs = e0;
  if (niter) *niter = nlanc;
  if (alanc_out) std::copy(al.begin(), al.end(), alanc_out);
  if (blanc_out) std::copy(bl.begin(), bl.begin() + nlanc, blanc_out);
  return 0;
}

// ------------------------------------------------------------------------------------
// cdmft_b200_eigh: device backend of the thick-restart Lanczos of trlan.h (sp_eigh, ED_DIAG.f90:150-170).
// The Krylov basis (ncv + 1 vectors) and the work vector live in the slot pool of the ground-state driver; one
// step = H x v + one classical Gram-Schmidt pass against the whole basis (k_multi_dot reads w once per 8 basis
// vectors, k_multi_axpy likewise; coefficients stay on the device between the two) + the norm; one read-back
// (coefficients + norm) per step.  Restarts rotate the basis in place (k_rotate: every thread owns one vector
// element of all basis vectors).  Fixed grids and reduction trees -> bitwise reproducible.
// ------------------------------------------------------------------------------------
constexpr int kMdVecs = 8;  // basis vectors per launch of the multi-dot / multi-axpy kernels
template <typename T>
struct MdArgs {
  const T *v[kMdVecs];
  int nv;
};
template <typename T>
__global__ void __launch_bounds__(256) k_multi_dot(int64_t n, MdArgs<T> a, const T *__restrict__ w, double2 *__restrict__ partial) {
  double re[kMdVecs], im[kMdVecs];
#pragma unroll
  for (int k = 0; k < kMdVecs; k++) re[k] = im[k] = 0.0;
  GRID_STRIDE(i, n) {
    const T x = w[i];
#pragma unroll
    for (int k = 0; k < kMdVecs; k++)
      if (k < a.nv) dot_acc(re[k], im[k], a.v[k][i], x);
  }
#pragma unroll
  for (int k = 0; k < kMdVecs; k++) {
    block_sum2(re[k], im[k]);
    if (threadIdx.x == 0) partial[(size_t)blockIdx.x * kMdVecs + k] = make_double2(re[k], im[k]);
    __syncthreads();  // block_sum2's shared scratch is reused by the next vector
  }
}
// out[g] = sum over the blocks of partial[(group(g) * nblocks + b) * kMdVecs + g % kMdVecs]; one CTA per basis vector
__global__ void __launch_bounds__(256) k_multi_reduce(int nblocks, const double2 *__restrict__ partial, double *__restrict__ out) {
  const int g = blockIdx.x, grp = g / kMdVecs, k = g % kMdVecs;
  double re = 0, im = 0;
  for (int b = threadIdx.x; b < nblocks; b += blockDim.x) {
    const double2 p = partial[((size_t)grp * nblocks + b) * kMdVecs + k];
    re += p.x;
    im += p.y;
  }
  block_sum2(re, im);
  if (threadIdx.x == 0) { out[2 * g] = re; out[2 * g + 1] = im; }
}
__device__ __forceinline__ double2 sub_cmul(double2 w, double hre, double him, double2 v) {  // w - h * v
  return make_double2(w.x - (hre * v.x - him * v.y), w.y - (hre * v.y + him * v.x));
}
__device__ __forceinline__ double sub_cmul(double w, double hre, double, double v) { return w - hre * v; }
// w -= sum_k h[k] v_k, coefficients (re, im) read from device memory
template <typename T>
__global__ void __launch_bounds__(256) k_multi_axpy(int64_t n, MdArgs<T> a, const double *__restrict__ h, T *__restrict__ w) {
  double hre[kMdVecs], him[kMdVecs];
#pragma unroll
  for (int k = 0; k < kMdVecs; k++) {
    hre[k] = k < a.nv ? h[2 * k] : 0.0;
    him[k] = k < a.nv ? h[2 * k + 1] : 0.0;
  }
  GRID_STRIDE(i, n) {
    T x = w[i];
#pragma unroll
    for (int k = 0; k < kMdVecs; k++)
      if (k < a.nv) x = sub_cmul(x, hre[k], him[k], a.v[k][i]);
    w[i] = x;
  }
}
template <typename T>
__global__ void __launch_bounds__(256) k_scale_to(int64_t n, T *__restrict__ out, const T *__restrict__ in, double s) {
  GRID_STRIDE(i, n) out[i] = scaled(in[i], s);
}
template <typename T>
__global__ void __launch_bounds__(256) k_rand_fill(int64_t n, int64_t goff, uint64_t seed, T *__restrict__ w) {
  GRID_STRIDE(i, n) set_real(w[i], trl_rand(seed, (uint64_t)(goff + i)));
}
template <typename T>
struct RotArgs {
  T *v[kTrlMaxNcv];
};
__device__ __forceinline__ void zero_of(double2 &x) { x = make_double2(0.0, 0.0); }
__device__ __forceinline__ void zero_of(double &x) { x = 0.0; }
// V[0..k) = V[0..m) Y in place: a thread reads its element of all m vectors before it writes any
template <typename T>
__global__ void __launch_bounds__(128) k_rotate(int64_t n, RotArgs<T> a, int m, int k, const double *__restrict__ Y) {
  extern __shared__ double sY[];
  for (int i = threadIdx.x; i < m * k; i += blockDim.x) sY[i] = Y[i];
  __syncthreads();
  GRID_STRIDE(i, n) {
    T x[kTrlMaxNcv];
    for (int r = 0; r < m; r++) x[r] = a.v[r][i];
    for (int c = 0; c < k; c++) {
      T acc;
      zero_of(acc);
      const double *y = sY + (size_t)c * m;
      for (int r = 0; r < m; r++) acc = axpy1(acc, y[r], x[r]);
      a.v[c][i] = acc;
    }
  }
}

template <typename T>
struct TrlDevBackend {
  typedef std::complex<double> cplx;
  int64_t n = 0, goff = 0;
  std::vector<T *> V;          // ncv + 1 basis vectors (slots of the pool)
  T *w = nullptr;              // work vector
  double2 *partial = nullptr;  // [ngroups * kRedBlocks * kMdVecs]
  double *coef = nullptr;      // device: (re, im) per basis vector, then |w|^2 at 2 * (kTrlMaxNcv + 1)
  double *ydev = nullptr;      // device copy of the rotation matrix
  std::vector<double> host;

  int rand_w(uint64_t seed) {
    Ctx &c = ctx();
    if (n > 0) { k_rand_fill<T><<<vec_grid(n), 256, 0, c.stream>>>(n, goff, seed, w); c.launches++; }
    return 0;
  }
  int orth_w(int nv, cplx *h, double *nrm2) {
    Ctx &c = ctx();
    const unsigned grid = std::min<unsigned>(vec_grid(n), kRedBlocks);
    const int ngrp = (nv + kMdVecs - 1) / kMdVecs;
    prof_begin(4);
    if (nv > 0) {
      for (int g = 0; g < ngrp; g++) {
        MdArgs<T> a{};
        a.nv = std::min(kMdVecs, nv - g * kMdVecs);
        for (int k = 0; k < a.nv; k++) a.v[k] = V[g * kMdVecs + k];
        if (n > 0) { k_multi_dot<T><<<grid, 256, 0, c.stream>>>(n, a, w, partial + (size_t)g * grid * kMdVecs); c.launches++; }
      }
      if (n > 0) { k_multi_reduce<<<nv, 256, 0, c.stream>>>((int)grid, partial, coef); c.launches++; }
      else CB_CUDA(cudaMemsetAsync(coef, 0, (size_t)2 * nv * sizeof(double), c.stream));
      CB_CHECK(nccl_allreduce_sum(coef, 2 * nv));
      for (int g = 0; g < ngrp; g++) {
        MdArgs<T> a{};
        a.nv = std::min(kMdVecs, nv - g * kMdVecs);
        for (int k = 0; k < a.nv; k++) a.v[k] = V[g * kMdVecs + k];
        if (n > 0) { k_multi_axpy<T><<<vec_grid(n), 256, 0, c.stream>>>(n, a, coef + 2 * g * kMdVecs, w); c.launches++; }
      }
    }
    double *nrm = coef + 2 * (kTrlMaxNcv + 1);
    CB_CHECK(zero_partials());
    if (n > 0) { k_dot<T><<<grid, 256, 0, c.stream>>>(n, w, w, (double2 *)c.red); c.launches++; }
    CB_CHECK(finish_reduce_dev(nrm));
    prof_end();
    host.resize(2 * (kTrlMaxNcv + 2));
    CB_CUDA(cudaMemcpyAsync(host.data(), coef, host.size() * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    CB_CUDA(cudaStreamSynchronize(c.stream));
    CB_CUDA(cudaGetLastError());
    for (int k = 0; k < nv; k++) h[k] = cplx(host[2 * k], host[2 * k + 1]);
    *nrm2 = host[2 * (kTrlMaxNcv + 1)];
    return 0;
  }
  int store_w(int j, double scale) {
    Ctx &c = ctx();
    if (n > 0) { k_scale_to<T><<<vec_grid(n), 256, 0, c.stream>>>(n, V[j], w, scale); c.launches++; }
    return 0;
  }
  int matvec(int j) { return hxv_t(V[j], w); }
  int rotate(int m, int k, const double *Y) {
    Ctx &c = ctx();
    CB_CUDA(cudaMemcpyAsync(ydev, Y, (size_t)m * k * sizeof(double), cudaMemcpyHostToDevice, c.stream));
    RotArgs<T> a{};
    for (int r = 0; r < m; r++) a.v[r] = V[r];
    if (n > 0) {
      prof_begin(4);
      const int64_t nb = (n + 127) / 128;
      const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(nb, (int64_t)c.sm_count * 8));
      k_rotate<T><<<grid, 128, (size_t)m * k * sizeof(double), c.stream>>>(n, a, m, k, ydev);
      c.launches++;
      prof_end();
    }
    CB_CUDA(cudaStreamSynchronize(c.stream));  // Y is a host temporary of the caller
    return 0;
  }
  int move(int dst, int src) {
    Ctx &c = ctx();
    if (n > 0) CB_CUDA(cudaMemcpyAsync(V[dst], V[src], (size_t)n * sizeof(T), cudaMemcpyDeviceToDevice, c.stream));
    return 0;
  }
};

template <typename T>
static int eigh_run(int64_t nloc, int nev, int ncv, int maxrestart, double tol, std::vector<double> &theta, TrlStats &st,
                    std::vector<T *> *ritz) {
  Ctx &c = ctx();
  TrlDevBackend<T> be;
  be.n = nloc;
  be.goff = (c.spmd && !c.rk.empty()) ? c.rk[0].dw.off * c.dimup : 0;
  const size_t vbytes = (size_t)std::max<int64_t>(nloc, 1) * sizeof(T);
  be.V.resize(ncv + 1);
  for (int j = 0; j <= ncv; j++) { void *p = nullptr; CB_CHECK(lz_slot(j, vbytes, &p)); be.V[j] = (T *)p; }
  { void *p = nullptr; CB_CHECK(lz_slot(ncv + 1, vbytes, &p)); be.w = (T *)p; }
  const int ngrp = (ncv + 1 + kMdVecs - 1) / kMdVecs;
  double *scratch = nullptr;
  const int64_t npart = (int64_t)ngrp * kRedBlocks * kMdVecs * 2, ncoef = 2 * (kTrlMaxNcv + 2), ny = (int64_t)kTrlMaxNcv * kTrlMaxNcv;
  CB_CHECK(dev_alloc(&scratch, npart + ncoef + ny));
  be.partial = (double2 *)scratch;
  be.coef = scratch + npart;
  be.ydev = be.coef + ncoef;
  cudaError_t e = cudaMemsetAsync(be.coef, 0, (size_t)ncoef * sizeof(double), c.stream);
  int rc = e == cudaSuccess ? trl_solve(be, nev, ncv, maxrestart, tol, theta, st) : fail("eigh: cudaMemsetAsync failed");
  cudaStreamSynchronize(c.stream);
  cudaFree(scratch);
  if (rc == -1) return fail("eigh: the start vector has zero norm");
  if (rc == -2) return fail("eigh: no direction left orthogonal to the Krylov basis (ncv too close to Dim)");
  CB_CHECK(rc);
  ritz->assign(be.V.begin(), be.V.begin() + nev);
  return 0;
}

// Fock map + Lin tables of one particle number, without hop terms (apply_op, scatter/gather helpers)
int cached_map_op(int npart, const SpinOp **out) {
  Ctx &c = ctx();
  auto it = c.map_ops.find(npart);
  if (it == c.map_ops.end()) {
    SpinOp *op = new SpinOp();
    std::vector<Term> none;
    std::vector<double> e0(c.ns, 0.0);
    const int rc = build_spin_op(*op, npart, none, e0, 0.0, false);
    if (rc) { free_spin_op(*op); delete op; return rc; }
    it = c.map_ops.emplace(npart, op).first;
  }
  *out = it->second;
  return 0;
}
void free_map_ops() {
  Ctx &c = ctx();
  for (auto &kv : c.map_ops) { free_spin_op(*kv.second); delete kv.second; }
  c.map_ops.clear();
}

}  // namespace cb

using namespace cb;

extern "C" {

int cdmft_b200_lanczos_tridiag(int64_t nloc, const void *v0, int32_t nitermax, double threshold, double *alanc,
                               double *blanc, int32_t *ndone) {
  CB_REQUIRE_INIT();
  Ctx &c = ctx();
  if (!c.hstatus) return fail("lanczos_tridiag: Hsector NOT set");
  if (nloc != local_n()) return fail("lanczos_tridiag: nloc mismatch");
  if (threshold <= 0) threshold = 1e-12;
  CB_CHECK(ensure_kv(std::max<int64_t>(nloc, 1), 3));
  double2 *z0 = c.kv[0];
  if (nloc > 0)
    CB_CUDA(cudaMemcpyAsync(z0, v0, (size_t)nloc * 16, is_device_ptr(v0) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c.stream));
  if (real_mode_possible()) {  // every rank takes the same branch: the reductions below are collective
    double *r = (double *)c.kv[1];  // u (real) in the first half of kv[1]
    double im2 = 1.0;
    CB_CHECK(split_real(nloc, z0, r, &im2));
    if (im2 == 0.0) {
      LanczosRun<double> L;
      L.n = nloc; L.u = r; L.um = r + nloc; L.t = (double *)c.kv[2];
      return tridiag_run(L, nitermax, threshold, alanc, blanc, ndone);
    }
  }
  LanczosRun<double2> L;
  L.n = nloc; L.u = c.kv[0]; L.um = c.kv[1]; L.t = c.kv[2];
  return tridiag_run(L, nitermax, threshold, alanc, blanc, ndone);
}

int cdmft_b200_lanczos_gs(int64_t nloc, void *vect, int32_t nitermax, double threshold, int32_t ncheck, double *egs,
                          int32_t *niter, double *alanc_out, double *blanc_out) {
  CB_REQUIRE_INIT();
  Ctx &c = ctx();
  if (!c.hstatus) return fail("lanczos_gs: Hsector NOT set");
  if (nloc != local_n()) return fail("lanczos_gs: nloc mismatch");
  if ((int64_t)nitermax > c.dim) nitermax = (int32_t)c.dim;
  if (ncheck <= 0) ncheck = 10;
  if (threshold <= 0) threshold = 1e-12;  // same default as lanczos_tridiag
  CB_CHECK(ensure_kv(std::max<int64_t>(nloc, 1), 3));
  const bool dev = is_device_ptr(vect);
  double2 *gs = nullptr;  // start vector, later the accumulated eigenvector
  if (dev) gs = (double2 *)vect;
  else {
    CB_CHECK(ensure_stage(std::max<int64_t>(nloc, 1)));
    gs = c.stage_v;
    if (nloc > 0) CB_CUDA(cudaMemcpyAsync(gs, vect, (size_t)nloc * 16, cudaMemcpyHostToDevice, c.stream));
  }
  std::complex<double> z;
  CB_CHECK(dot(nloc, gs, gs, &z));
  if (z.real() == 0.0) {  // SciFortran start vector is unpinned; constant 1/sqrt(Dim)
    if (nloc > 0) { k_fill<double2><<<vec_grid(nloc), 256, 0, c.stream>>>(nloc, gs, 1.0 / std::sqrt((double)c.dim)); c.launches++; }
  }
  bool done = false;
  if (real_mode_possible()) {  // every rank takes the same branch: the reductions below are collective
    // real mode: u, um in kv[0] (two halves), t and the real eigenvector in kv[1]
    double *gr = (double *)c.kv[1] + nloc;
    double im2 = 1.0;
    CB_CHECK(split_real(nloc, gs, gr, &im2));
    if (im2 == 0.0) {
      LanczosRun<double> L;
      L.n = nloc; L.u = (double *)c.kv[0]; L.um = (double *)c.kv[0] + nloc; L.t = (double *)c.kv[1];
      CB_CHECK(gs_run(L, gr, nitermax, threshold, ncheck, egs, niter, alanc_out, blanc_out));
      k_r2c<<<vec_grid(nloc), 256, 0, c.stream>>>(nloc, gr, gs);
      c.launches++;
      done = true;
    }
  }
  if (!done) {
    LanczosRun<double2> L;
    L.n = nloc; L.u = c.kv[0]; L.um = c.kv[1]; L.t = c.kv[2];
    CB_CHECK(gs_run(L, gs, nitermax, threshold, ncheck, egs, niter, alanc_out, blanc_out));
  }
  if (!dev && nloc > 0) CB_CUDA(cudaMemcpyAsync(vect, gs, (size_t)nloc * 16, cudaMemcpyDeviceToHost, c.stream));
  CB_CUDA(cudaStreamSynchronize(c.stream));
  return 0;
}

// sp_eigh(MatVec, eig_values, eig_basis, Nblock, Nitermax, tol) -- the reference's default LANC_METHOD (ED_DIAG.f90:94-97,
// 150-170; SciFortran wraps (P)ARPACK, which = 'SA'): the neigen lowest eigenpairs of the active sector by a device-resident
// thick-restart Lanczos (trlan.h).  Collective in SPMD mode (every dot product is all-reduced, the start vector is a
// function of the global index): the P-ARPACK call of the reference, without a single vector crossing PCIe.
int cdmft_b200_eigh(int64_t nloc, int32_t neigen, int32_t nblock, int32_t nitermax, double tol, double *eig_values,
                    void *eig_basis, int32_t *nconv, int32_t *nmatvec) {
  CB_REQUIRE_INIT();
  Ctx &c = ctx();
  if (!c.hstatus) return fail("eigh: Hsector NOT set");
  if (nloc != local_n()) return fail("eigh: nloc mismatch");
  if (neigen < 1) return fail("eigh: Neigen must be positive");
  if ((int64_t)neigen + 1 >= c.dim)
    return fail("eigh: Neigen = %d needs a sector of more than %d states (the reference diagonalises such sectors densely, ED_DIAG.f90:104-106)",
                neigen, neigen + 1);
  // ncv: ARPACK wants nev < ncv <= n; here additionally ncv + 1 orthonormal vectors must exist and the rotation kernel
  // holds at most kTrlMaxNcv of them
  int ncv = nblock > 0 ? nblock : std::max(2 * neigen, 20);
  ncv = (int)std::min<int64_t>(std::min<int64_t>(ncv, kTrlMaxNcv), c.dim - 1);
  if (ncv <= neigen) ncv = neigen + 1;
  if (ncv > kTrlMaxNcv) return fail("eigh: Neigen = %d exceeds the %d basis vectors of the restart kernel", neigen, kTrlMaxNcv - 1);
  if (nitermax < 0) nitermax = 0;
  const bool real = real_mode_possible();  // the start vector is real: real H, no Jx/Jp -> real Krylov vectors
  const size_t vbytes = (size_t)std::max<int64_t>(nloc, 1) * (real ? sizeof(double) : sizeof(double2));
  int have = 0;
  CB_CHECK(lz_slot_budget(vbytes, ncv + 2, &have));
  if (have < ncv + 2) {
    if (have < neigen + 3) return fail("eigh: HBM holds only %d vectors of this sector, Neigen = %d needs %d", have, neigen, neigen + 3);
    ncv = have - 2;
  }
  std::vector<double> theta;
  TrlStats st;
  std::vector<double *> rz_r;
  std::vector<double2 *> rz_c;
  if (real) CB_CHECK(eigh_run<double>(nloc, neigen, ncv, nitermax, tol, theta, st, &rz_r));
  else CB_CHECK(eigh_run<double2>(nloc, neigen, ncv, nitermax, tol, theta, st, &rz_c));
  for (int i = 0; i < neigen; i++) eig_values[i] = theta[i];
  if (nconv) *nconv = st.nconv;
  if (nmatvec) *nmatvec = st.nmatvec;
  if (eig_basis && nloc > 0) {
    const bool dev = is_device_ptr(eig_basis);
    if (real && !dev) CB_CHECK(ensure_kv(nloc, 1));
    for (int i = 0; i < neigen; i++) {
      double2 *dst = (double2 *)eig_basis + (size_t)i * nloc;
      if (real) {
        double2 *z = dev ? dst : c.kv[0];
        k_r2c<<<vec_grid(nloc), 256, 0, c.stream>>>(nloc, rz_r[i], z);
        c.launches++;
        if (!dev) CB_CUDA(cudaMemcpyAsync(dst, z, (size_t)nloc * 16, cudaMemcpyDeviceToHost, c.stream));
      } else {
        CB_CUDA(cudaMemcpyAsync(dst, rz_c[i], (size_t)nloc * 16, dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c.stream));
      }
    }
  }
  CB_CUDA(cudaStreamSynchronize(c.stream));
  CB_CUDA(cudaGetLastError());
  return 0;
}

// Host-only test hook (no CUDA call, like cdmft_b200_schedule_host): the restart logic of cdmft_b200_eigh (trl_solve)
// on host vectors around the CALLER's mat-vec -- the CPU tests drive it with the oracle's H x v.  Not a product path.
// Sharded runs (the gloo test): n = local length, goff = global index of the first local element, ntot = global length,
// allreduce = in-place sum over the ranks (NULL with one rank) -- the same contract the device backend has with NCCL.
int cdmft_b200_eigh_logic_host(int64_t n, int64_t goff, int64_t ntot, void (*matvec)(int64_t, const double *, double *, void *),
                               void (*allreduce)(double *, int64_t, void *), void *user, int32_t neigen, int32_t nblock,
                               int32_t nitermax, double tol, double *eig_values, double *eig_basis, int32_t *nconv,
                               int32_t *nmatvec) {
  if (!matvec || n < 0 || goff < 0 || ntot < 3 || n + goff > ntot || neigen < 1 || (int64_t)neigen + 1 >= ntot)
    return fail("eigh_logic_host: bad arguments");
  int ncv = nblock > 0 ? nblock : std::max(2 * neigen, 20);
  ncv = (int)std::min<int64_t>(std::min<int64_t>(ncv, kTrlMaxNcv), ntot - 1);
  if (ncv <= neigen) ncv = neigen + 1;
  if (ncv > kTrlMaxNcv) return fail("eigh_logic_host: Neigen too large");
  TrlHostBackend be;
  be.init(n, goff, ncv + 1, matvec, allreduce, user);
  std::vector<double> theta;
  TrlStats st;
  const int rc = trl_solve(be, neigen, ncv, std::max(nitermax, 0), tol, theta, st);
  if (rc) return fail("eigh_logic_host: thick-restart Lanczos failed (%d)", rc);
  for (int i = 0; i < neigen; i++) {
    eig_values[i] = theta[i];
    if (eig_basis && n > 0) std::copy((const double *)be.V[i].data(), (const double *)be.V[i].data() + 2 * n, eig_basis + (size_t)2 * n * i);
  }
  if (nconv) *nconv = st.nconv;
  if (nmatvec) *nmatvec = st.nmatvec;
  return 0;
}

int cdmft_b200_apply_op(int32_t isector, int32_t iop, int32_t ispin, int32_t nops, const int32_t *pos, const double *coef,
                        const void *state, void *out, int32_t *jsector) {
  CB_REQUIRE_INIT();
  Ctx &c = ctx();
  if (!c.have_model) return fail("apply_op: no model set");
  if (ispin != 1 && ispin != 2) return fail("apply_op: ispin must be 1 or 2");
  // SPMD: the vector is sharded along Ndw, so a spin-UP operator maps every column onto the same column of the
  // target sector (same DimDw, same split): each rank transforms its own shard, no communication.  Spin-down
  // operators change the number of columns and with it the split: the state is gathered on rank 0, transformed
  // there and scattered with the target sector's split -- the reference's own route (es_return_cvector gathers on
  // the master, ED_EIGENSPACE.f90:499-569; scatter_vector_MPI before sp_lanc_tridiag, ED_GF_NORMAL.f90:214).
  const bool spmd_dw = c.spmd && c.nranks > 1 && ispin != 1;
  const int ns = c.ns;
  for (int k = 0; k < nops; k++)  // validate everything before any resource exists
    if (pos[k] < 1 || pos[k] > ns) return fail("apply_op: pos out of range");
  int nup = (isector - 1) / (ns + 1), ndw = (isector - 1) % (ns + 1);
  int jnup = nup + (ispin == 1 ? (iop > 0 ? 1 : -1) : 0), jndw = ndw + (ispin == 2 ? (iop > 0 ? 1 : -1) : 0);
  if (jnup < 0 || jnup > ns || jndw < 0 || jndw > ns) { if (jsector) *jsector = 0; return 0; }  // getCsector = 0
  if (jsector) *jsector = 1 + jnup * (ns + 1) + jndw;
  int64_t idimup, idimdw, idim, jdimup, jdimdw, jdim;
  CB_CHECK(cdmft_b200_get_sector_dims(isector, &idimup, &idimdw, &idim));
  CB_CHECK(cdmft_b200_get_sector_dims(1 + jnup * (ns + 1) + jndw, &jdimup, &jdimdw, &jdim));
  if (spmd_dw) {
    // gather (source sector split) -> full-vector transform on rank 0 -> scatter (target sector split)
    const bool root = c.rank == 0;
    double2 *gfull = nullptr, *ofull = nullptr;
    auto body2 = [&]() -> int {
      if (root) {
        CB_CHECK(dev_alloc(&gfull, idim));
        CB_CHECK(dev_alloc(&ofull, jdim));
      }
      CB_CHECK(scatter_gather_dims(gfull, (void *)state, 0, false, idimup, idimdw));
      if (root) {
        const SpinOp *src = nullptr, *dst = nullptr;
        CB_CHECK(cached_map_op(ndw, &src));
        CB_CHECK(cached_map_op(jndw, &dst));
        CB_CUDA(cudaMemsetAsync(ofull, 0, jdim * 16, c.stream));
        for (int k = 0; k < nops; k++) {
          if (idim == 0) continue;
          k_apply_op<<<(unsigned)((idim + 255) / 256), 256, 0, c.stream>>>(idim, idimup, jdimup, ispin, iop, pos[k] - 1, src->map, dst->lin_lo,
                                                                            dst->lin_hi, ns / 2, make_double2(coef[2 * k], coef[2 * k + 1]),
                                                                            gfull, ofull);
          c.launches++;
        }
      }
      CB_CHECK(scatter_gather_dims(ofull, out, 0, true, jdimup, jdimdw));
      return 0;
    };
    const int rc2 = body2();
    if (gfull) cudaFree(gfull);
    if (ofull) cudaFree(ofull);
    return rc2;
  }
  if (c.spmd) {  // local shard: this rank's columns (ED_HAMILTONIAN.f90:92-105 with P_eff = min(P, DimDw))
    const int peff = (int)std::min<int64_t>(c.nranks, idimdw);
    const int64_t q = c.rank < peff ? split_of(idimdw, peff, c.rank).q : 0;
    idim = idimup * q;
    jdim = jdimup * q;
  }
  // Fock maps + Lin tables of the two particle numbers: cached per particle number (the GF loops of
  // ED_GF_NORMAL.f90 call this once per channel with the same sectors; the cache dies with the model)
  const SpinOp *src = nullptr, *dst = nullptr;
  CB_CHECK(cached_map_op(ispin == 1 ? nup : ndw, &src));
  CB_CHECK(cached_map_op(ispin == 1 ? jnup : jndw, &dst));
  const bool ds = is_device_ptr(state), dd = is_device_ptr(out);
  double2 *d_state = (double2 *)state, *d_out = (double2 *)out;
  int rc = 0;
  auto body = [&]() -> int {
    if (!ds) { CB_CHECK(dev_alloc(&d_state, idim)); CB_CUDA(cudaMemcpyAsync(d_state, state, idim * 16, cudaMemcpyHostToDevice, c.stream)); }
    if (!dd) CB_CHECK(dev_alloc(&d_out, jdim));
    CB_CUDA(cudaMemsetAsync(d_out, 0, jdim * 16, c.stream));
    for (int k = 0; k < nops; k++) {
      if (idim == 0) continue;
      k_apply_op<<<(unsigned)((idim + 255) / 256), 256, 0, c.stream>>>(idim, idimup, jdimup, ispin, iop, pos[k] - 1, src->map,
                                                                        dst->lin_lo, dst->lin_hi, ns / 2,
                                                                        make_double2(coef[2 * k], coef[2 * k + 1]), d_state, d_out);
      c.launches++;
    }
    if (!dd) CB_CUDA(cudaMemcpyAsync(out, d_out, jdim * 16, cudaMemcpyDeviceToHost, c.stream));
    CB_CUDA(cudaStreamSynchronize(c.stream));
    CB_CUDA(cudaGetLastError());
    return 0;
  };
  rc = body();  // single clean-up path, whatever happened
  if (!ds && d_state) cudaFree(d_state);
  if (!dd && d_out) cudaFree(d_out);
  return rc;
}

}  // extern "C"
namespace cb {
int expect_terms(const std::vector<Term> &tu, const std::vector<Term> &td, const double2 *dv, double out[2]) {
  Ctx &c = ctx();
  const int64_t nloc = local_n();
  CB_CHECK(ensure_stage(std::max<int64_t>(nloc, 1)));
  SpinOp ku, kd;
  const std::vector<double> e0(c.ns, 0.0);
  int rc = build_spin_op(ku, c.up.npart, tu, e0, 0.0, false);
  if (rc == 0) rc = build_spin_op(kd, c.dw.npart, td, e0, 0.0, false);
  if (rc == 0) {
    std::swap(c.up, ku);
    std::swap(c.dw, kd);
    const bool jh = c.jhflag, tb = c.tables;
    c.jhflag = false;
    c.tables = false;  // one-term operators: the matrix-free kernels, nothing to build
    c.kin_only = true;
    rc = hxv_device(dv, c.stage_hv);
    c.kin_only = false;
    c.tables = tb;
    c.jhflag = jh;
    std::swap(c.up, ku);
    std::swap(c.dw, kd);
  }
  free_spin_op(ku);
  free_spin_op(kd);
  CB_CHECK(rc);
  std::complex<double> z;
  CB_CHECK(dot(nloc, dv, (const double2 *)c.stage_hv, &z));
  out[0] = z.real();
  out[1] = z.imag();
  return 0;
}
}  // namespace cb
extern "C" {

// <vec| K |vec> for K = the impurity block of the hopping part of H (off-diagonal impHloc, both spins): the only piece of
// lanc_local_energy (ED_OBSERVABLES.f90:246-460) that is not a function of the impurity occupations -- the reference
// accumulates impHloc(is,js) sg1 sg2 vec(i) conjg(vec(j)) over the hops |j> = c^+_is c_js |i> (:305-345).  One product
// with the restricted operator through the regular H x v path (any layout: single rank, simulated ranks, SPMD), then
// a dot product; out = (Re, Im), all-reduced over the ranks.  vec = local shard of a vector of the ACTIVE sector.
int cdmft_b200_imp_kinetic(int64_t nloc, const void *vec, double out[2]) {
  CB_REQUIRE_INIT();
  Ctx &c = ctx();
  if (!c.hstatus) return fail("imp_kinetic: Hsector NOT set");
  if (nloc != local_n()) return fail("imp_kinetic: nloc mismatch");
  if (!c.kin_built) {
    auto imp_only = [&](const std::vector<Term> &all) {
      std::vector<Term> t;
      for (const Term &x : all)
        if (x.a < c.nimp && x.b < c.nimp) t.push_back(x);
      return t;
    };
    const std::vector<double> e0(c.ns, 0.0);
    int rc = build_spin_op(c.kin_up, c.up.npart, imp_only(c.terms_up), e0, 0.0, c.tables);
    if (rc == 0) rc = build_spin_op(c.kin_dw, c.dw.npart, imp_only(c.terms_dw), e0, 0.0, c.tables);
    if (rc) { free_spin_op(c.kin_up); free_spin_op(c.kin_dw); return rc; }
    c.kin_built = true;
  }
  CB_CHECK(ensure_stage(std::max<int64_t>(nloc, 1)));
  const double2 *dv = (const double2 *)vec;
  if (nloc > 0 && !is_device_ptr(vec)) {
    CB_CUDA(cudaMemcpyAsync(c.stage_v, vec, (size_t)nloc * 16, cudaMemcpyHostToDevice, c.stream));
    dv = c.stage_v;
  }
  std::swap(c.up, c.kin_up);
  std::swap(c.dw, c.kin_dw);
  const bool jh = c.jhflag;
  c.jhflag = false;
  c.kin_only = true;
  int rc = hxv_device(dv, c.stage_hv);
  c.kin_only = false;
  c.jhflag = jh;
  std::swap(c.up, c.kin_up);
  std::swap(c.dw, c.kin_dw);
  CB_CHECK(rc);
  std::complex<double> z;
  CB_CHECK(dot(nloc, dv, (const double2 *)c.stage_hv, &z));
  out[0] = z.real();
  out[1] = z.imag();
  return 0;
}

int cdmft_b200_add_to_lanczos_gf_full(const double vnorm2[2], double ei, double egs, int32_t finite_t, double beta, int32_t nlanc,
                                      const double *alanc, const double *blanc, int32_t isign, double zeta, int32_t lmats,
                                      const double *wm, double *gmats, int32_t lreal, const double *wr, double eps, double *greal,
                                      double *poles, double *weights) {
  if (nlanc <= 0) return fail("add_to_lanczos_gf: nlanc must be positive");
  std::vector<double> d(alanc, alanc + nlanc), e(blanc, blanc + nlanc), Z;
  CB_CHECK(tridiag_eigh(nlanc, d, e, &Z));
  const std::complex<double> vn(vnorm2[0], vnorm2[1]);
  std::complex<double> pesoBZ(0.0, 0.0);  // Boltzmann factor of the state the channel was built on
  if (finite_t) {
    if (beta * (ei - egs) < 200.0) pesoBZ = vn * std::exp(-beta * (ei - egs)) / zeta;
  } else {
    pesoBZ = vn / zeta;
  }
  std::complex<double> *gm = (std::complex<double> *)gmats, *gr = (std::complex<double> *)greal;
  for (int j = 0; j < nlanc; j++) {
    const double pole = (double)isign * (d[j] - ei);
    const std::complex<double> peso = pesoBZ * Z[j] * Z[j];  // first row of the eigenvector matrix
    if (poles) poles[j] = pole;
    if (weights) { weights[2 * j] = peso.real(); weights[2 * j + 1] = peso.imag(); }
    for (int i = 0; i < lmats; i++) gm[i] += peso / (std::complex<double>(0.0, wm[i]) - pole);
    for (int i = 0; i < lreal; i++) gr[i] += peso / (std::complex<double>(wr[i], eps) - pole);
  }
  return 0;
}

int cdmft_b200_add_to_lanczos_gf(const double vnorm2[2], double ei, int32_t nlanc, const double *alanc, const double *blanc,
                                 int32_t isign, double zeta, int32_t lmats, const double *wm, double *g, double *poles,
                                 double *weights) {
  return cdmft_b200_add_to_lanczos_gf_full(vnorm2, ei, ei, 0, 0.0, nlanc, alanc, blanc, isign, zeta, lmats, wm, g, 0, nullptr, 0.0,
                                           nullptr, poles, weights);
}

}  // extern "C"
