// hxv_real.cu -- real-valued specialisation of the H x v kernels.
//
// The reference stores every vector as complex(8) (ED_VARS_GLOBAL.f90:72-78).  When the Hamiltonian
// is real (impHloc / Hbath without imaginary parts: BASELINE configs K1-K3, K5) and a Krylov run
// starts from a real vector, every Lanczos vector stays real -- the imaginary parts are exactly zero
// in the reference's complex arithmetic too -- so the fused Krylov drivers (lanczos.cu) may keep the
// vectors as 8-byte reals: half the HBM bytes and half the L1 wavefronts of the coalesced gathers.
// The C-ABI mat-vec (cdmft_b200_hxv) always stays complex.  Single rank, non-sharded layouts only.
//
// Same two-pass structure and thread mappings as the complex generic kernels in hxv.cu:
//   column pass  out(i,c)  = d(i,c) v(i,c) + sum_k Hup(i,j_k) v(j_k,c)      thread per row, CB columns
//   row pass     out(i,c) += sum_k Hdw(c,j_k) v(i,j_k)                      lanes along i, operator row uniform
#include "ctx.h"
#include "hxv_common.cuh"

namespace cb {

OpArgs op_args(const SpinOp &s);
int rowpass_real_as_pairs(int64_t nrows, const double *v, double *out, bool accum);  // hxv.cu
bool rowtile_applicable(const SpinOp &s);  // rowtile.cu
DiagArgs diag_args(int64_t coloff);

template <bool DIRECT, int CB>
__global__ void __launch_bounds__(256) k_colpass_r(int64_t n, int64_t ncols, const double *__restrict__ v,
                                                    double *__restrict__ out, OpArgs op, DiagArgs dg) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t c0 = (int64_t)blockIdx.y * CB;
  if (i >= n) return;
  double acc[CB];
  const uint32_t mu_imp = (uint32_t)__ldg(dg.map_row + i) & ((1u << dg.nimp) - 1u);
#pragma unroll
  for (int cc = 0; cc < CB; cc++) {
    acc[cc] = 0.0;
    if (c0 + cc < ncols) acc[cc] = diag_value(dg, i, mu_imp, c0 + cc) * __ldg(v + i + (c0 + cc) * n);
  }
  if (!DIRECT) {
    const int len = __ldg(op.rowlen + i);
    for (int k = 0; k < len; k++) {
      const int32_t j = __ldg(op.ell_col + (int64_t)k * n + i);
      const double h = __ldg(&op.ell_val[(int64_t)k * n + i].x);
#pragma unroll
      for (int cc = 0; cc < CB; cc++)
        if (c0 + cc < ncols) acc[cc] = fma(h, __ldg(v + j + (c0 + cc) * n), acc[cc]);
    }
  } else {
    const uint32_t s = (uint32_t)__ldg(op.map + i);
    for (int t = 0; t < op.nterms; t++) {
      const Term tm = op.terms[t];
      if (((s >> tm.a) & 1u) && !((s >> tm.b) & 1u)) {
        const uint32_t m = (s & ~(1u << tm.a)) | (1u << tm.b);
        const int32_t j = lin_rank_d(op.lin_lo, op.lin_hi, op.lbits, m);
        const double h = tm.re * hop_sign_d(s, tm.a, tm.b);
#pragma unroll
        for (int cc = 0; cc < CB; cc++)
          if (c0 + cc < ncols) acc[cc] = fma(h, __ldg(v + j + (c0 + cc) * n), acc[cc]);
      }
    }
  }
#pragma unroll
  for (int cc = 0; cc < CB; cc++)
    if (c0 + cc < ncols) out[i + (c0 + cc) * n] = acc[cc];
}

template <bool DIRECT>
__global__ void __launch_bounds__(256) k_rowpass_r(int64_t n /*rows*/, int64_t ncols, const double *__restrict__ v,
                                                    double *__restrict__ out, const int32_t *__restrict__ rowptr,
                                                    const int32_t *__restrict__ col, const double2 *__restrict__ val,
                                                    OpArgs op) {
  const int64_t c = blockIdx.x;
  const int64_t i = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double acc0 = 0.0, acc1 = 0.0;
  if (!DIRECT) {
    const int32_t p0 = __ldg(rowptr + c), p1 = __ldg(rowptr + c + 1);
    int32_t p = p0;
    for (; p + 4 <= p1; p += 4) {
      const double x0 = __ldg(v + i + (int64_t)__ldg(col + p) * n);
      const double x1 = __ldg(v + i + (int64_t)__ldg(col + p + 1) * n);
      const double x2 = __ldg(v + i + (int64_t)__ldg(col + p + 2) * n);
      const double x3 = __ldg(v + i + (int64_t)__ldg(col + p + 3) * n);
      acc0 = fma(__ldg(&val[p].x), x0, acc0);
      acc1 = fma(__ldg(&val[p + 1].x), x1, acc1);
      acc0 = fma(__ldg(&val[p + 2].x), x2, acc0);
      acc1 = fma(__ldg(&val[p + 3].x), x3, acc1);
    }
    for (; p < p1; p++) acc0 = fma(__ldg(&val[p].x), __ldg(v + i + (int64_t)__ldg(col + p) * n), acc0);
  } else {
    const uint32_t s = (uint32_t)__ldg(op.map + c);
    for (int t = 0; t < op.nterms; t++) {
      const Term tm = op.terms[t];
      if (((s >> tm.a) & 1u) && !((s >> tm.b) & 1u)) {
        const uint32_t m = (s & ~(1u << tm.a)) | (1u << tm.b);
        const int64_t j = lin_rank_d(op.lin_lo, op.lin_hi, op.lbits, m);
        acc0 = fma(tm.re * hop_sign_d(s, tm.a, tm.b), __ldg(v + i + j * n), acc0);
      }
    }
  }
  out[i + c * n] += acc0 + acc1;
}

// diag + Hup on a real vector: column-resident kernel when a column fits in shared memory, else the generic one
int colpass_real(const SpinOp &s, int64_t ncols, const double *v, double *out, const DiagArgs &dg, bool accum, bool final) {
  Ctx &c = ctx();
  if (ncols <= 0 || s.n <= 0) return 0;
  const bool direct = c.mode == CDMFT_B200_DIRECT;
  prof_begin(0);
  int rc = kColresNA;
  if (c.opt.colpass_variant == 6) {
    if (c.opt.colres_rows > 0 && !accum) rc = launch_colblk<double>(s, ncols, v, out, dg);
    if (rc == kColresNA) rc = launch_colres<double>(s, ncols, v, out, dg, accum, final);
    if (rc == kColresNA && !accum) rc = launch_colblk<double>(s, ncols, v, out, dg);
  }
  if (rc != 0 && rc != kColresNA) { prof_end(); return rc; }
  if (rc == kColresNA && accum) { prof_end(); return fail("internal: accumulating real column pass without the column-resident kernel"); }
  if (rc == kColresNA) {
    dim3 grid((unsigned)((s.n + 255) / 256), (unsigned)((ncols + 3) / 4));
    if (grid.y > 65535) { prof_end(); return fail("colpass_r: too many column groups"); }
    if (direct) k_colpass_r<true, 4><<<grid, 256, 0, c.stream>>>(s.n, ncols, v, out, op_args(s), dg);
    else k_colpass_r<false, 4><<<grid, 256, 0, c.stream>>>(s.n, ncols, v, out, op_args(s), dg);
    c.launches++;
  }
  prof_end();
  return 0;
}

int hxv_sharded_real(const double *v, double *hv);  // hxv.cu

// real-vector H x v; requires a real Hamiltonian and no Jx/Jp term.  One rank: column pass + row pass.
// Sharded layouts (SPMD, simulated ranks): the paired-row view of hxv.cu (even DimUp).
int hxv_device_real(const double *v, double *hv) {
  Ctx &c = ctx();
  if (!c.real_h || c.jhflag) return fail("hxv_device_real: not applicable");
  if (c.spmd || c.sim || c.opt.force_sharded) return hxv_sharded_real(v, hv);
  const bool direct = c.mode == CDMFT_B200_DIRECT;
  // preferred order (see hxv_local_terms): tile-resident row pass writes hv, column-resident pass accumulates
  if (!direct && !(c.dimup & 1) && rowtile_applicable(c.dw) && colres_applicable<double>(c.up, diag_args(0))) {
    CB_CHECK(rowpass_real_as_pairs(c.dimup, v, hv, false));
    return colpass_real(c.up, c.dimdw, v, hv, diag_args(0), true, true);
  }
  CB_CHECK(colpass_real(c.up, c.dimdw, v, hv, diag_args(0), false, false));
  if (!direct && !(c.dimup & 1)) {  // 16-byte gathers (two rows per lane)
    c.dot_final_rowpass = true;  // last contribution (real mode has no Jx/Jp term); pairs: u0*y0 + u1*y1
    const int rc = rowpass_real_as_pairs(c.dimup, v, hv, true);
    c.dot_final_rowpass = false;
    return rc;
  }
  {
    const SpinOp &s = c.dw;
    dim3 grid((unsigned)s.n, (unsigned)((c.dimup + 255) / 256));
    if (grid.y > 65535) return fail("rowpass_r: too many row chunks");
    prof_begin(1);
    if (direct) k_rowpass_r<true><<<grid, 256, 0, c.stream>>>(c.dimup, s.n, v, hv, s.rowptr, s.col, s.val, op_args(s));
    else k_rowpass_r<false><<<grid, 256, 0, c.stream>>>(c.dimup, s.n, v, hv, s.rowptr, s.col, s.val, op_args(s));
    c.launches++;
    prof_end();
  }
  return 0;
}

}  // namespace cb
