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

// ------------------------------------------------------------------------------------
// Column-resident pass on TWO real columns at a time (real Krylov vectors, fast decodes).  The one-column kernel
// k_colres<double> is bound by instruction issue, not by shared-memory bandwidth: every 8-byte gather carries the
// whole decode of its operator word.  Two columns (c, c+1) resident side by side -- exactly the footprint of one
// complex column -- share the decode, the operator stream and the metadata: per step one word, two LDS.64 at the
// same offset of the two planes, two FMAs.  Same edge-coloured schedule (G = 16), dealt for 32 warps.
// ------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(1024, 1) k_colres2(int64_t n, int64_t ncols, const double *__restrict__ v, double *__restrict__ out,
                                                      ColResArgs a, DiagArgs dg) {
  constexpr int G = 16;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // layout: [0,8) mbarrier | two diagonal tables [2^nimp doubles each] | plane 0 | plane 1 (column + G zero elements each)
  uint64_t *bar = (uint64_t *)smem_raw;
  const int ndt = dg.enabled ? (1 << dg.nimp) : 0;
  double *dtab0 = (double *)(smem_raw + 128);
  double *dtab1 = dtab0 + ndt;
  const int64_t npad = (n + G - 1) / G * G;
  const uint32_t pstride = (uint32_t)((npad + G) * 8);  // bytes between the planes
  double *xs0 = (double *)(smem_raw + 128 + (((size_t)ndt * 16 + 127) & ~(size_t)127));
  double *xs1 = (double *)((char *)xs0 + pstride);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) mbar_init(bar, 1);
  for (int64_t k = n + threadIdx.x; k < npad + G; k += blockDim.x) { xs0[k] = 0.0; xs1[k] = 0.0; }  // idle lanes gather these
  const int t0 = __ldg(a.tbase + warp), t1 = __ldg(a.tbase + warp + 1);
  const int64_t q0 = __ldg(a.qbase + warp);
  const uint4 *mp = a.meta + (int64_t)t0 * 32 + lane;
  using WQ = typename std::conditional<MODE == 3, uint32_t, uint4>::type;
  const WQ *wbase = (const WQ *)a.words + q0 * 32 + lane;
  __syncthreads();
  const uint32_t bytes = (uint32_t)(n * 8);
  const char *xs_b = (const char *)xs0;
  uint32_t phase = 0;
  double dsum = 0.0;
  const int64_t npairs = (ncols + 1) / 2;
  auto step = [&](double &acc0, double &acc1, uint32_t off, double h) {
    acc0 = fma(h, *(const double *)(xs_b + off), acc0);
    acc1 = fma(h, *(const double *)(xs_b + off + pstride), acc1);
  };
  for (int64_t cp = blockIdx.x; cp < npairs; cp += gridDim.x) {
    const int64_t c0 = 2 * cp;
    const bool two = c0 + 1 < ncols;
    const int64_t c1 = two ? c0 + 1 : c0;  // odd tail: the second plane repeats the column, its results are dropped
    if (threadIdx.x == 0) {
      mbar_expect_tx(bar, 2 * bytes);
      for (int pl = 0; pl < 2; pl++) {
        const char *src = (const char *)(v + (pl ? c1 : c0) * n);
        char *dst = (char *)(pl ? xs1 : xs0);
        for (uint32_t off = 0; off < bytes; off += 32768u) bulk_g2s(dst + off, src + off, min(32768u, bytes - off), bar);
      }
    }
    for (int mu = threadIdx.x; mu < 2 * ndt; mu += blockDim.x) {  // diagonal of both columns per impurity configuration of the row
      const int64_t cg = dg.coloff + (mu < ndt ? c0 : c1);
      const int m = mu < ndt ? mu : mu - ndt;
      double val = __ldg(dg.f_col + cg);
      uint32_t md = (uint32_t)__ldg(dg.map_col + cg) & ((1u << dg.nimp) - 1u);
      while (md) {
        const int b = __ffs(md) - 1;
        md &= md - 1;
        val += __ldg(dg.cross_tab + ((int64_t)b << dg.nimp) + m);
      }
      dtab0[mu] = val;  // dtab1 = dtab0 + ndt
    }
    const WQ *wp = wbase;
    WQ wa = __ldg(wp), wb = __ldg(wp + 32);
    WQ wc = wa, wd = wa;
    if constexpr (MODE == 3) { wc = __ldg(wp + 64); wd = __ldg(wp + 96); }
    uint4 m = t0 < t1 ? __ldg(mp) : make_uint4(0u, 0u, 0xFFFFFFFFu, 0u);
    mbar_wait(bar, phase);
    phase ^= 1u;
    __syncthreads();
    double *o0 = out + c0 * n, *o1 = out + c1 * n;
    for (int t = t0; t < t1; t++) {
      uint4 mnext = make_uint4(0u, 0u, 0xFFFFFFFFu, 0u);
      if (t + 1 < t1) mnext = __ldg(mp + (int64_t)(t + 1 - t0) * 32);
      const int nquad = (int)(m.w >> 16);
      const bool valid = m.z != 0xFFFFFFFFu;
      double acc0 = 0.0, acc1 = 0.0, y0 = 0.0, y1 = 0.0;
      if (a.accum && valid) { y0 = o0[m.z]; y1 = o1[m.z]; }  // requested first, needed last
      if (dg.enabled && valid) {
        const double fr = __hiloint2double((int)m.y, (int)m.x);
        acc0 = (fr + dtab0[m.w & 0xFFFFu]) * xs0[m.z];
        acc1 = (fr + dtab1[m.w & 0xFFFFu]) * xs1[m.z];
      }
      for (int kq = 0; kq < nquad; kq++) {
        const WQ w = wa;
        wp += 32;
        if constexpr (MODE == 3) {
          // two 16-bit words (negative << 15 | class << 14 | source row) per register
          wa = wb; wb = wc; wc = wd;
          wd = __ldg(wp + 96);
          step(acc0, acc1, (w << 3) & 0x1FFF8u, colres_signed((w & 0x4000u) ? a.m1 : a.m0, (w << 16) & 0x80000000u));
          step(acc0, acc1, (w >> 13) & 0x1FFF8u, colres_signed((w & 0x40000000u) ? a.m1 : a.m0, w & 0x80000000u));
        } else {
          // 32-bit words: negative << 31 | byte offset | class
          wa = wb;
          wb = __ldg(wp + 32);
          step(acc0, acc1, w.x & 0x7FFFFFF8u, colres_signed((w.x & 1u) ? a.m1 : a.m0, w.x & 0x80000000u));
          step(acc0, acc1, w.y & 0x7FFFFFF8u, colres_signed((w.y & 1u) ? a.m1 : a.m0, w.y & 0x80000000u));
          step(acc0, acc1, w.z & 0x7FFFFFF8u, colres_signed((w.z & 1u) ? a.m1 : a.m0, w.z & 0x80000000u));
          step(acc0, acc1, w.w & 0x7FFFFFF8u, colres_signed((w.w & 1u) ? a.m1 : a.m0, w.w & 0x80000000u));
        }
      }
      if (valid) {
        acc0 += y0;
        acc1 += y1;
        o0[m.z] = acc0;
        if (two) o1[m.z] = acc1;
        if (a.dot_partial) dsum += xs0[m.z] * acc0 + (two ? xs1[m.z] * acc1 : 0.0);
      }
      m = mnext;
    }
    __syncthreads();  // every gather of this pair is done before the next bulk copies land
  }
  if (a.dot_partial) {  // fixed summation order: lanes -> warps -> CTA, one partial per CTA
    __shared__ double wsum[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dsum += __shfl_down_sync(0xffffffffu, dsum, o);
    if (lane == 0) wsum[warp] = dsum;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); w++) tot += wsum[w];
      a.dot_partial[blockIdx.x] = tot;
    }
  }
}

static size_t colres2_smem(int64_t n, int nimp_diag) {
  const size_t ndt = nimp_diag >= 0 ? ((size_t)1 << nimp_diag) : 0;
  const int64_t npad = (n + 15) / 16 * 16;
  return 128 + ((ndt * 16 + 127) & ~(size_t)127) + 2 * (size_t)(npad + 16) * 8;
}
// kColresNA when the two-column kernel does not apply
static int launch_colres2(const SpinOp &s, int64_t ncols, const double *v, double *out, const DiagArgs &dg, bool accum, bool final) {
  Ctx &c = ctx();
  const Sched &sc = s.sc16x2;
  if (!c.opt.colres_pair || c.opt.colpass_variant != 6 || c.opt.colres_rows > 0) return kColresNA;
  if (!use_tables() || !sc.words || sc.ntask <= 0 || (sc.fmt != 1 && sc.fmt != 2)) return kColresNA;
  if ((s.n & 1) || !c.real_h || ncols < 2) return kColresNA;
  if (dg.enabled && dg.f_row != s.f) return kColresNA;
  const size_t smem = colres2_smem(s.n, dg.enabled ? dg.nimp : -1);
  if (smem > 232448) return kColresNA;
  ColResArgs a{};
  a.accum = accum ? 1 : 0;
  a.tbase = sc.tbase; a.qbase = sc.qbase; a.meta = (const uint4 *)sc.meta; a.words = sc.words;
  a.m0 = s.sc_mag[0]; a.m1 = s.sc_mag[1];
  void (*kern)(int64_t, int64_t, const double *, double *, ColResArgs, DiagArgs) = sc.fmt == 2 ? k_colres2<3> : k_colres2<2>;
  static std::map<const void *, size_t> max_smem;
  size_t &ms = max_smem[(const void *)kern];
  if (smem > ms) {
    CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ms = smem;
  }
  if (sc.nwarps != 32) return kColresNA;
  const int64_t grid = std::min<int64_t>((ncols + 1) / 2, (int64_t)c.sm_count);
  if (final && c.dot_request) {
    if (c.dot_cap < grid) {
      dev_free(c.dot_partial);
      CB_CHECK(dev_alloc(&c.dot_partial, grid));
      c.dot_cap = grid;
    }
    a.dot_partial = c.dot_partial;
    c.dot_npartial = grid;
    c.dot_done = true;
  }
  kern<<<(unsigned)grid, 1024, smem, c.stream>>>(s.n, ncols, v, out, a, dg);
  c.launches++;
  return 0;
}

// diag + Hup on a real vector: column-resident kernel when a column fits in shared memory, else the generic one
int colpass_real(const SpinOp &s, int64_t ncols, const double *v, double *out, const DiagArgs &dg, bool accum, bool final) {
  Ctx &c = ctx();
  if (ncols <= 0 || s.n <= 0) return 0;
  const bool direct = !use_tables();  // matrix-free kernels
  prof_begin(0);
  int rc = kColresNA;
  if (c.opt.colpass_variant == 6) {
    if (c.opt.colres_rows > 0 && !accum) rc = launch_colblk<double>(s, ncols, v, out, dg);
    if (rc == kColresNA) rc = launch_colres2(s, ncols, v, out, dg, accum, final);
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
  const bool direct = !use_tables();  // matrix-free kernels
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
