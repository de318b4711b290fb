// hxv.cu -- the Hamiltonian-times-vector kernels:  Hv = D.v + (1 (x) Hup) v + (Hdw (x) 1) v
// on the DimUp x DimDw sector vector v(iup + idw*DimUp) (iup fastest), complex(8).
//
// Reference paths replaced (all four share this code, selected by mode / rank layout):
//   spMatVec_main            ED_HAMILTONIAN_SPARSE_HxV.f90:167-227     SPARSE, one rank
//   spMatVec_mpi_main        ED_HAMILTONIAN_SPARSE_HxV.f90:230-315     SPARSE, Ndw-sharded
//   directMatVec_main        ED_HAMILTONIAN_DIRECT_HxV.f90:37-90       DIRECT, one rank
//   directMatVec_MPI_main    ED_HAMILTONIAN_DIRECT_HxV.f90:94-171      DIRECT, Ndw-sharded
//   vector_transpose_MPI     ED_HAMILTONIAN_COMMON.f90:30-101          k_transpose_block + NCCL
//
// "column pass": hop matrix acts on the contiguous index (Hup on v, Hdw on the transposed vt);
// "row pass":    hop matrix acts on the strided index (Hdw on v without transposing; one rank).
// Pull (gather) formulation only: every output element is written once, no atomics, so results
// are deterministic run to run.
#include <algorithm>

#include "ctx.h"
#include "hxv_common.cuh"

namespace cb {

// ------------------------------------------------------------------------------------
// Column pass, generic variant (any size): one thread per row i, CB columns per thread so the
// operator row (ELL entries or matrix-free hops) is fetched once per CB outputs.
//   out(i,c) = [diag] d(i,c) v(i,c) + sum_k H(i,j_k) v(j_k,c)          (overwrites out)
// ------------------------------------------------------------------------------------
template <bool REALH, bool DIRECT, int CB>
__global__ void __launch_bounds__(256) k_colpass(int64_t n, int64_t ncols, const double2 *__restrict__ v,
                                                  double2 *__restrict__ out, OpArgs op, DiagArgs dg) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t c0 = (int64_t)blockIdx.y * CB;
  if (i >= n) return;
  double2 acc[CB];
  uint32_t mu_imp = 0;
  if (dg.enabled) mu_imp = (uint32_t)__ldg(dg.map_row + i) & ((1u << dg.nimp) - 1u);
#pragma unroll
  for (int cc = 0; cc < CB; cc++) {
    acc[cc] = make_double2(0.0, 0.0);
    if (dg.enabled && c0 + cc < ncols) {
      double d = diag_value(dg, i, mu_imp, c0 + cc);
      double2 x = ldg2(v + i + (c0 + cc) * n);
      acc[cc] = make_double2(d * x.x, d * x.y);
    }
  }
  if (!DIRECT) {
    const int len = __ldg(op.rowlen + i);
    for (int k = 0; k < len; k++) {
      const int32_t j = __ldg(op.ell_col + (int64_t)k * n + i);
      const double2 h = ldg2(op.ell_val + (int64_t)k * n + i);
#pragma unroll
      for (int cc = 0; cc < CB; cc++)
        if (c0 + cc < ncols) {
          double2 x = ldg2(v + j + (c0 + cc) * n);
          if (REALH) rfma(acc[cc], h.x, x); else cfma(acc[cc], h, x);
        }
    }
  } else {
    const uint32_t s = (uint32_t)__ldg(op.map + i);
    for (int t = 0; t < op.nterms; t++) {
      const Term tm = op.terms[t];
      if (((s >> tm.a) & 1u) && !((s >> tm.b) & 1u)) {
        const uint32_t m = (s & ~(1u << tm.a)) | (1u << tm.b);
        const int32_t j = lin_rank_d(op.lin_lo, op.lin_hi, op.lbits, m);
        const double sg = hop_sign_d(s, tm.a, tm.b);
        const double2 h = make_double2(tm.re * sg, tm.im * sg);
#pragma unroll
        for (int cc = 0; cc < CB; cc++)
          if (c0 + cc < ncols) {
            double2 x = ldg2(v + j + (c0 + cc) * n);
            if (REALH) rfma(acc[cc], h.x, x); else cfma(acc[cc], h, x);
          }
      }
    }
  }
#pragma unroll
  for (int cc = 0; cc < CB; cc++)
    if (c0 + cc < ncols) out[i + (c0 + cc) * n] = acc[cc];
}

// ------------------------------------------------------------------------------------
// Row pass (one rank, no transpose): out(i,c) += sum_k Hd(c,j_k) v(i,j_k).
// Threads run along i (contiguous), the operator row of column c is warp-uniform (broadcast
// loads).  Grid x = columns (fastest) inside one slab of rows, so a slab (rows x all columns)
// stays L2-resident while every column of it is produced.
// ------------------------------------------------------------------------------------
template <bool REALH, bool DIRECT>
__global__ void __launch_bounds__(256) k_rowpass(int64_t n /*rows=DimUp*/, int64_t ncols /*DimDw*/,
                                                  const double2 *__restrict__ v, double2 *__restrict__ out,
                                                  const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                                                  const double2 *__restrict__ val, OpArgs op) {
  const int64_t c = blockIdx.x;
  const int64_t i = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double2 acc = make_double2(0.0, 0.0);
  if (!DIRECT) {
    const int32_t p0 = __ldg(rowptr + c), p1 = __ldg(rowptr + c + 1);
    int32_t p = p0;
    for (; p + 4 <= p1; p += 4) {
      double2 x0 = ldg2(v + i + (int64_t)__ldg(col + p) * n);
      double2 x1 = ldg2(v + i + (int64_t)__ldg(col + p + 1) * n);
      double2 x2 = ldg2(v + i + (int64_t)__ldg(col + p + 2) * n);
      double2 x3 = ldg2(v + i + (int64_t)__ldg(col + p + 3) * n);
      double2 h0 = ldg2(val + p), h1 = ldg2(val + p + 1), h2 = ldg2(val + p + 2), h3 = ldg2(val + p + 3);
      if (REALH) { rfma(acc, h0.x, x0); rfma(acc, h1.x, x1); rfma(acc, h2.x, x2); rfma(acc, h3.x, x3); }
      else { cfma(acc, h0, x0); cfma(acc, h1, x1); cfma(acc, h2, x2); cfma(acc, h3, x3); }
    }
    for (; p < p1; p++) {
      double2 x = ldg2(v + i + (int64_t)__ldg(col + p) * n);
      double2 h = ldg2(val + p);
      if (REALH) rfma(acc, h.x, x); else cfma(acc, h, x);
    }
  } else {
    const uint32_t s = (uint32_t)__ldg(op.map + c);
    for (int t = 0; t < op.nterms; t++) {
      const Term tm = op.terms[t];
      if (((s >> tm.a) & 1u) && !((s >> tm.b) & 1u)) {
        const uint32_t m = (s & ~(1u << tm.a)) | (1u << tm.b);
        const int64_t j = lin_rank_d(op.lin_lo, op.lin_hi, op.lbits, m);
        const double sg = hop_sign_d(s, tm.a, tm.b);
        double2 x = ldg2(v + i + j * n);
        if (REALH) rfma(acc, tm.re * sg, x); else cfma(acc, make_double2(tm.re * sg, tm.im * sg), x);
      }
    }
  }
  double2 o = out[i + c * n];
  o.x += acc.x;
  o.y += acc.y;
  out[i + c * n] = o;
}

// ------------------------------------------------------------------------------------
// Row pass, RB row chunks per thread (SPARSE): the warp-uniform operator entry (column index +
// coefficient = 2 L1 wavefronts) is fetched once per RB coalesced gathers instead of once per gather,
// and the RB gathers of an entry are independent loads in flight.
// ------------------------------------------------------------------------------------
// DOT: also reduce Re<v, out> over the CTA's outputs (out is final after this pass) into dot_partial[CTA] --
// the Lanczos alpha without a separate sweep over two vectors (fixed summation order: deterministic).
template <bool REALH, int RB, bool DOT>
__global__ void __launch_bounds__(256) k_rowpass_rb(int64_t n /*rows=DimUp*/, const double2 *__restrict__ v,
                                                     double2 *__restrict__ out, const int32_t *__restrict__ rowptr,
                                                     const int32_t *__restrict__ col, const double2 *__restrict__ val,
                                                     double *__restrict__ dot_partial) {
  const int64_t c = blockIdx.x;
  const int64_t i0 = (int64_t)blockIdx.y * (blockDim.x * RB) + threadIdx.x;
  double2 acc[RB];
  const double2 *vi[RB];
#pragma unroll
  for (int r = 0; r < RB; r++) {
    acc[r] = make_double2(0.0, 0.0);
    vi[r] = v + min(i0 + (int64_t)r * blockDim.x, n - 1);  // clamped: rows past the end are not stored
  }
  const int32_t p0 = __ldg(rowptr + c), p1 = __ldg(rowptr + c + 1);
  for (int32_t p = p0; p < p1; p++) {
    const int64_t off = (int64_t)__ldg(col + p) * n;
    const double2 h = ldg2(val + p);
    double2 x[RB];
#pragma unroll
    for (int r = 0; r < RB; r++) x[r] = ldg2(vi[r] + off);
    // the coefficient is warp-uniform: purely real / purely imaginary entries (every hop of the BHZ model) take
    // two FMAs instead of four
    if (REALH || h.y == 0.0) {
#pragma unroll
      for (int r = 0; r < RB; r++) rfma(acc[r], h.x, x[r]);
    } else if (h.x == 0.0) {
#pragma unroll
      for (int r = 0; r < RB; r++) {
        acc[r].x = fma(-h.y, x[r].y, acc[r].x);
        acc[r].y = fma(h.y, x[r].x, acc[r].y);
      }
    } else {
#pragma unroll
      for (int r = 0; r < RB; r++) cfma(acc[r], h, x[r]);
    }
  }
  double dsum = 0.0;
#pragma unroll
  for (int r = 0; r < RB; r++) {
    const int64_t i = i0 + (int64_t)r * blockDim.x;
    if (i < n) {
      double2 *o = out + i + c * n;
      double2 y = *o;
      y.x += acc[r].x;
      y.y += acc[r].y;
      *o = y;
      if (DOT) {
        const double2 u = ldg2(v + i + c * n);
        dsum = fma(u.x, y.x, fma(u.y, y.y, dsum));
      }
    }
  }
  if (DOT) {
    __shared__ double wsum[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dsum += __shfl_down_sync(0xffffffffu, dsum, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = dsum;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
      for (int w = 0; w < (int)((blockDim.x + 31) >> 5); w++) tot += wsum[w];
      dot_partial[(int64_t)blockIdx.y * gridDim.x + blockIdx.x] = tot;
    }
  }
}

// ------------------------------------------------------------------------------------
// Tiled transpose of a sub-block: dst[(dcol_off + c) + r*ld_dst] (=|+=) src[(srow_off + r) + c*ld_src]
// for r in [0,nr), c in [0,nc).  This is pack + exchange + unpack + local_transpose of
// vector_transpose_MPI in one kernel when dst is the destination rank's buffer.
// ------------------------------------------------------------------------------------
template <bool ACCUM>
__global__ void __launch_bounds__(256) k_transpose_block(const double2 *__restrict__ src, int64_t ld_src, int64_t srow_off,
                                                          int64_t nr, int64_t nc, double2 *__restrict__ dst,
                                                          int64_t ld_dst, int64_t dcol_off) {
  __shared__ double2 tile[32][33];
  const int64_t r0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    int64_t r = r0 + tx, c = c0 + ty + k;
    if (r < nr && c < nc) tile[ty + k][tx] = src[(srow_off + r) + c * ld_src];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    int64_t c = c0 + tx, r = r0 + ty + k;
    if (r < nr && c < nc) {
      double2 x = tile[tx][ty + k];
      double2 *d = dst + (dcol_off + c) + r * ld_dst;
      if (ACCUM) { double2 o = *d; x.x += o.x; x.y += o.y; }
      *d = x;
    }
  }
}

// strided block copy: dst[(doff + c) + r*ld_dst] (=|+=) src[c + r*nc]   (unpack of an NCCL block)
template <bool ACCUM>
__global__ void __launch_bounds__(256) k_copy_block(const double2 *__restrict__ src, int64_t nr, int64_t nc,
                                                     double2 *__restrict__ dst, int64_t ld_dst, int64_t doff) {
  const int64_t total = nr * nc;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = k / nc, c = k - r * nc;
    double2 x = src[k];
    double2 *d = dst + (doff + c) + r * ld_dst;
    if (ACCUM) { double2 o = *d; x.x += o.x; x.y += o.y; }
    *d = x;
  }
}

// Multi-destination tiled transpose: the rows r (contiguous index of src) are partitioned into up to 16 segments
// [b[p], b[p+1]) -- one per destination rank -- and every segment has its own destination:
//   dst_p[(dcol_off_p + c) + (r - b[p]) * ld_p] (=|+=) src[r + c*ld_src]        r in segment p, c in [0, nc)
// One launch packs every block of a distributed transpose (and writes the own block in place), instead of one
// small launch per rank: at 8 ranks the product is otherwise bound by the launch rate, not by any kernel.
constexpr int kMaxSeg = 16;
struct XposeSeg {
  double2 *dst;
  int64_t ld, dcol_off;
  int32_t accum, pad;
};
struct XposeArgs {
  int64_t b[kMaxSeg + 1];
  XposeSeg seg[kMaxSeg];
  int nseg;
};
__global__ void __launch_bounds__(256) k_xpose_multi(const double2 *__restrict__ src, int64_t ld_src, int64_t nr, int64_t nc, XposeArgs a) {
  __shared__ double2 tile[32][33];
  const int64_t r0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int64_t r = r0 + tx, c = c0 + ty + k;
    if (r < nr && c < nc) tile[ty + k][tx] = src[r + c * ld_src];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int64_t c = c0 + tx, r = r0 + ty + k;
    if (r < nr && c < nc) {
      int p = 0;
      while (p + 1 < a.nseg && r >= a.b[p + 1]) p++;
      double2 x = tile[tx][ty + k];
      double2 *d = a.seg[p].dst + (a.seg[p].dcol_off + c) + (r - a.b[p]) * a.seg[p].ld;
      if (a.seg[p].accum) { const double2 o = *d; x.x += o.x; x.y += o.y; }
      *d = x;
    }
  }
}
// Fused unpack of the way back: hv(i, r) += the element sender p packed for it, for every row i outside my own
// up-range.  The receive window holds, per sender p and chunk ch of its up-rows, a block [i_chunk + r * q_chunk]
// at q_dw * (up_off(p) + chunk_off).  ub = up-split boundaries (nseg + 1), nch = chunks per sender.
struct UnpackArgs {
  int64_t ub[kMaxSeg + 1];
  int nseg, me, nch_opt;
};
__global__ void __launch_bounds__(256) k_unpack_multi(const double2 *__restrict__ recv, double2 *__restrict__ hv, int64_t DU, int64_t qdw,
                                                      UnpackArgs a) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t r = blockIdx.y;
  if (i >= DU || r >= qdw) return;
  int p = 0;
  while (p + 1 < a.nseg && i >= a.ub[p + 1]) p++;
  if (p == a.me) return;
  const int64_t uoff = a.ub[p], uq = a.ub[p + 1] - a.ub[p];
  const int64_t nch = uq < a.nch_opt ? (uq > 1 ? uq : 1) : a.nch_opt;  // = max(1, min(xchg_chunks, uq)) of the sender
  // split_of(uq, nch, ch): the first (uq mod nch) chunks hold one more row
  const int64_t q = uq / nch, rem = uq % nch, il = i - uoff;
  int64_t ch, coff, cq;
  if (il < rem * (q + 1)) { ch = il / (q + 1); coff = ch * (q + 1); cq = q + 1; }
  else { ch = rem + (il - rem * (q + 1)) / (q > 0 ? q : 1); coff = ch * q + rem; cq = q; }
  const double2 x = recv[qdw * (uoff + coff) + (il - coff) + r * cq];
  double2 *d = hv + i + r * DU;
  double2 o = *d;
  o.x += x.x;
  o.y += x.y;
  *d = o;
}

// ------------------------------------------------------------------------------------
OpArgs op_args(const SpinOp &s) {
  Ctx &c = ctx();
  OpArgs o{};
  o.ell_col = s.ell_col; o.ell_val = s.ell_val; o.rowlen = s.rowlen; o.ell_w = s.ell_w;
  o.map = s.map; o.lin_lo = s.lin_lo; o.lin_hi = s.lin_hi; o.lbits = c.ns / 2;
  o.terms = s.terms; o.nterms = s.nterms;
  return o;
}

template <int CB>
static int launch_colpass(const SpinOp &s, int64_t ncols, const double2 *v, double2 *out, const DiagArgs &dg) {
  Ctx &c = ctx();
  if (ncols <= 0 || s.n <= 0) return 0;
  dim3 grid((unsigned)((s.n + 255) / 256), (unsigned)((ncols + CB - 1) / CB));
  if (grid.y > 65535) return fail("colpass: too many column groups (%u)", grid.y);
  OpArgs op = op_args(s);
  const bool direct = !use_tables();  // matrix-free kernels
  if (c.real_h) {
    if (direct) k_colpass<true, true, CB><<<grid, 256, 0, c.stream>>>(s.n, ncols, v, out, op, dg);
    else k_colpass<true, false, CB><<<grid, 256, 0, c.stream>>>(s.n, ncols, v, out, op, dg);
  } else {
    if (direct) k_colpass<false, true, CB><<<grid, 256, 0, c.stream>>>(s.n, ncols, v, out, op, dg);
    else k_colpass<false, false, CB><<<grid, 256, 0, c.stream>>>(s.n, ncols, v, out, op, dg);
  }
  c.launches++;
  return 0;
}

// tile-resident row pass (rowtile.cu)
bool rowtile_applicable(const SpinOp &s);
int launch_rowtile(const SpinOp &s, int64_t nrows, const double2 *v, double2 *out, bool accum);

// out (=|+=) [diag] D.v + H_s v with H_s on the contiguous index.  accum is only available from the whole-column
// kernel (callers check colres_applicable first); final marks the last contribution to H x v.
static int colpass_impl(const SpinOp &s, int64_t ncols, const double2 *v, double2 *out, const DiagArgs &dg, bool accum, bool final);
static int colpass(const SpinOp &s, int64_t ncols, const double2 *v, double2 *out, const DiagArgs &dg, bool accum = false,
                   bool final = false) {
  prof_begin(0);
  int rc = colpass_impl(s, ncols, v, out, dg, accum, final);
  prof_end();
  return rc;
}
static int colpass_impl(const SpinOp &s, int64_t ncols, const double2 *v, double2 *out, const DiagArgs &dg, bool accum, bool final) {
  if (ncols <= 0 || s.n <= 0) return 0;
  // variant 6 = column-resident shared-memory kernels (default), 1 = generic global-gather kernel
  if (ctx().opt.colpass_variant == 6) {
    // forced block split (tests) -> block-split kernel; else whole column; else block split for big columns
    int rc = (ctx().opt.colres_rows > 0 && !accum) ? launch_colblk<double2>(s, ncols, v, out, dg) : kColresNA;
    if (rc == kColresNA) rc = launch_colres<double2>(s, ncols, v, out, dg, accum, final);
    if (rc == kColresNA && !accum) rc = launch_colblk<double2>(s, ncols, v, out, dg);
    if (rc != kColresNA) return rc;
  }
  if (accum) return fail("internal: accumulating column pass without the column-resident kernel");
  // DIRECT mode, or a column larger than shared memory without block schedules: generic kernel
  switch (ctx().opt.col_batch) {
    case 1: return launch_colpass<1>(s, ncols, v, out, dg);
    case 2: return launch_colpass<2>(s, ncols, v, out, dg);
    case 8: return launch_colpass<8>(s, ncols, v, out, dg);
    default: return launch_colpass<4>(s, ncols, v, out, dg);
  }
}

// out (=|+=) v . H_s^T with H_s on the strided index.  accum = false (plain store) only from the tile-resident
// kernel (callers check rowtile_applicable first); the generic kernels always accumulate.
static int rowpass_impl(const SpinOp &s, int64_t nrows, const double2 *v, double2 *out, bool accum);
static int rowpass(const SpinOp &s, int64_t nrows, const double2 *v, double2 *out, bool accum = true) {
  prof_begin(1);
  int rc = rowpass_impl(s, nrows, v, out, accum);
  prof_end();
  return rc;
}
static int rowpass_impl(const SpinOp &s, int64_t nrows, const double2 *v, double2 *out, bool accum) {
  Ctx &c = ctx();
  if (s.n <= 0 || nrows <= 0) return 0;
  {
    const int rc = launch_rowtile(s, nrows, v, out, accum);
    if (rc != kColresNA) return rc;
  }
  if (!accum) return fail("internal: storing row pass without the tile-resident kernel");
  dim3 grid((unsigned)s.n, (unsigned)((nrows + 255) / 256));
  if (grid.y > 65535) return fail("rowpass: too many row chunks");
  OpArgs op = op_args(s);
  const bool direct = !use_tables();  // matrix-free kernels
  if (!direct && (c.opt.row_rb > 1 || c.opt.row_slab != 256)) {
    // RB row chunks per thread; the slab of rows one grid.y index sweeps (threads*RB rows x all columns) must stay
    // L2-resident: 128 rows x 12870 columns x 16 B = 26 MB at K3 -> threads = row_slab / RB, at most 256
    const int rb = c.opt.row_rb >= 4 ? 4 : (c.opt.row_rb >= 2 ? 2 : 1);
    const int slab = (int)std::max<int64_t>(64, std::min<int64_t>(256 * rb, c.opt.row_slab));
    const int thr = std::min(256, std::max(32, slab / rb / 32 * 32));
    dim3 g2((unsigned)s.n, (unsigned)((nrows + (int64_t)thr * rb - 1) / ((int64_t)thr * rb)));
    if (g2.y > 65535) return fail("rowpass: too many row chunks");
    // fused Lanczos dot (lanczos.cu sets dot_request): one partial per CTA; only when this pass is the last one
    double *dp = nullptr;
    if (c.dot_request && c.dot_final_rowpass) {
      const int64_t np = (int64_t)g2.x * g2.y;
      if (c.dot_cap < np) {
        dev_free(c.dot_partial);
        CB_CHECK(dev_alloc(&c.dot_partial, np));
        c.dot_cap = np;
      }
      dp = c.dot_partial;
      c.dot_npartial = np;
      c.dot_done = true;
    }
#define CB_ROWRB(RH_, RB_)                                                                                                   \
  do {                                                                                                                       \
    if (dp) k_rowpass_rb<RH_, RB_, true><<<g2, thr, 0, c.stream>>>(nrows, v, out, s.rowptr, s.col, s.val, dp);                \
    else k_rowpass_rb<RH_, RB_, false><<<g2, thr, 0, c.stream>>>(nrows, v, out, s.rowptr, s.col, s.val, nullptr);             \
  } while (0)
    if (c.real_h) {
      if (rb == 4) CB_ROWRB(true, 4); else if (rb == 2) CB_ROWRB(true, 2); else CB_ROWRB(true, 1);
    } else {
      if (rb == 4) CB_ROWRB(false, 4); else if (rb == 2) CB_ROWRB(false, 2); else CB_ROWRB(false, 1);
    }
#undef CB_ROWRB
    c.launches++;
    return 0;
  }
  if (c.real_h) {
    if (direct) k_rowpass<true, true><<<grid, 256, 0, c.stream>>>(nrows, s.n, v, out, s.rowptr, s.col, s.val, op);
    else k_rowpass<true, false><<<grid, 256, 0, c.stream>>>(nrows, s.n, v, out, s.rowptr, s.col, s.val, op);
  } else {
    if (direct) k_rowpass<false, true><<<grid, 256, 0, c.stream>>>(nrows, s.n, v, out, s.rowptr, s.col, s.val, op);
    else k_rowpass<false, false><<<grid, 256, 0, c.stream>>>(nrows, s.n, v, out, s.rowptr, s.col, s.val, op);
  }
  c.launches++;
  return 0;
}

// Row pass of Hdw on a REAL vector with an even number of rows: adjacent rows (i, i+1) of the real vector are
// one double2 of a "complex" vector with nrows/2 rows, and a real coefficient acts on both halves alike, so the
// complex kernels run unchanged on half as many rows: every gather moves 16 bytes per lane instead of 8.
int rowpass_real_as_pairs(int64_t nrows, const double *v, double *out, bool accum) {
  Ctx &c = ctx();
  if ((nrows & 1) || !c.real_h) return fail("rowpass_real_as_pairs: needs a real H and an even DimUp");
  return rowpass(c.dw, nrows / 2, (const double2 *)v, (double2 *)out, accum);
}

template <bool ACCUM>
static void transpose_block(const double2 *src, int64_t ld_src, int64_t srow_off, int64_t nr, int64_t nc, double2 *dst,
                            int64_t ld_dst, int64_t dcol_off) {
  Ctx &c = ctx();
  if (nr <= 0 || nc <= 0) return;
  dim3 grid((unsigned)((nr + 31) / 32), (unsigned)((nc + 31) / 32));
  k_transpose_block<ACCUM><<<grid, 256, 0, c.stream>>>(src, ld_src, srow_off, nr, nc, dst, ld_dst, dcol_off);
  c.launches++;
}
template <bool ACCUM>
static void copy_block(const double2 *src, int64_t nr, int64_t nc, double2 *dst, int64_t ld_dst, int64_t doff) {
  Ctx &c = ctx();
  if (nr <= 0 || nc <= 0) return;
  int64_t nb = std::min<int64_t>((nr * nc + 255) / 256, (int64_t)c.sm_count * 16);
  k_copy_block<ACCUM><<<(unsigned)nb, 256, 0, c.stream>>>(src, nr, nc, dst, ld_dst, doff);
  c.launches++;
}

DiagArgs diag_args(int64_t coloff) {
  Ctx &c = ctx();
  DiagArgs d{};
  d.enabled = c.kin_only ? 0 : 1;  // kin_only: the operator is the bare hopping part (cdmft_b200_imp_kinetic)
  d.f_row = c.up.f; d.f_col = c.dw.f;
  d.map_row = c.up.map; d.map_col = c.dw.map;
  d.cross_tab = c.cross_tab;
  d.nimp = c.nimp;
  d.coloff = coloff;
  return d;
}

// H x v on the local shard(s); v/hv device pointers laid out as the concatenation of the local
// ranks' shards (exactly one shard in SPMD mode, all P in sim mode, the whole vector otherwise).
// ------------------------------------------------------------------------------------
// Non-local Kanamori terms (Jhflag: Norb>1 and Jx/=0 or Jp/=0): spin-exchange and pair-hopping,
// ED_HAMILTONIAN/sparse/H_non_local.f90:4-100 == direct/HxV_non_local.f90:4-86.  One thread per
// local row i=(iup,idw): conditions on the row state, column j = (c^+_is c_js)_up (c^+_js c_is)_dw |i>
// (S-E) or (c^+_is c_js)_up (c^+_is c_js)_dw |i> (P-H), Hv(i) += J * sg1 sg2 sg3 sg4 * v(j).
// v is the FULL vector (the reference all-gathers it, ED_HAMILTONIAN_SPARSE_HxV.f90:302-305).
// ------------------------------------------------------------------------------------
struct NonLocalArgs {
  const int32_t *map_up, *map_dw;
  const int32_t *up_lo, *up_hi, *dw_lo, *dw_hi;
  int lbits, nlat, norb;
  double jx, jp;
  int64_t dimup;
  int64_t row0;  // global index of the first local row
};
__device__ __forceinline__ double hop_sign2(uint32_t s, int a, int b) { return hop_sign_d(s, a, b); }

__global__ void __launch_bounds__(256) k_nonlocal(int64_t nloc, const double2 *__restrict__ vfull, double2 *__restrict__ hv,
                                                   NonLocalArgs a) {
  const int64_t il = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (il >= nloc) return;
  const int64_t ig = il + a.row0;
  const int64_t iup = ig % a.dimup, idw = ig / a.dimup;
  const uint32_t mup = (uint32_t)__ldg(a.map_up + iup), mdw = (uint32_t)__ldg(a.map_dw + idw);
  double2 acc = make_double2(0.0, 0.0);
  for (int ilat = 0; ilat < a.nlat; ilat++)
    for (int io = 0; io < a.norb; io++)
      for (int jo = 0; jo < a.norb; jo++) {
        if (io == jo) continue;
        const int is = io + ilat * a.norb, js = jo + ilat * a.norb;  // imp_state_index - 1
        const uint32_t bi = 1u << is, bj = 1u << js;
        const bool nup_i = mup & bi, nup_j = mup & bj, ndw_i = mdw & bi, ndw_j = mdw & bj;
        // S-E: nup(j)=1, ndw(i)=1, ndw(j)=0, nup(i)=0 : dw c(is) cdg(js) ; up c(js) cdg(is)
        if (a.jx != 0.0 && nup_j && ndw_i && !ndw_j && !nup_i) {
          const uint32_t kdw = (mdw & ~bi) | bj, kup = (mup & ~bj) | bi;
          const double sg = hop_sign2(mdw, js, is) * hop_sign2(mup, is, js);
          const int64_t jdw = lin_rank_d(a.dw_lo, a.dw_hi, a.lbits, kdw), jup = lin_rank_d(a.up_lo, a.up_hi, a.lbits, kup);
          const double2 x = ldg2(vfull + jup + jdw * a.dimup);
          acc.x += a.jx * sg * x.x;
          acc.y += a.jx * sg * x.y;
        }
        // P-H: nup(j)=1, ndw(j)=1, ndw(i)=0, nup(i)=0 : dw c(js) cdg(is) ; up c(js) cdg(is)
        if (a.jp != 0.0 && nup_j && ndw_j && !ndw_i && !nup_i) {
          const uint32_t kdw = (mdw & ~bj) | bi, kup = (mup & ~bj) | bi;
          const double sg = hop_sign2(mdw, is, js) * hop_sign2(mup, is, js);
          const int64_t jdw = lin_rank_d(a.dw_lo, a.dw_hi, a.lbits, kdw), jup = lin_rank_d(a.up_lo, a.up_hi, a.lbits, kup);
          const double2 x = ldg2(vfull + jup + jdw * a.dimup);
          acc.x += a.jp * sg * x.x;
          acc.y += a.jp * sg * x.y;
        }
      }
  double2 o = hv[il];
  o.x += acc.x;
  o.y += acc.y;
  hv[il] = o;
}

static int hxv_local_terms(const double2 *v, double2 *hv, bool pairs);
// allgather_vector_MPI (ED_SETUP.f90:672-708): every active rank sends its shard to every active rank; *out = the full
// sector vector on this rank (c.vfull, allocated on first use).  One process (single rank, simulated ranks): v itself.
int allgather_full(const double2 *v, const double2 **out) {
  Ctx &c = ctx();
  *out = v;
  if (!(c.spmd && c.p_eff > 1) || c.rk.empty()) return 0;
  if (!c.vfull) CB_CHECK(dev_alloc(&c.vfull, c.dim));
  std::vector<int64_t> cs(c.nranks, 0), os(c.nranks, 0), cr(c.nranks, 0), orr(c.nranks, 0);
  for (int p = 0; p < c.p_eff; p++) {
    Split pd = split_of(c.dimdw, c.p_eff, p);
    cs[p] = c.rk[0].nloc; os[p] = 0;
    cr[p] = pd.q * c.dimup; orr[p] = pd.off * c.dimup;
  }
  prof_begin(3);
  CB_CHECK(nccl_all_to_all(v, c.vfull, cs.data(), os.data(), cr.data(), orr.data()));
  prof_end();
  *out = c.vfull;
  return 0;
}

int hxv_device(const double2 *v, double2 *hv) {
  Ctx &c = ctx();
  CB_CHECK(hxv_local_terms(v, hv, false));
  if (!c.jhflag) return 0;
  NonLocalArgs a{};
  a.map_up = c.up.map; a.map_dw = c.dw.map;
  a.up_lo = c.up.lin_lo; a.up_hi = c.up.lin_hi; a.dw_lo = c.dw.lin_lo; a.dw_hi = c.dw.lin_hi;
  a.lbits = c.ns / 2; a.nlat = c.m.nlat; a.norb = c.m.norb; a.jx = c.m.jx; a.jp = c.m.jp; a.dimup = c.dimup;
  const double2 *vfull = v;
  if (c.spmd && c.p_eff > 1) {
    if (c.rk.empty()) return 0;
    CB_CHECK(allgather_full(v, &vfull));
  }
  int64_t off = 0;
  for (auto &r : c.rk) {
    if (r.nloc > 0) {
      a.row0 = r.dw.off * c.dimup;
      // single process (one rank or simulated ranks): v is already the gathered vector
      k_nonlocal<<<(unsigned)((r.nloc + 255) / 256), 256, 0, c.stream>>>(r.nloc, vfull, hv + off, a);
      c.launches++;
    }
    off += r.nloc;
  }
  return 0;
}

// ------------------------------------------------------------------------------------
// Copy-engine exchange (SPMD with CUDA-IPC peer windows): the two distributed transposes of spMatVec_mpi_main /
// directMatVec_MPI_main (vector_transpose_MPI, ED_HAMILTONIAN_COMMON.f90:30-94) without a single SM cycle spent
// on NVLink traffic:
//   way out : k_transpose_block packs my columns, already transposed, one block per destination; the DMA engines
//             copy every block straight into the owner's vt (2-D copy: rows of my q_dw elements at pitch DimDw),
//             one stream per peer, all peers at once -- while the SMs run the diag + Hup column pass;
//   way back: the Hdw pass on vt runs in chunks of my up-rows; behind every chunk its transposed blocks are
//             packed and handed to the DMA engines while the next chunk computes; the owners add the received
//             blocks to Hv (k_copy_block, accumulating);
//   two stream-ordered barriers per product (all blocks of every vt have landed / every receive window is
//   complete); the one of product n also tells everybody that vt and the windows of product n-1 are free again.
// v / hv: this rank's shard, 16-byte elements (complex, or the paired-row view of a real vector), DU rows.
// ------------------------------------------------------------------------------------
int colpass_real(const SpinOp &s, int64_t ncols, const double *v, double *out, const DiagArgs &dg, bool accum, bool final);  // hxv_real.cu
// chunks of the Hdw pass pipelined against the way back.  A peer-to-peer DMA copy carries ~20 us of fixed cost
// (measured on the B200 box, tools/p2p_bw.py: 8 MB 280 GB/s, 40 MB 566, 83 MB 664, 331 MB 745, 1 GB 772), and
// copies to different peers do not run faster side by side, so small blocks are expensive, while large ones leave a
// long unhidden tail: about 96 MB per copy, between 2 and 6 chunks (K3: 4 chunks at 2 ranks, 2 at 8; K5 at 8 ranks: 6;
// measured: K3/8 ranks 2.01 | 2.04 | 2.21 ms with 2 | 4 | 8 chunks, K5 32.0 | 30.1 ms with 2 | 4).
// Option xchg_chunks > 0 overrides.  Every rank computes the same number (it enters the receive-window layout).
static int64_t xchg_chunks_eff(int64_t DU) {
  Ctx &c = ctx();
  if (c.opt.xchg_chunks > 0) return c.opt.xchg_chunks;
  const int P = c.p_eff;
  const int64_t block_bytes = (c.dimdw / P + 1) * (DU / P + 1) * 16;  // one sender -> one receiver, one direction
  return std::max<int64_t>(2, std::min<int64_t>(6, (block_bytes + (96ll << 20) - 1) / (96ll << 20)));
}
static int hxv_sharded_ce(const double2 *v, double2 *hv, bool pairs, int64_t DU) {
  Ctx &c = ctx();
  RankState &me = c.rk[0];
  const int P = c.p_eff;
  auto usplit = [&](int p) { return split_of(DU, P, p); };
  const Split me_up = usplit(me.rank);
  cudaStream_t S = c.stream;
  const int nsplit = (int)std::max<int64_t>(1, std::min<int64_t>(c.opt.xchg_split, 4));  // DMA streams per peer
  if ((int)c.peer_stream.size() < P * nsplit) {
    int lo = 0, hi = 0;
    CB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    while ((int)c.peer_stream.size() < P * nsplit) {
      cudaStream_t st;
      cudaEvent_t ev;
      CB_CUDA(cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, hi));
      CB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      c.peer_stream.push_back(st);
      c.peer_done.push_back(ev);
    }
    if (!c.ev_pack) CB_CUDA(cudaEventCreateWithFlags(&c.ev_pack, cudaEventDisableTiming));
  }
  DiagArgs nodiag{};
  // ---- way out: pack (transposing), DMA into the owners' vt, diag + Hup meanwhile
  std::vector<int64_t> os(P, 0);
  {
    int64_t so = 0;
    for (int p = 0; p < P; p++) { os[p] = so; so += usplit(p).q * me.dw.q; }
  }
  if (P > kMaxSeg) return fail("copy-engine exchange: more than %d ranks", kMaxSeg);
  prof_begin(2);
  if (me.dw.q > 0 && DU > 0) {  // one launch: every destination's block, transposed (the own block straight into vt)
    XposeArgs xa{};
    xa.nseg = P;
    for (int p = 0; p < P; p++) {
      const Split pu = usplit(p);
      xa.b[p] = pu.off;
      xa.b[p + 1] = pu.off + pu.q;
      if (p == me.rank) xa.seg[p] = XposeSeg{me.vt, c.dimdw, me.dw.off, 0, 0};
      else xa.seg[p] = XposeSeg{me.sendbuf + os[p], me.dw.q, 0, 0, 0};
    }
    dim3 grid((unsigned)((DU + 31) / 32), (unsigned)((me.dw.q + 31) / 32));
    k_xpose_multi<<<grid, 256, 0, S>>>(v, DU, DU, me.dw.q, xa);
    c.launches++;
  }
  prof_end();
  CB_CUDA(cudaEventRecord(c.ev_pack, S));
  for (int k = 1; k < P; k++) {
    const int p = (me.rank + k) % P;  // staggered: at step k every rank targets a different GPU
    const Split pu = usplit(p);
    if (pu.q <= 0 || me.dw.q <= 0) continue;
    for (int j = 0; j < nsplit; j++) {  // rows [r.off, r.off + r.q) of the block on stream j of this peer
      const Split r = split_of(pu.q, nsplit, j);
      if (r.q <= 0) continue;
      cudaStream_t st = c.peer_stream[p * nsplit + j];
      CB_CUDA(cudaStreamWaitEvent(st, c.ev_pack, 0));
      CB_CUDA(cudaMemcpy2DAsync(c.peer_vt[p] + me.dw.off + r.off * c.dimdw, (size_t)c.dimdw * 16, me.sendbuf + os[p] + r.off * me.dw.q,
                                (size_t)me.dw.q * 16, (size_t)me.dw.q * 16, (size_t)r.q, cudaMemcpyDeviceToDevice, st));
      CB_CUDA(cudaEventRecord(c.peer_done[p * nsplit + j], st));
    }
  }
  if (pairs) CB_CHECK(colpass_real(c.up, me.dw.q, (const double *)v, (double *)hv, diag_args(me.dw.off), false, false));
  else CB_CHECK(colpass(c.up, me.dw.q, v, hv, diag_args(me.dw.off)));
  prof_begin(3);  // what is left of my outgoing DMA copies (way out) after the overlap with the column pass
  for (int k = 1; k < P; k++)
    for (int j = 0; j < nsplit; j++)
      if (split_of(usplit((me.rank + k) % P).q, nsplit, j).q > 0) CB_CUDA(cudaStreamWaitEvent(S, c.peer_done[((me.rank + k) % P) * nsplit + j], 0));
  prof_end();
  prof_begin(5);  // barrier: all blocks of every vt have landed (includes waiting for the slowest rank)
  CB_CHECK(nccl_barrier());
  prof_end();
  // ---- Hdw on vt in chunks of my up-rows; way back pipelined behind the chunks
  const int nch = (int)std::max<int64_t>(1, std::min<int64_t>(xchg_chunks_eff(DU), me_up.q));
  std::vector<int64_t> ob(P, 0);  // receive-window layout of peer p: block of sender s at q_dw(p) * up_off(s)
  {
    int64_t so = 0;
    for (int p = 0; p < P; p++) { ob[p] = so; so += split_of(c.dimdw, P, p).q * me_up.q; }
  }
  for (int ch = 0; ch < nch; ch++) {
    const Split cs = split_of(me_up.q, nch, ch);  // rows [cs.off, cs.off + cs.q) of my up range
    if (cs.q <= 0) continue;
    CB_CHECK(colpass(c.dw, cs.q, me.vt + cs.off * c.dimdw, me.hvt + cs.off * c.dimdw, nodiag));
    prof_begin(2);
    {  // one launch: p's columns (rows of hvt) x this chunk of my up-rows, transposed [iup_chunk + idw_p * cs.q]; own block += into hv
      XposeArgs xa{};
      xa.nseg = P;
      for (int p = 0; p < P; p++) {
        const Split pd = split_of(c.dimdw, P, p);
        xa.b[p] = pd.off;
        xa.b[p + 1] = pd.off + pd.q;
        if (p == me.rank) xa.seg[p] = XposeSeg{hv, DU, me_up.off + cs.off, 1, 0};
        else xa.seg[p] = XposeSeg{me.sendbuf + ob[p] + pd.q * cs.off, cs.q, 0, 0, 0};
      }
      dim3 grid((unsigned)((c.dimdw + 31) / 32), (unsigned)((cs.q + 31) / 32));
      k_xpose_multi<<<grid, 256, 0, S>>>(me.hvt + cs.off * c.dimdw, c.dimdw, c.dimdw, cs.q, xa);
      c.launches++;
    }
    prof_end();
    CB_CUDA(cudaEventRecord(c.ev_pack, S));
    for (int k = 1; k < P; k++) {
      const int p = (me.rank + k) % P;
      const Split pd = split_of(c.dimdw, P, p);
      if (pd.q <= 0) continue;
      // into p's window: sender block at q_dw(p) * up_off(me), inside it this chunk at q_dw(p) * cs.off
      const int64_t tot = pd.q * cs.q;
      for (int j = 0; j < nsplit; j++) {
        const Split r = split_of(tot, nsplit, j);
        cudaStream_t st = c.peer_stream[p * nsplit + j];
        CB_CUDA(cudaStreamWaitEvent(st, c.ev_pack, 0));
        if (r.q > 0)
          CB_CUDA(cudaMemcpyAsync(c.peer_recv[p] + pd.q * (me_up.off + cs.off) + r.off, me.sendbuf + ob[p] + pd.q * cs.off + r.off,
                                  (size_t)r.q * 16, cudaMemcpyDeviceToDevice, st));
        if (ch == nch - 1) CB_CUDA(cudaEventRecord(c.peer_done[p * nsplit + j], st));
      }
    }
  }
  prof_begin(6);  // what is left of the way back
  for (int k = 1; k < P; k++)
    for (int j = 0; j < nsplit; j++)
      if (split_of(c.dimdw, P, (me.rank + k) % P).q > 0) CB_CUDA(cudaStreamWaitEvent(S, c.peer_done[((me.rank + k) % P) * nsplit + j], 0));
  prof_end();
  prof_begin(5);  // every receive window is complete
  CB_CHECK(nccl_barrier());
  prof_end();
  prof_begin(2);
  if (me.dw.q > 0 && DU > 0) {  // one launch adds every sender's blocks to Hv
    UnpackArgs ua{};
    ua.nseg = P; ua.me = me.rank; ua.nch_opt = (int)std::max<int64_t>(1, xchg_chunks_eff(DU));
    for (int p = 0; p < P; p++) { ua.ub[p] = usplit(p).off; ua.ub[p + 1] = usplit(p).off + usplit(p).q; }
    if (me.dw.q > 65535) return fail("copy-engine exchange: more than 65535 local columns");
    dim3 grid((unsigned)((DU + 255) / 256), (unsigned)me.dw.q);
    k_unpack_multi<<<grid, 256, 0, S>>>(me.recvbuf, hv, DU, me.dw.q, ua);
    c.launches++;
  }
  prof_end();
  return 0;
}

// PAIRS (real Krylov mode on a sharded layout): v / hv hold REAL vectors.  The diag + Hup pass runs the real
// kernels; for everything after it two adjacent up-rows of the real vector are one double2 of a "complex"
// vector with DimUp/2 rows -- transposes, exchanges and the Hdw pass act on the other index, and a real
// coefficient treats both halves alike, so the complex code runs unchanged on half the bytes (HBM and NVLink).
static int hxv_local_terms(const double2 *v, double2 *hv, bool pairs) {
  Ctx &c = ctx();
  const bool sharded = c.spmd || c.sim || c.opt.force_sharded;
  DiagArgs nodiag{};
  if (pairs && (!sharded || (c.dimup & 1) || !c.real_h)) return fail("hxv: paired-row view needs a sharded layout, a real H and an even DimUp");
  const int64_t DU = pairs ? c.dimup / 2 : c.dimup;  // rows of the (possibly paired) view
  auto usplit = [&](int p) { return split_of(DU, c.p_eff, p); };
  if (!sharded) {
    // one rank.  Preferred order: the tile-resident row pass WRITES hv = v.Hdw^T (no read-modify-write of hv in
    // the pass that is bound by the L2 -> SM fill rate), then the column-resident pass adds diag + Hup.v, reading
    // the old hv with one coalesced load per output while its gathers run in shared memory.
    if (rowtile_applicable(c.dw) && colres_applicable<double2>(c.up, diag_args(0))) {
      CB_CHECK(rowpass(c.dw, c.dimup, v, hv, false));
      CB_CHECK(colpass(c.up, c.dimdw, v, hv, diag_args(0), true, !c.jhflag));
      return 0;
    }
    // otherwise: diag + up (column pass), then dw on the strided index (accumulating row pass)
    CB_CHECK(colpass(c.up, c.dimdw, v, hv, diag_args(0)));
    c.dot_final_rowpass = !c.jhflag;
    const int rc = rowpass(c.dw, c.dimup, v, hv, true);
    c.dot_final_rowpass = false;
    return rc;
  }
  // sharded: diag -> UP -> transpose -> DW on vt -> transpose back -> add  (spMatVec_mpi_main order)
  int64_t off = 0;
  std::vector<int64_t> offs;
  const int P = c.p_eff;
  // measured on B200 (K3): peer-memory transposes win at P=2 (8.4 vs 8.7 ms), the overlapped NCCL path wins
  // at P=4 (4.6 vs 4.9) and P=8 (2.5 vs 2.8) -> use_ipc 1 = auto (P<=2), 2 = always, 0 = never
  // (the stream-ordered barriers of that path span the full communicator: only when no rank was shrunk away)
  const bool ipc_ok = c.spmd && P > 1 && P == c.nranks && !c.rk.empty() && c.ipc_ready;
  if (ipc_ok && c.opt.use_ipc == 1) return hxv_sharded_ce(v, hv, pairs, DU);
  const bool use_ipc = ipc_ok && c.opt.use_ipc == 2;
  const bool overlap = !use_ipc && c.spmd && P > 1 && !c.rk.empty() && c.opt.overlap && c.comm_stream;
  if (use_ipc) {
    // peer-memory transpose: every rank stores its transposed blocks straight into the owners' vt
    // over NVLink (pack + exchange + unpack + local_transpose of vector_transpose_MPI in one kernel);
    // barrier A: every peer has finished reading its vt from the previous product
    // The forward transpose only reads v: it runs on the communication stream, overlapped with the
    // diag+Hup pass on the compute stream.
    RankState &me = c.rk[0];
    const Split me_up = usplit(me.rank);
    (void)me_up;
    cudaStream_t main = c.stream;
    // (measured on 2 B200: overlapping it with the column pass slows both -- they share the LSU
    // pipe -- so it is sequential unless overlap == 2)
    const bool ov = c.opt.overlap == 2 && c.comm_stream;
    if (ov) {
      CB_CUDA(cudaEventRecord(c.ev_in, main));
      CB_CUDA(cudaStreamWaitEvent(c.comm_stream, c.ev_in, 0));
      c.stream = c.comm_stream;
    }
    int rc = nccl_barrier();
    prof_begin(2);
    for (int k = 0; k < P && rc == 0; k++) {
      const int p = (me.rank + k) % P;  // staggered schedule: at step k every rank targets a different GPU
      Split pu = usplit(p);
      transpose_block<false>(v, DU, pu.off, pu.q, me.dw.q, c.peer_vt[p], c.dimdw, me.dw.off);
    }
    prof_end();
    if (rc == 0) rc = nccl_barrier();  // B: all blocks of every vt have landed
    if (ov) {
      cudaEventRecord(c.ev_comm, c.comm_stream);
      c.stream = main;
    }
    if (rc) return rc;
  }
  std::vector<int64_t> cs(c.nranks, 0), os(c.nranks, 0), cr(c.nranks, 0), orr(c.nranks, 0);
  if (overlap) {
    // the transpose of v does not depend on the diag+Hup pass: run pack -> all-to-all -> unpack on the
    // communication stream while the column pass runs on the compute stream
    RankState &me = c.rk[0];
    const Split me_up = usplit(me.rank);
    (void)me_up;
    cudaStream_t main = c.stream;
    int64_t so = 0, ro = 0;
    prof_begin(2);  // pack on the compute stream (alone, at full speed) ...
    for (int p = 0; p < P; p++) {
      Split pu = usplit(p), pd = split_of(c.dimdw, P, p);
      cs[p] = pu.q * me.dw.q; os[p] = so; so += cs[p];
      cr[p] = me_up.q * pd.q; orr[p] = ro; ro += cr[p];
      transpose_block<false>(v, DU, pu.off, pu.q, me.dw.q, me.sendbuf + os[p], me.dw.q, 0);
    }
    prof_end();
    CB_CUDA(cudaEventRecord(c.ev_in, main));
    CB_CUDA(cudaStreamWaitEvent(c.comm_stream, c.ev_in, 0));
    c.stream = c.comm_stream;  // ... all-to-all + unpack on the (high-priority) communication stream
    prof_begin(3);
    int rc = nccl_all_to_all(me.sendbuf, me.recvbuf, cs.data(), os.data(), cr.data(), orr.data());
    prof_end();
    if (rc == 0) {
      prof_begin(2);
      for (int p = 0; p < P; p++) {
        Split pd = split_of(c.dimdw, P, p);
        copy_block<false>(me.recvbuf + orr[p], me_up.q, pd.q, me.vt, c.dimdw, pd.off);
      }
      prof_end();
      cudaEventRecord(c.ev_comm, c.comm_stream);
    }
    c.stream = main;
    if (rc) return rc;
  }
  for (auto &r : c.rk) {
    offs.push_back(off);
    if (pairs) CB_CHECK(colpass_real(c.up, r.dw.q, (const double *)(v + off), (double *)(hv + off), diag_args(r.dw.off), false, false));
    else CB_CHECK(colpass(c.up, r.dw.q, v + off, hv + off, diag_args(r.dw.off)));
    off += pairs ? r.nloc / 2 : r.nloc;
  }
  if (!c.spmd || P == 1) {
    // device-local exchange: write straight into the destination rank's buffer
    prof_begin(2);
    for (size_t a = 0; a < c.rk.size(); a++)
      for (size_t b = 0; b < c.rk.size(); b++) {
        RankState &src = c.rk[a], &dst = c.rk[b];
        transpose_block<false>(v + offs[a], DU, usplit(dst.rank).off, usplit(dst.rank).q, src.dw.q, dst.vt, c.dimdw, src.dw.off);
      }
    prof_end();
    for (auto &r : c.rk) CB_CHECK(colpass(c.dw, usplit(r.rank).q, r.vt, r.hvt, nodiag));
    prof_begin(2);
    for (size_t a = 0; a < c.rk.size(); a++)
      for (size_t b = 0; b < c.rk.size(); b++) {
        RankState &src = c.rk[a], &dst = c.rk[b];
        transpose_block<true>(src.hvt, c.dimdw, dst.dw.off, dst.dw.q, usplit(src.rank).q, hv + offs[b], DU, usplit(src.rank).off);
      }
    prof_end();
    return 0;
  }
  // SPMD over NCCL: pack (transposing) -> grouped send/recv -> unpack
  if (c.rk.empty()) return 0;  // rank outside the shrunk communicator
  RankState &me = c.rk[0];
  const Split me_up = usplit(me.rank);
  if (use_ipc) {
    if (c.opt.overlap == 2 && c.comm_stream) CB_CUDA(cudaStreamWaitEvent(c.stream, c.ev_comm, 0));
    CB_CHECK(colpass(c.dw, me_up.q, me.vt, me.hvt, nodiag));
    prof_begin(2);
    for (int k = 0; k < P; k++) {  // back: my rows (up) x p's columns (dw) -> p's receive window, transposed
      const int p = (me.rank + k) % P;
      Split pd = split_of(c.dimdw, P, p);
      transpose_block<false>(me.hvt, c.dimdw, pd.off, pd.q, me_up.q, c.peer_recv[p] + pd.q * me_up.off, me_up.q, 0);
    }
    prof_end();
    CB_CHECK(nccl_barrier());  // C: my receive window is complete
    prof_begin(2);
    for (int p = 0; p < P; p++) {
      Split pu = usplit(p);
      copy_block<true>(me.recvbuf + me.dw.q * pu.off, me.dw.q, pu.q, hv, DU, pu.off);
    }
    prof_end();
    return 0;
  }
  int64_t so = 0, ro = 0;
  if (overlap) {
    CB_CUDA(cudaStreamWaitEvent(c.stream, c.ev_comm, 0));
  } else {
    prof_begin(2);
    for (int p = 0; p < P; p++) {
      Split pu = usplit(p), pd = split_of(c.dimdw, P, p);
      cs[p] = pu.q * me.dw.q; os[p] = so; so += cs[p];   // my columns, p's rows
      cr[p] = me_up.q * pd.q; orr[p] = ro; ro += cr[p];  // my rows, p's columns
      transpose_block<false>(v, DU, pu.off, pu.q, me.dw.q, me.sendbuf + os[p], me.dw.q, 0);
    }
    prof_end();
    prof_begin(3);
    CB_CHECK(nccl_all_to_all(me.sendbuf, me.recvbuf, cs.data(), os.data(), cr.data(), orr.data()));
    prof_end();
    prof_begin(2);
    for (int p = 0; p < P; p++) {
      Split pd = split_of(c.dimdw, P, p);
      copy_block<false>(me.recvbuf + orr[p], me_up.q, pd.q, me.vt, c.dimdw, pd.off);
    }
    prof_end();
  }
  CB_CHECK(colpass(c.dw, me_up.q, me.vt, me.hvt, nodiag));
  so = ro = 0;
  prof_begin(2);
  for (int p = 0; p < P; p++) {
    Split pu = usplit(p), pd = split_of(c.dimdw, P, p);
    cs[p] = pd.q * me_up.q; os[p] = so; so += cs[p];   // my rows (up), p's columns (dw)
    cr[p] = me.dw.q * pu.q; orr[p] = ro; ro += cr[p];
    transpose_block<false>(me.hvt, c.dimdw, pd.off, pd.q, me_up.q, me.sendbuf + os[p], me_up.q, 0);
  }
  prof_end();
  prof_begin(3);
  CB_CHECK(nccl_all_to_all(me.sendbuf, me.recvbuf, cs.data(), os.data(), cr.data(), orr.data()));
  prof_end();
  prof_begin(2);
  for (int p = 0; p < P; p++) {
    Split pu = usplit(p);
    copy_block<true>(me.recvbuf + orr[p], me.dw.q, pu.q, hv, DU, pu.off);
  }
  prof_end();
  return 0;
}

// real vectors on a sharded layout (called from hxv_real.cu): see PAIRS above
int hxv_sharded_real(const double *v, double *hv) { return hxv_local_terms((const double2 *)v, (double2 *)hv, true); }

// spH0d of the local rows
__global__ void k_diag_only(int64_t n, int64_t ncols, double *__restrict__ out, DiagArgs dg) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t c = blockIdx.y;
  if (i >= n || c >= ncols) return;
  uint32_t mu_imp = (uint32_t)__ldg(dg.map_row + i) & ((1u << dg.nimp) - 1u);
  out[i + c * n] = diag_value(dg, i, mu_imp, c);
}

// ------------------------------------------------------------------------------------
// Dense sector Hamiltonian, the optional `Hmat` of build_Hv_sector (ED_HAMILTONIAN.f90:123-127 ->
// ED_HAMILTONIAN_SPARSE_HxV.f90:112-148: Hmat = spH0d [+ spH0nd] + kron(spH0dws, 1) + kron(1, spH0ups); caller
// ED_DIAG.f90:199, the LAPACK branch for sectors below lanc_dim_threshold).  One thread per row i = (iup, idw)
// of the GLOBAL sector: every rank assembles the whole matrix, as every rank of the reference does.
// H is column-major [Dim, Dim], zeroed by the caller.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_dense_hmat(int64_t dim, int64_t dimup, DiagArgs dg, OpArgs up, OpArgs dw, NonLocalArgs nl,
                                                    int jhflag, double2 *__restrict__ H) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= dim) return;
  const int64_t iup = i % dimup, idw = i / dimup;
  const uint32_t mup = (uint32_t)__ldg(up.map + iup), mdw = (uint32_t)__ldg(dw.map + idw);
  auto add = [&](int64_t j, double re, double im) {
    double2 *p = H + i + j * dim;
    p->x += re;
    p->y += im;
  };
  add(i, diag_value(dg, iup, mup & ((1u << dg.nimp) - 1u), idw), 0.0);
  for (int t = 0; t < up.nterms; t++) {  // <i|H|j> = h(a,b) sg for |j> = c^+_b c_a |i>, as in the matrix-free column pass
    const Term tm = up.terms[t];
    if (((mup >> tm.a) & 1u) && !((mup >> tm.b) & 1u)) {
      const uint32_t m = (mup & ~(1u << tm.a)) | (1u << tm.b);
      const double sg = hop_sign_d(mup, tm.a, tm.b);
      add(lin_rank_d(up.lin_lo, up.lin_hi, up.lbits, m) + idw * dimup, tm.re * sg, tm.im * sg);
    }
  }
  for (int t = 0; t < dw.nterms; t++) {
    const Term tm = dw.terms[t];
    if (((mdw >> tm.a) & 1u) && !((mdw >> tm.b) & 1u)) {
      const uint32_t m = (mdw & ~(1u << tm.a)) | (1u << tm.b);
      const double sg = hop_sign_d(mdw, tm.a, tm.b);
      add(iup + (int64_t)lin_rank_d(dw.lin_lo, dw.lin_hi, dw.lbits, m) * dimup, tm.re * sg, tm.im * sg);
    }
  }
  if (!jhflag) return;
  for (int ilat = 0; ilat < nl.nlat; ilat++)  // same conditions and signs as k_nonlocal
    for (int io = 0; io < nl.norb; io++)
      for (int jo = 0; jo < nl.norb; jo++) {
        if (io == jo) continue;
        const int is = io + ilat * nl.norb, js = jo + ilat * nl.norb;
        const uint32_t bi = 1u << is, bj = 1u << js;
        const bool nup_i = mup & bi, nup_j = mup & bj, ndw_i = mdw & bi, ndw_j = mdw & bj;
        if (nl.jx != 0.0 && nup_j && ndw_i && !ndw_j && !nup_i) {
          const uint32_t kdw = (mdw & ~bi) | bj, kup = (mup & ~bj) | bi;
          const double sg = hop_sign2(mdw, js, is) * hop_sign2(mup, is, js);
          add(lin_rank_d(nl.up_lo, nl.up_hi, nl.lbits, kup) + (int64_t)lin_rank_d(nl.dw_lo, nl.dw_hi, nl.lbits, kdw) * dimup, nl.jx * sg, 0.0);
        }
        if (nl.jp != 0.0 && nup_j && ndw_j && !ndw_i && !nup_i) {
          const uint32_t kdw = (mdw & ~bj) | bi, kup = (mup & ~bj) | bi;
          const double sg = hop_sign2(mdw, is, js) * hop_sign2(mup, is, js);
          add(lin_rank_d(nl.up_lo, nl.up_hi, nl.lbits, kup) + (int64_t)lin_rank_d(nl.dw_lo, nl.dw_hi, nl.lbits, kdw) * dimup, nl.jp * sg, 0.0);
        }
      }
}

}  // namespace cb

using namespace cb;

extern "C" {

int cdmft_b200_hxv64(int64_t nloc, const void *v, void *hv) {
  CB_REQUIRE_INIT();
  Ctx &c = ctx();
  if (!c.hstatus) return fail("hxv: Hsector NOT set (call build_hv_sector)");  // directMatVec_cc ERROR
  int64_t expect = 0;
  for (auto &r : c.rk) expect += r.nloc;
  if (nloc != expect) return fail("hxv: Nloc=%lld != local dimension %lld of sector %d", (long long)nloc, (long long)expect, c.hsector);
  if (v == hv) return fail("hxv: v and hv must not alias");
  if (nloc == 0) return 0;
  const bool dv = is_device_ptr(v), dh = is_device_ptr(hv);
  if (dv != dh) return fail("hxv: v and hv must both be host or both be device pointers");
  if (dv) {
    CB_CHECK(hxv_device((const double2 *)v, (double2 *)hv));
    CB_CUDA(cudaGetLastError());
    return 0;
  }
  CB_CHECK(ensure_stage(nloc));
  CB_CUDA(cudaMemcpyAsync(c.stage_v, v, (size_t)nloc * 16, cudaMemcpyHostToDevice, c.stream));
  CB_CHECK(hxv_device(c.stage_v, c.stage_hv));
  CB_CUDA(cudaGetLastError());
  CB_CUDA(cudaMemcpyAsync(hv, c.stage_hv, (size_t)nloc * 16, cudaMemcpyDeviceToHost, c.stream));
  CB_CUDA(cudaStreamSynchronize(c.stream));
  return 0;
}

int cdmft_b200_hxv(int32_t nloc, const void *v, void *hv) { return cdmft_b200_hxv64((int64_t)nloc, v, hv); }

int cdmft_b200_get_diag(int64_t nloc, double *d) {
  CB_REQUIRE_INIT();
  Ctx &c = ctx();
  if (!c.hstatus) return fail("get_diag: no active sector");
  int64_t expect = 0;
  for (auto &r : c.rk) expect += r.nloc;
  if (nloc != expect) return fail("get_diag: nloc mismatch");
  double *dd = nullptr;
  CB_CHECK(dev_alloc(&dd, nloc));
  int64_t off = 0;
  for (auto &r : c.rk) {
    if (r.dw.q > 0) {
      dim3 grid((unsigned)((c.dimup + 255) / 256), (unsigned)r.dw.q);
      k_diag_only<<<grid, 256, 0, c.stream>>>(c.dimup, r.dw.q, dd + off, diag_args(r.dw.off));
      c.launches++;
    }
    off += r.nloc;
  }
  CB_CUDA(cudaMemcpyAsync(d, dd, nloc * 8, cudaMemcpyDeviceToHost, c.stream));
  CB_CUDA(cudaStreamSynchronize(c.stream));
  cudaFree(dd);
  return 0;
}

int cdmft_b200_build_hmat(void *hmat) {
  CB_REQUIRE_INIT();
  Ctx &c = ctx();
  if (!c.hstatus) return fail("build_hmat: Hsector NOT set (call build_hv_sector first)");
  if (c.dim > 32768) return fail("build_hmat: Dim=%lld is too large for a dense matrix (the reference uses it below lanc_dim_threshold)", (long long)c.dim);
  const int64_t n2 = c.dim * c.dim;
  const bool dev = is_device_ptr(hmat);
  double2 *d = (double2 *)hmat;
  if (!dev) CB_CHECK(dev_alloc(&d, n2));
  int rc = 0;
  auto body = [&]() -> int {
    CB_CUDA(cudaMemsetAsync(d, 0, (size_t)n2 * sizeof(double2), c.stream));
    NonLocalArgs a{};
    a.map_up = c.up.map; a.map_dw = c.dw.map;
    a.up_lo = c.up.lin_lo; a.up_hi = c.up.lin_hi; a.dw_lo = c.dw.lin_lo; a.dw_hi = c.dw.lin_hi;
    a.lbits = c.ns / 2; a.nlat = c.m.nlat; a.norb = c.m.norb; a.jx = c.m.jx; a.jp = c.m.jp; a.dimup = c.dimup;
    if (c.dim > 0) {
      k_dense_hmat<<<(unsigned)((c.dim + 127) / 128), 128, 0, c.stream>>>(c.dim, c.dimup, diag_args(0), op_args(c.up), op_args(c.dw), a,
                                                                          c.jhflag ? 1 : 0, d);
      c.launches++;
    }
    if (!dev) CB_CUDA(cudaMemcpyAsync(hmat, d, (size_t)n2 * sizeof(double2), cudaMemcpyDeviceToHost, c.stream));
    CB_CUDA(cudaStreamSynchronize(c.stream));
    CB_CUDA(cudaGetLastError());
    return 0;
  };
  rc = body();
  if (!dev && d) cudaFree(d);
  return rc;
}

// scatter_vector_MPI / gather_vector_MPI (ED_SETUP.f90:575-668): the root's full sector vector <-> every rank's
// Ndw shard (shards are contiguous: rank r owns columns [off_r, off_r + q_r) of v(DimUp, DimDw)).  The reference
// needs them around sp_lanc_tridiag (ED_GF_NORMAL.f90:214) and es_return_cvector (ED_EIGENSPACE.f90:499-569).
// vfull is significant on the root only; host or device pointers.  Single rank / simulated ranks: plain copies.
static int shard_bounds_of(int64_t dimup, int64_t dimdw, int rank, int64_t *off, int64_t *cnt) {
  Ctx &c = ctx();
  const int peff = (int)std::min<int64_t>(c.nranks, dimdw);  // ED_HAMILTONIAN.f90:62-90
  *off = *cnt = 0;
  if (rank >= peff) return 0;
  const Split s = split_of(dimdw, peff, rank);
  *off = s.off * dimup;
  *cnt = s.q * dimup;
  return 0;
}
}  // extern "C"
namespace cb {
// scatter (root's full vector -> shards) / gather of a vector of ANY sector (dimup x dimdw), SPMD or not
int scatter_gather_dims(void *vfull, void *vloc, int root, bool scatter, int64_t dimup, int64_t dimdw) {
  Ctx &c = ctx();
  const int64_t dim = dimup * dimdw;
  auto shard_bounds = [&](int rank, int64_t *off, int64_t *cnt) { return shard_bounds_of(dimup, dimdw, rank, off, cnt); };
  const bool spmd = c.spmd && c.nranks > 1;
  if (root < 0 || root >= (spmd ? c.nranks : 1)) return fail("scatter/gather_vector: bad root %d", root);
  if (!spmd) {  // one process holds every shard back to back = the full vector
    if (vfull == vloc || dim == 0) return 0;
    CB_CUDA(cudaMemcpyAsync(scatter ? vloc : vfull, scatter ? vfull : vloc, (size_t)dim * 16, cudaMemcpyDefault, c.stream));
    CB_CUDA(cudaStreamSynchronize(c.stream));
    return 0;
  }
  int64_t moff = 0, mcnt = 0;
  shard_bounds(c.rank, &moff, &mcnt);
  const bool me_root = c.rank == root;
  // device staging for host pointers
  double2 *dfull = nullptr, *dloc = nullptr;
  const bool full_dev = me_root && vfull && is_device_ptr(vfull), loc_dev = mcnt > 0 && is_device_ptr(vloc);
  int rc = 0;
  auto body = [&]() -> int {
    if (me_root) {
      if (full_dev) dfull = (double2 *)vfull;
      else {
        CB_CHECK(dev_alloc(&dfull, dim));
        if (scatter) CB_CUDA(cudaMemcpyAsync(dfull, vfull, (size_t)dim * 16, cudaMemcpyHostToDevice, c.stream));
      }
    }
    if (mcnt > 0) {
      if (loc_dev) dloc = (double2 *)vloc;
      else {
        CB_CHECK(dev_alloc(&dloc, mcnt));
        if (!scatter) CB_CUDA(cudaMemcpyAsync(dloc, vloc, (size_t)mcnt * 16, cudaMemcpyHostToDevice, c.stream));
      }
    }
    std::vector<int64_t> cs(c.nranks, 0), os(c.nranks, 0), cr(c.nranks, 0), orr(c.nranks, 0);
    if (scatter) {
      if (me_root) for (int p = 0; p < c.nranks; p++) shard_bounds(p, &os[p], &cs[p]);
      cr[root] = mcnt;
      CB_CHECK(nccl_all_to_all(dfull, dloc, cs.data(), os.data(), cr.data(), orr.data()));
    } else {
      cs[root] = mcnt;
      if (me_root) for (int p = 0; p < c.nranks; p++) shard_bounds(p, &orr[p], &cr[p]);
      CB_CHECK(nccl_all_to_all(dloc, dfull, cs.data(), os.data(), cr.data(), orr.data()));
    }
    if (scatter && mcnt > 0 && !loc_dev) CB_CUDA(cudaMemcpyAsync(vloc, dloc, (size_t)mcnt * 16, cudaMemcpyDeviceToHost, c.stream));
    if (!scatter && me_root && !full_dev) CB_CUDA(cudaMemcpyAsync(vfull, dfull, (size_t)dim * 16, cudaMemcpyDeviceToHost, c.stream));
    CB_CUDA(cudaStreamSynchronize(c.stream));
    return 0;
  };
  rc = body();
  if (me_root && !full_dev && dfull) cudaFree(dfull);
  if (mcnt > 0 && !loc_dev && dloc) cudaFree(dloc);
  return rc;
}
}  // namespace cb
extern "C" {
int cdmft_b200_scatter_vector(const void *vfull, void *vloc, int32_t root) {
  CB_REQUIRE_INIT();
  Ctx &c = ctx();
  if (!c.hstatus) return fail("scatter_vector: Hsector NOT set");
  return scatter_gather_dims((void *)vfull, vloc, root, true, c.dimup, c.dimdw);
}
int cdmft_b200_gather_vector(const void *vloc, void *vfull, int32_t root) {
  CB_REQUIRE_INIT();
  Ctx &c = ctx();
  if (!c.hstatus) return fail("gather_vector: Hsector NOT set");
  return scatter_gather_dims(vfull, (void *)vloc, root, false, c.dimup, c.dimdw);
}

}  // extern "C"
