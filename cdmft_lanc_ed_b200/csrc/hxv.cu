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

namespace cb {

__device__ __forceinline__ double2 ldg2(const double2 *p) { return __ldg(p); }
__device__ __forceinline__ void cfma(double2 &acc, double2 h, double2 x) {  // acc += h*x (complex)
  acc.x = fma(h.x, x.x, acc.x);
  acc.x = fma(-h.y, x.y, acc.x);
  acc.y = fma(h.x, x.y, acc.y);
  acc.y = fma(h.y, x.x, acc.y);
}
__device__ __forceinline__ void rfma(double2 &acc, double h, double2 x) {  // real coefficient
  acc.x = fma(h, x.x, acc.x);
  acc.y = fma(h, x.y, acc.y);
}
__device__ __forceinline__ int32_t lin_rank_d(const int32_t *__restrict__ lo, const int32_t *__restrict__ hi, int lbits,
                                              uint32_t s) {
  return __ldg(hi + (s >> lbits)) + __ldg(lo + (s & ((1u << lbits) - 1u)));
}
__device__ __forceinline__ double hop_sign_d(uint32_t s, int a, int b) {
  int lo = min(a, b), hi = max(a, b);
  uint32_t between = ((1u << hi) - 1u) & ~((2u << lo) - 1u);
  return (__popc(s & between) & 1) ? -1.0 : 1.0;
}

struct DiagArgs {  // diagonal d(i,c) = f_row[i] + f_col[coloff+c] + sum_{b in md} T[b][mu]
  int enabled;
  const double *f_row;      // indexed by the contiguous index i
  const double *f_col;      // indexed by the global column
  const int32_t *map_row;   // Fock state of row i
  const int32_t *map_col;   // Fock state of global column
  const double *cross_tab;  // [Nimp][2^Nimp]
  int nimp;
  int64_t coloff;
};

struct OpArgs {  // gather operator on the contiguous index
  const int32_t *ell_col;
  const double2 *ell_val;
  const int32_t *rowlen;
  int ell_w;
  // matrix-free
  const int32_t *map;
  const int32_t *lin_lo, *lin_hi;
  int lbits;
  const Term *terms;
  int nterms;
};

__device__ __forceinline__ double diag_value(const DiagArgs &d, int64_t i, uint32_t mu_imp, int64_t c) {
  double val = __ldg(d.f_row + i) + __ldg(d.f_col + d.coloff + c);
  uint32_t md = (uint32_t)__ldg(d.map_col + d.coloff + c) & ((1u << d.nimp) - 1u);
  const int64_t nst = (int64_t)1 << d.nimp;
  while (md) {
    int b = __ffs(md) - 1;
    md &= md - 1;
    val += __ldg(d.cross_tab + (int64_t)b * nst + mu_imp);
  }
  return val;
}

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

// ------------------------------------------------------------------------------------
static OpArgs op_args(const SpinOp &s) {
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
  const bool direct = c.mode == CDMFT_B200_DIRECT;
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

static int colpass(const SpinOp &s, int64_t ncols, const double2 *v, double2 *out, const DiagArgs &dg) {
  switch (ctx().opt.col_batch) {
    case 1: return launch_colpass<1>(s, ncols, v, out, dg);
    case 2: return launch_colpass<2>(s, ncols, v, out, dg);
    case 8: return launch_colpass<8>(s, ncols, v, out, dg);
    default: return launch_colpass<4>(s, ncols, v, out, dg);
  }
}

static int rowpass(const SpinOp &s, int64_t nrows, const double2 *v, double2 *out) {
  Ctx &c = ctx();
  if (s.n <= 0 || nrows <= 0) return 0;
  dim3 grid((unsigned)s.n, (unsigned)((nrows + 255) / 256));
  if (grid.y > 65535) return fail("rowpass: too many row chunks");
  OpArgs op = op_args(s);
  const bool direct = c.mode == CDMFT_B200_DIRECT;
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

static DiagArgs diag_args(int64_t coloff) {
  Ctx &c = ctx();
  DiagArgs d{};
  d.enabled = 1;
  d.f_row = c.up.f; d.f_col = c.dw.f;
  d.map_row = c.up.map; d.map_col = c.dw.map;
  d.cross_tab = c.cross_tab;
  d.nimp = c.nimp;
  d.coloff = coloff;
  return d;
}

// H x v on the local shard(s); v/hv device pointers laid out as the concatenation of the local
// ranks' shards (exactly one shard in SPMD mode, all P in sim mode, the whole vector otherwise).
int hxv_device(const double2 *v, double2 *hv) {
  Ctx &c = ctx();
  const bool sharded = c.spmd || c.sim || c.opt.force_sharded;
  DiagArgs nodiag{};
  if (!sharded) {
    // one rank: diag + up (column pass), then dw on the strided index (row pass)
    CB_CHECK(colpass(c.up, c.dimdw, v, hv, diag_args(0)));
    CB_CHECK(rowpass(c.dw, c.dimup, v, hv));
    return 0;
  }
  // sharded: diag -> UP -> transpose -> DW on vt -> transpose back -> add  (spMatVec_mpi_main order)
  int64_t off = 0;
  std::vector<int64_t> offs;
  for (auto &r : c.rk) {
    offs.push_back(off);
    CB_CHECK(colpass(c.up, r.dw.q, v + off, hv + off, diag_args(r.dw.off)));
    off += r.nloc;
  }
  const int P = c.p_eff;
  if (!c.spmd || P == 1) {
    // device-local exchange: write straight into the destination rank's buffer
    for (size_t a = 0; a < c.rk.size(); a++)
      for (size_t b = 0; b < c.rk.size(); b++) {
        RankState &src = c.rk[a], &dst = c.rk[b];
        transpose_block<false>(v + offs[a], c.dimup, dst.up.off, dst.up.q, src.dw.q, dst.vt, c.dimdw, src.dw.off);
      }
    for (auto &r : c.rk) CB_CHECK(colpass(c.dw, r.up.q, r.vt, r.hvt, nodiag));
    for (size_t a = 0; a < c.rk.size(); a++)
      for (size_t b = 0; b < c.rk.size(); b++) {
        RankState &src = c.rk[a], &dst = c.rk[b];
        transpose_block<true>(src.hvt, c.dimdw, dst.dw.off, dst.dw.q, src.up.q, hv + offs[b], c.dimup, src.up.off);
      }
    return 0;
  }
  // SPMD over NCCL: pack (transposing) -> grouped send/recv -> unpack
  if (c.rk.empty()) return 0;  // rank outside the shrunk communicator
  RankState &me = c.rk[0];
  std::vector<int64_t> cs(c.nranks, 0), os(c.nranks, 0), cr(c.nranks, 0), orr(c.nranks, 0);
  int64_t so = 0, ro = 0;
  for (int p = 0; p < P; p++) {
    Split pu = split_of(c.dimup, P, p), pd = split_of(c.dimdw, P, p);
    cs[p] = pu.q * me.dw.q; os[p] = so; so += cs[p];   // my columns, p's rows
    cr[p] = me.up.q * pd.q; orr[p] = ro; ro += cr[p];  // my rows, p's columns
    transpose_block<false>(v, c.dimup, pu.off, pu.q, me.dw.q, me.sendbuf + os[p], me.dw.q, 0);
  }
  CB_CHECK(nccl_all_to_all(me.sendbuf, me.recvbuf, cs.data(), os.data(), cr.data(), orr.data()));
  for (int p = 0; p < P; p++) {
    Split pd = split_of(c.dimdw, P, p);
    copy_block<false>(me.recvbuf + orr[p], me.up.q, pd.q, me.vt, c.dimdw, pd.off);
  }
  CB_CHECK(colpass(c.dw, me.up.q, me.vt, me.hvt, nodiag));
  so = ro = 0;
  for (int p = 0; p < P; p++) {
    Split pu = split_of(c.dimup, P, p), pd = split_of(c.dimdw, P, p);
    cs[p] = pd.q * me.up.q; os[p] = so; so += cs[p];   // my rows (up), p's columns (dw)
    cr[p] = me.dw.q * pu.q; orr[p] = ro; ro += cr[p];
    transpose_block<false>(me.hvt, c.dimdw, pd.off, pd.q, me.up.q, me.sendbuf + os[p], me.up.q, 0);
  }
  CB_CHECK(nccl_all_to_all(me.sendbuf, me.recvbuf, cs.data(), os.data(), cr.data(), orr.data()));
  for (int p = 0; p < P; p++) {
    Split pu = split_of(c.dimup, P, p);
    copy_block<true>(me.recvbuf + orr[p], me.dw.q, pu.q, hv, c.dimup, pu.off);
  }
  return 0;
}

// spH0d of the local rows
__global__ void k_diag_only(int64_t n, int64_t ncols, double *__restrict__ out, DiagArgs dg) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t c = blockIdx.y;
  if (i >= n || c >= ncols) return;
  uint32_t mu_imp = (uint32_t)__ldg(dg.map_row + i) & ((1u << dg.nimp) - 1u);
  out[i + c * n] = diag_value(dg, i, mu_imp, c);
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

}  // extern "C"
