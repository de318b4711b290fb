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
// Column pass, shared-memory tile variant.  One CTA owns (row block) x (8 columns): the block is a
// run of Fock states sharing their top bits, so every hop that leaves those bits alone lands
// inside the tile.  The tile is staged once with cp.async (16 B, coalesced along the rows) into
// tile[row][8 slots], slot = column ^ (row & 7); in the compute phase 8 consecutive lanes own the
// 8 columns of ONE output row, so each gather reads one full 128-byte line of shared memory:
// conflict-free by construction.  The CSR row is fetched cooperatively (lane c loads entry
// base+c) and broadcast inside the 8-lane group with shuffles.  Hops that change the top bits
// (a minority) gather from global memory / L2.
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

template <bool REALH, bool DIRECT>
__global__ void __launch_bounds__(1024, 1) k_colpass_tile(int64_t n, int64_t ncols, const double2 *__restrict__ v,
                                                           double2 *__restrict__ out, const int2 *__restrict__ blocks,
                                                           int nblocks, const int32_t *__restrict__ rowptr,
                                                           const int32_t *__restrict__ col,
                                                           const double2 *__restrict__ val, OpArgs op, DiagArgs dg) {
  extern __shared__ double2 tile[];
  const int blk = blockIdx.x % nblocks;
  const int64_t c0 = (int64_t)(blockIdx.x / nblocks) * 8;
  const int2 b = blocks[blk];
  const int g0 = b.x, ng = b.y;
  const int nb = (int)min((int64_t)8, ncols - c0);
  // ---- stage the tile: thread -> (row g, column cc), rows fastest => 512 B coalesced per warp
  for (int idx = threadIdx.x; idx < ng * 8; idx += blockDim.x) {
    const int cc = idx / ng, g = idx - cc * ng;
    if (cc < nb) cp_async16(&tile[g * 8 + (cc ^ (g & 7))], v + (g0 + g) + (c0 + cc) * n);
  }
  cp_async_wait_all();
  __syncthreads();
  // ---- compute: 8 lanes = 8 columns of one row
  const int c = threadIdx.x & 7;
  const int rsub = threadIdx.x >> 3, rstep = blockDim.x >> 3;
  const int64_t colg = c0 + min(c, nb - 1);  // ragged last group: clamp loads, skip the store
  const unsigned gmask = 0xFFu << (threadIdx.x & 24);
  for (int g = rsub; g < ng; g += rstep) {
    const int64_t i = g0 + g;
    double2 acc = make_double2(0.0, 0.0);
    if (dg.enabled) {
      uint32_t mu_imp = (uint32_t)__ldg(dg.map_row + i) & ((1u << dg.nimp) - 1u);
      double d = diag_value(dg, i, mu_imp, colg);
      double2 x = tile[g * 8 + (c ^ (g & 7))];
      acc = make_double2(d * x.x, d * x.y);
    }
    if (!DIRECT) {
      const int32_t p0 = __ldg(rowptr + i), p1 = __ldg(rowptr + i + 1);
      for (int32_t base = p0; base < p1; base += 8) {
        const int32_t my = base + c;
        int32_t jc = 0;
        double2 hv = make_double2(0.0, 0.0);
        if (my < p1) { jc = __ldg(col + my); hv = ldg2(val + my); }
        const int cnt = min(8, p1 - base);
        for (int kk = 0; kk < cnt; kk++) {
          const int32_t j = __shfl_sync(gmask, jc, kk, 8);
          const double hx = __shfl_sync(gmask, hv.x, kk, 8);
          double hy = 0.0;
          if (!REALH) hy = __shfl_sync(gmask, hv.y, kk, 8);
          const unsigned rel = (unsigned)(j - g0);
          double2 x;
          if (rel < (unsigned)ng) x = tile[rel * 8 + (c ^ (rel & 7))];
          else x = ldg2(v + j + colg * n);
          if (REALH) rfma(acc, hx, x); else cfma(acc, make_double2(hx, hy), x);
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
          const unsigned rel = (unsigned)(j - g0);
          double2 x;
          if (rel < (unsigned)ng) x = tile[rel * 8 + (c ^ (rel & 7))];
          else x = ldg2(v + j + colg * n);
          if (REALH) rfma(acc, tm.re * sg, x); else cfma(acc, make_double2(tm.re * sg, tm.im * sg), x);
        }
      }
    }
    if (c < nb) out[i + colg * n] = acc;
  }
}

// ------------------------------------------------------------------------------------
// Packed shared-memory tile kernel (SPARSE mode, the fast path for both passes).
// Same tiling as k_colpass_tile / k_rowpass_tile, but the operator rows come from the packed tile
// CSR (see ctx.h): per row one list of sources inside the row block (shared memory) and one of
// sources outside it (global memory / L2), 32-bit words in rounds of 8.  Lane c of an 8-lane group
// loads word c of the round, the group broadcasts the words with SHFL; an in-block entry then costs
// one shift, one LOP3 (slot ^ lane, mask), one AND (coefficient offset), LDS.128 + LDS.64, 2 DFMA.
//   COLMODE = true : column pass. tile[row][slot], slot = column ^ (row&7); element (g, cc) of the
//                    tile lives at v[(g0+g) + (c0+cc)*ld]; off-block source j at v[j + (c0+cc)*ld]
//   COLMODE = false: row pass. tile[dwstate][iup 0..7]; element (g, bb) at v[(i0+bb) + (g0+g)*ld];
//                    off-block source j at v[(i0+bb) + j*ld]; out is accumulated (+=)
// Control flow is warp-uniform (the 4 row groups of a warp run the same number of rounds; short
// rows run null rounds with coefficient id 0) so the shuffles are plain full-mask SHFL.IDX.
// ------------------------------------------------------------------------------------
template <bool REALH, bool COLMODE>
__global__ void __launch_bounds__(1024, 1) k_tile_pk(int64_t ld, int64_t nbatch, const double2 *__restrict__ v,
                                                      double2 *__restrict__ out, const int2 *__restrict__ blocks,
                                                      int nblocks, const int32_t *__restrict__ in_ptr,
                                                      const uint32_t *__restrict__ pk_in,
                                                      const int32_t *__restrict__ off_ptr,
                                                      const uint32_t *__restrict__ pk_off,
                                                      const double2 *__restrict__ coef_g, DiagArgs dg) {
  extern __shared__ double2 smem[];
  double2 *coef = smem;        // [128]
  double2 *tile = smem + 128;  // [ng*8]
  const int blk = blockIdx.x % nblocks;
  const int64_t b0 = (int64_t)(blockIdx.x / nblocks) * 8;
  const int2 b = blocks[blk];
  const int g0 = b.x, ng = b.y;
  const int nb = (int)min((int64_t)8, nbatch - b0);
  if (threadIdx.x < 128) coef[threadIdx.x] = coef_g[threadIdx.x];
  if (COLMODE) {
    for (int idx = threadIdx.x; idx < ng * 8; idx += blockDim.x) {
      const int cc = idx / ng, g = idx - cc * ng;
      if (cc < nb) cp_async16(&tile[g * 8 + (cc ^ (g & 7))], v + (g0 + g) + (b0 + cc) * ld);
    }
  } else {
    for (int idx = threadIdx.x; idx < ng * 8; idx += blockDim.x) {
      const int bb = idx & 7, g = idx >> 3;
      if (bb < nb) cp_async16(&tile[idx], v + (b0 + bb) + (int64_t)(g0 + g) * ld);
    }
  }
  cp_async_wait_all();
  __syncthreads();
  const int c = threadIdx.x & 7;
  const unsigned c16 = (unsigned)c << 4;
  const int rsub = threadIdx.x >> 3, rstep = blockDim.x >> 3;
  const int64_t bg = b0 + min(c, nb - 1);  // my batch entry (clamped on the ragged last group)
  // off-block gathers: COLMODE v[j + bg*ld], row mode v[bg + j*ld]
  const double2 *vbase = COLMODE ? v + bg * ld : v + bg;
  const int64_t gstride = COLMODE ? 1 : ld;
  const char *tile_b = (const char *)tile;
  const char *coef_b = (const char *)coef;
  const int niter = (ng + rstep - 1) / rstep;
  for (int it = 0; it < niter; it++) {
    const int g_raw = rsub + it * rstep;
    const bool valid = g_raw < ng;
    const int g = valid ? g_raw : ng - 1;
    const int64_t i = g0 + g;
    double2 acc = make_double2(0.0, 0.0);
    if (COLMODE && dg.enabled) {
      uint32_t mu_imp = (uint32_t)__ldg(dg.map_row + i) & ((1u << dg.nimp) - 1u);
      double d = diag_value(dg, i, mu_imp, bg);
      double2 x = tile[(g << 3) | ((c ^ g) & 7)];
      acc = make_double2(d * x.x, d * x.y);
    }
    // ---- sources outside the row block (long latency) first
    {
      const int32_t o0 = __ldg(off_ptr + i);
      const int no = valid ? __ldg(off_ptr + i + 1) - o0 : 0;
      int nomax = max(no, __shfl_xor_sync(0xffffffffu, no, 8));
      nomax = max(nomax, __shfl_xor_sync(0xffffffffu, nomax, 16));
      for (int rr = 0; rr < nomax; rr++) {
        uint32_t mine = 0u;
        if (rr < no) mine = __ldg(pk_off + (int64_t)(o0 + rr) * 8 + c);
#pragma unroll
        for (int kk = 0; kk < 8; kk++) {
          const uint32_t w = __shfl_sync(0xffffffffu, mine, kk, 8);
          const uint32_t idb = w & 0x7F0u;
          if (idb) {
            const double2 x = ldg2(vbase + (int64_t)(w >> 11) * gstride);
            if (REALH) rfma(acc, *(const double *)(coef_b + idb), x);
            else cfma(acc, *(const double2 *)(coef_b + idb), x);
          }
        }
      }
    }
    // ---- sources inside the row block: shared memory
    {
      const int32_t r0 = __ldg(in_ptr + i);
      const int nr = valid ? __ldg(in_ptr + i + 1) - r0 : 0;
      int nrmax = max(nr, __shfl_xor_sync(0xffffffffu, nr, 8));
      nrmax = max(nrmax, __shfl_xor_sync(0xffffffffu, nrmax, 16));
      const uint32_t nullw = (((uint32_t)g << 3) | (COLMODE ? ((uint32_t)g & 7u) : 0u)) << 11;  // coef id 0 = 0.0
      uint32_t mine = nullw;
      if (0 < nr) mine = __ldg(pk_in + (int64_t)r0 * 8 + c);
      for (int rr = 0; rr < nrmax; rr++) {
        uint32_t next = nullw;  // prefetch the next round's word
        if (rr + 1 < nr) next = __ldg(pk_in + (int64_t)(r0 + rr + 1) * 8 + c);
#pragma unroll
        for (int kk = 0; kk < 8; kk++) {
          const uint32_t w = __shfl_sync(0xffffffffu, mine, kk, 8);
          const uint32_t addr = ((w >> 7) ^ c16) & 0xFFFFFFF0u;
          const double2 x = *(const double2 *)(tile_b + addr);
          if (REALH) rfma(acc, *(const double *)(coef_b + (w & 0x7F0u)), x);
          else cfma(acc, *(const double2 *)(coef_b + (w & 0x7F0u)), x);
        }
        mine = next;
      }
    }
    if (valid && c < nb) {
      double2 *o = COLMODE ? out + i + bg * ld : out + bg + i * ld;
      if (!COLMODE) { double2 y = *o; acc.x += y.x; acc.y += y.y; }
      *o = acc;
    }
  }
}

// ------------------------------------------------------------------------------------
// Column pass, "rotating slot" shared-memory kernel (SPARSE mode).
// One CTA owns (row block) x (8 columns), tile[row][8 slots] UNswizzled.  Each THREAD owns one output
// row and all 8 columns of it (8 accumulators), so an operator entry is decoded once per 8 outputs
// and no shuffles are needed.  Lane l reads, at step kk, slot (kk + l) & 7 of its source row: the 8
// lanes of a quarter-warp then hit 8 different 16-byte bank groups WHATEVER rows they gather from ->
// every LDS.128 is conflict-free (128 B per wavefront, the shared-memory peak).  acc[kk] therefore
// holds column (kk + l) & 7; the permutation is undone in the address of the final store.
// Rows are dealt to lanes as row = 4*(l&7) + (l>>3) inside each group of 32, so the four lanes that
// touch the same column at a step own 4 consecutive rows: stores and off-block loads move 64-byte
// runs (full sectors).  Hops that change the top bits gather from global memory / L2.
// ------------------------------------------------------------------------------------
template <bool REALH, bool INBLOCK_ONLY>
__global__ void __launch_bounds__(576, 1) k_colpass_rot(int64_t n, int64_t ncols, const double2 *__restrict__ v,
                                                         double2 *__restrict__ out, const int2 *__restrict__ blocks,
                                                         int nblocks, const uint32_t *__restrict__ pkell,
                                                         const int32_t *__restrict__ rowlen,
                                                         const int2 *__restrict__ rowsplit,
                                                         const double2 *__restrict__ coef_g, DiagArgs dg) {
  extern __shared__ double2 smem[];
  double2 *coef = smem;        // [128]
  double2 *tile = smem + 128;  // [ng*8]
  const int blk = blockIdx.x % nblocks;
  const int64_t c0 = (int64_t)(blockIdx.x / nblocks) * 8;
  const int2 b = blocks[blk];
  const int g0 = b.x, ng = b.y;
  const int nb = (int)min((int64_t)8, ncols - c0);
  if (threadIdx.x < 128) coef[threadIdx.x] = coef_g[threadIdx.x];
  // staging: 8 consecutive lanes = the 8 columns of one row -> conflict-free 128-byte smem lines
  for (int idx = threadIdx.x; idx < ng * 8; idx += blockDim.x) {
    const int cc = idx & 7, g = idx >> 3;
    if (cc < nb) cp_async16(&tile[idx], v + (g0 + g) + (c0 + cc) * n);
    else tile[idx] = make_double2(0.0, 0.0);
  }
  cp_async_wait_all();
  __syncthreads();
  const int lane = threadIdx.x & 31, L = lane & 7;
  const int rowperm = (L << 2) | (lane >> 3);  // row inside the group of 32
  const char *tile_b = (const char *)tile;
  const char *coef_b = (const char *)coef;
  unsigned slot16[8];    // byte offset of the slot this lane reads at step kk
  const double2 *vcol[8];  // global column base of that slot (clamped on the ragged last group)
#pragma unroll
  for (int kk = 0; kk < 8; kk++) {
    const int col = (kk + L) & 7;
    slot16[kk] = (unsigned)col << 4;
    vcol[kk] = v + (c0 + min(col, nb - 1)) * n;
  }
  for (int gbase = (threadIdx.x & ~31); gbase < ng; gbase += blockDim.x) {
    const int g = gbase + rowperm;
    if (g >= ng) continue;
    const int64_t i = g0 + g;
    double2 acc[8];
    if (dg.enabled) {
      const uint32_t mu_imp = (uint32_t)__ldg(dg.map_row + i) & ((1u << dg.nimp) - 1u);
#pragma unroll
      for (int kk = 0; kk < 8; kk++) {
        const int col = (kk + L) & 7;
        const double d = diag_value(dg, i, mu_imp, c0 + min(col, nb - 1));
        const double2 x = *(const double2 *)(tile_b + ((unsigned)g << 7) + slot16[kk]);
        acc[kk] = make_double2(d * x.x, d * x.y);
      }
    } else {
#pragma unroll
      for (int kk = 0; kk < 8; kk++) acc[kk] = make_double2(0.0, 0.0);
    }
    const int len = __ldg(rowlen + i);
    const int2 sp = __ldg(rowsplit + i);  // entries [sp.x, sp.y) gather inside the tile
    const uint32_t *pw = pkell + i;
    // three homogeneous loops (columns are ascending: below the block, inside, above): no branch
    // divergence between the lanes of a warp, only different trip counts
    auto off_block = [&](int k0, int k1) {
      for (int k = k0; k < k1; k++) {
        const uint32_t w = __ldg(pw + (int64_t)k * n);
        const uint32_t j = w >> 7;
        const double2 h = *(const double2 *)(coef_b + ((w & 127u) << 4));
        double2 x[8];
#pragma unroll
        for (int kk = 0; kk < 8; kk++) x[kk] = ldg2(vcol[kk] + j);
#pragma unroll
        for (int kk = 0; kk < 8; kk++) {
          if (REALH) rfma(acc[kk], h.x, x[kk]); else cfma(acc[kk], h, x[kk]);
        }
      }
    };
    if (!INBLOCK_ONLY) {  // otherwise the row-pass kernel adds the off-block terms (see k_rowpass UPOFF)
      off_block(0, sp.x);
      off_block(sp.y, len);
    }
    {
      uint32_t wnext = sp.x < sp.y ? __ldg(pw + (int64_t)sp.x * n) : 0u;
      for (int k = sp.x; k < sp.y; k++) {
        const uint32_t w = wnext;  // next operator word is prefetched
        if (k + 1 < sp.y) wnext = __ldg(pw + (int64_t)(k + 1) * n);
        const double2 h = *(const double2 *)(coef_b + ((w & 127u) << 4));
        const char *src = tile_b + (((w >> 7) - (unsigned)g0) << 7);
#pragma unroll
        for (int kk = 0; kk < 8; kk++) {
          const double2 x = *(const double2 *)(src + slot16[kk]);
          if (REALH) rfma(acc[kk], h.x, x); else cfma(acc[kk], h, x);
        }
      }
    }
#pragma unroll
    for (int kk = 0; kk < 8; kk++) {
      const int col = (kk + L) & 7;
      if (col < nb) out[i + (c0 + col) * n] = acc[kk];
    }
  }
}

// ------------------------------------------------------------------------------------
// Row pass (one rank, no transpose): out(i,c) += sum_k Hd(c,j_k) v(i,j_k).
// Threads run along i (contiguous), the operator row of column c is warp-uniform (broadcast
// loads).  Grid x = columns (fastest) inside one slab of rows, so a slab (rows x all columns)
// stays L2-resident while every column of it is produced.
// ------------------------------------------------------------------------------------
struct UpOffArgs {  // off-block Hup terms folded into the row pass (colpass_variant 5)
  const uint32_t *pkell;  // packed ELL of Hup: (col << 7) | coef_id, [k*n + i]
  const int32_t *rowlen;
  const int2 *rowsplit;   // entries [x,y) are in-block (done by k_colpass_rot<.,true>)
  const double2 *coef;    // [128]
};

template <bool REALH, bool DIRECT, bool UPOFF>
__global__ void __launch_bounds__(256) k_rowpass(int64_t n /*rows=DimUp*/, int64_t ncols /*DimDw*/,
                                                  const double2 *__restrict__ v, double2 *__restrict__ out,
                                                  const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                                                  const double2 *__restrict__ val, OpArgs op, UpOffArgs uo) {
  __shared__ double2 ucoef[UPOFF ? 128 : 1];
  if (UPOFF) {
    if (threadIdx.x < 128) ucoef[threadIdx.x] = uo.coef[threadIdx.x];
    __syncthreads();
  }
  const int64_t c = blockIdx.x;
  const int64_t i = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double2 acc = make_double2(0.0, 0.0);
  if (UPOFF) {
    // Hup entries whose source row lies outside row i's block: lanes = adjacent rows of ONE column, so
    // neighbouring sources share 128-byte lines (the cheap way to do the scattered gathers)
    const int len = __ldg(uo.rowlen + i);
    const int2 sp = __ldg(uo.rowsplit + i);
    const double2 *vc = v + c * n;
    const uint32_t *pw = uo.pkell + i;
    for (int k = 0; k < sp.x; k++) {
      const uint32_t w = __ldg(pw + (int64_t)k * n);
      const double2 x = ldg2(vc + (w >> 7)), h = ucoef[w & 127u];
      if (REALH) rfma(acc, h.x, x); else cfma(acc, h, x);
    }
    for (int k = sp.y; k < len; k++) {
      const uint32_t w = __ldg(pw + (int64_t)k * n);
      const double2 x = ldg2(vc + (w >> 7)), h = ucoef[w & 127u];
      if (REALH) rfma(acc, h.x, x); else cfma(acc, h, x);
    }
  }
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
// Row pass, shared-memory tile variant: one CTA owns (block of dw states) x (8 consecutive iup).
// Same structure as k_colpass_tile with the roles swapped: the 8 batch entries of one dw state
// are 128 contiguous bytes in global memory, so staging, the gathers (tile[j][0..7]), the
// off-block gathers from L2 and the read-modify-write of out are all full 128-byte lines.
//   out(i, c) += sum_k Hd(c, j_k) v(i, j_k)
// ------------------------------------------------------------------------------------
template <bool REALH, bool DIRECT>
__global__ void __launch_bounds__(1024, 1) k_rowpass_tile(int64_t n /*rows = DimUp*/, const double2 *__restrict__ v,
                                                           double2 *__restrict__ out, const int2 *__restrict__ blocks,
                                                           int nblocks, const int32_t *__restrict__ rowptr,
                                                           const int32_t *__restrict__ col,
                                                           const double2 *__restrict__ val, OpArgs op) {
  extern __shared__ double2 tile[];
  const int blk = blockIdx.x % nblocks;
  const int64_t i0 = (int64_t)(blockIdx.x / nblocks) * 8;
  const int2 b = blocks[blk];
  const int g0 = b.x, ng = b.y;
  const int nb = (int)min((int64_t)8, n - i0);
  for (int idx = threadIdx.x; idx < ng * 8; idx += blockDim.x) {
    const int bb = idx & 7, g = idx >> 3;
    if (bb < nb) cp_async16(&tile[idx], v + (i0 + bb) + (int64_t)(g0 + g) * n);
  }
  cp_async_wait_all();
  __syncthreads();
  const int c = threadIdx.x & 7;
  const int rsub = threadIdx.x >> 3, rstep = blockDim.x >> 3;
  const int64_t ig = i0 + min(c, nb - 1);
  const unsigned gmask = 0xFFu << (threadIdx.x & 24);
  for (int g = rsub; g < ng; g += rstep) {
    const int64_t cdw = g0 + g;  // dw state = global column
    double2 acc = make_double2(0.0, 0.0);
    if (!DIRECT) {
      const int32_t p0 = __ldg(rowptr + cdw), p1 = __ldg(rowptr + cdw + 1);
      for (int32_t base = p0; base < p1; base += 8) {
        const int32_t my = base + c;
        int32_t jc = 0;
        double2 hv = make_double2(0.0, 0.0);
        if (my < p1) { jc = __ldg(col + my); hv = ldg2(val + my); }
        const int cnt = min(8, p1 - base);
        for (int kk = 0; kk < cnt; kk++) {
          const int32_t j = __shfl_sync(gmask, jc, kk, 8);
          const double hx = __shfl_sync(gmask, hv.x, kk, 8);
          double hy = 0.0;
          if (!REALH) hy = __shfl_sync(gmask, hv.y, kk, 8);
          const unsigned rel = (unsigned)(j - g0);
          double2 x;
          if (rel < (unsigned)ng) x = tile[rel * 8 + c];
          else x = ldg2(v + ig + (int64_t)j * n);
          if (REALH) rfma(acc, hx, x); else cfma(acc, make_double2(hx, hy), x);
        }
      }
    } else {
      const uint32_t s = (uint32_t)__ldg(op.map + cdw);
      for (int t = 0; t < op.nterms; t++) {
        const Term tm = op.terms[t];
        if (((s >> tm.a) & 1u) && !((s >> tm.b) & 1u)) {
          const uint32_t m = (s & ~(1u << tm.a)) | (1u << tm.b);
          const int32_t j = lin_rank_d(op.lin_lo, op.lin_hi, op.lbits, m);
          const double sg = hop_sign_d(s, tm.a, tm.b);
          const unsigned rel = (unsigned)(j - g0);
          double2 x;
          if (rel < (unsigned)ng) x = tile[rel * 8 + c];
          else x = ldg2(v + ig + (int64_t)j * n);
          if (REALH) rfma(acc, tm.re * sg, x); else cfma(acc, make_double2(tm.re * sg, tm.im * sg), x);
        }
      }
    }
    if (c < nb) {
      double2 *o = out + ig + cdw * n;
      double2 y = *o;
      y.x += acc.x;
      y.y += acc.y;
      *o = y;
    }
  }
}

// ------------------------------------------------------------------------------------
// Row pass, L1-blocked variant: one CTA (32 warps) owns 32 consecutive iup rows x one block of
// dw states sharing their top bits.  Lanes run along iup, so every gather is one coalesced 512-byte
// line set and the operator row is warp-uniform (broadcast loads, no shuffles).  All warps of the
// CTA work on the same 32 rows, so the block's slice of v (block x 512 B, <= ~128 KB) stays in the
// SM's L1 after the first touch: hops inside the block hit L1, only hops that change the top bits
// go to L2.  No shared memory, no staging, no barriers.
//   out(i, c) += sum_k Hd(c, j_k) v(i, j_k)
// ------------------------------------------------------------------------------------
template <bool REALH>
__global__ void __launch_bounds__(1024, 1) k_rowpass_l1(int64_t n /*rows = DimUp*/, const double2 *__restrict__ v,
                                                         double2 *__restrict__ out, const int2 *__restrict__ blocks,
                                                         int nblocks, const int32_t *__restrict__ rowptr,
                                                         const int32_t *__restrict__ col,
                                                         const double2 *__restrict__ val) {
  const int blk = blockIdx.x % nblocks;
  const int64_t i = (int64_t)(blockIdx.x / nblocks) * 32 + (threadIdx.x & 31);
  const int2 b = blocks[blk];
  const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const bool live = i < n;
  const double2 *vi = v + (live ? i : n - 1);
  for (int g = warp; g < b.y; g += nwarps) {
    const int64_t c = b.x + g;
    const int32_t p0 = __ldg(rowptr + c), p1 = __ldg(rowptr + c + 1);
    double2 acc0 = make_double2(0.0, 0.0), acc1 = make_double2(0.0, 0.0);
    int32_t p = p0;
    for (; p + 4 <= p1; p += 4) {
      const double2 x0 = ldg2(vi + (int64_t)__ldg(col + p) * n);
      const double2 x1 = ldg2(vi + (int64_t)__ldg(col + p + 1) * n);
      const double2 x2 = ldg2(vi + (int64_t)__ldg(col + p + 2) * n);
      const double2 x3 = ldg2(vi + (int64_t)__ldg(col + p + 3) * n);
      const double2 h0 = ldg2(val + p), h1 = ldg2(val + p + 1), h2 = ldg2(val + p + 2), h3 = ldg2(val + p + 3);
      if (REALH) { rfma(acc0, h0.x, x0); rfma(acc1, h1.x, x1); rfma(acc0, h2.x, x2); rfma(acc1, h3.x, x3); }
      else { cfma(acc0, h0, x0); cfma(acc1, h1, x1); cfma(acc0, h2, x2); cfma(acc1, h3, x3); }
    }
    for (; p < p1; p++) {
      const double2 x = ldg2(vi + (int64_t)__ldg(col + p) * n);
      const double2 h = ldg2(val + p);
      if (REALH) rfma(acc0, h.x, x); else cfma(acc0, h, x);
    }
    if (live) {
      double2 *o = out + i + c * n;
      double2 y = *o;
      y.x += acc0.x + acc1.x;
      y.y += acc0.y + acc1.y;
      *o = y;
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

// packed tile kernel: ld = leading dimension of v, nbatch = number of batch entries (columns for the
// column pass, iup rows for the row pass)
template <bool REALH, bool COLMODE>
static int launch_tile_pk(const SpinOp &s, int64_t ld, int64_t nbatch, const double2 *v, double2 *out, const DiagArgs &dg) {
  Ctx &c = ctx();
  const size_t smem = ((size_t)s.max_block * 8 + 128) * sizeof(double2);
  static size_t configured = 0;
  if (smem > configured) {
    CB_CUDA(cudaFuncSetAttribute(k_tile_pk<REALH, COLMODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const int64_t nct = ((nbatch + 7) / 8) * s.nblocks;
  if (nct > 0x7fffffffLL) return fail("tile_pk: grid too large");
  int threads = (int)std::min<int64_t>(1024, std::max<int64_t>(256, ((int64_t)s.max_block * 8 + 255) / 256 * 256));
  k_tile_pk<REALH, COLMODE><<<(unsigned)nct, threads, smem, c.stream>>>(ld, nbatch, v, out, s.blocks, s.nblocks, s.pk_in_ptr, s.pk_in, s.pk_off_ptr, s.pk_off,
                                                                          s.coef, dg);
  c.launches++;
  return 0;
}

template <bool REALH, bool DIRECT>
static int launch_colpass_tile_t(const SpinOp &s, int64_t ncols, const double2 *v, double2 *out, const DiagArgs &dg) {
  Ctx &c = ctx();
  const size_t smem = (size_t)s.max_block * 8 * sizeof(double2);
  static size_t configured = 0;  // per instantiation
  if (smem > configured) {
    CB_CUDA(cudaFuncSetAttribute(k_colpass_tile<REALH, DIRECT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const int64_t ngroups = (ncols + 7) / 8;
  const int64_t nct = ngroups * s.nblocks;
  if (nct > 0x7fffffffLL) return fail("colpass_tile: grid too large");
  // enough threads to cover the biggest block once, in multiples of 256 (8 lanes per row)
  int threads = (int)std::min<int64_t>(1024, std::max<int64_t>(256, ((int64_t)s.max_block * 8 + 255) / 256 * 256));
  k_colpass_tile<REALH, DIRECT><<<(unsigned)nct, threads, smem, c.stream>>>(s.n, ncols, v, out, s.blocks, s.nblocks, s.rowptr,
                                                                             s.col, s.val, op_args(s), dg);
  c.launches++;
  return 0;
}

static int launch_colpass_tile(const SpinOp &s, int64_t ncols, const double2 *v, double2 *out, const DiagArgs &dg) {
  Ctx &c = ctx();
  const bool direct = c.mode == CDMFT_B200_DIRECT;
  if (c.real_h) return direct ? launch_colpass_tile_t<true, true>(s, ncols, v, out, dg) : launch_colpass_tile_t<true, false>(s, ncols, v, out, dg);
  return direct ? launch_colpass_tile_t<false, true>(s, ncols, v, out, dg) : launch_colpass_tile_t<false, false>(s, ncols, v, out, dg);
}

static int colpass_impl(const SpinOp &s, int64_t ncols, const double2 *v, double2 *out, const DiagArgs &dg);
static int colpass(const SpinOp &s, int64_t ncols, const double2 *v, double2 *out, const DiagArgs &dg) {
  prof_begin(0);
  int rc = colpass_impl(s, ncols, v, out, dg);
  prof_end();
  return rc;
}
static int colpass_impl(const SpinOp &s, int64_t ncols, const double2 *v, double2 *out, const DiagArgs &dg) {
  if (ncols <= 0 || s.n <= 0) return 0;
  // variant 6 = column-resident shared-memory kernel (default), 1 = generic global-gather kernel,
  // 0/2 = shared-memory tiles, 4/5 = rotating-slot tiles
  int64_t var = ctx().opt.colpass_variant;
  if (var == 6) {
    // forced block split (tests) -> block-split kernel; else whole column; else block split for big columns
    int rc = ctx().opt.colres_rows > 0 ? launch_colblk<double2>(s, ncols, v, out, dg) : kColresNA;
    if (rc == kColresNA) rc = launch_colres<double2>(s, ncols, v, out, dg);
    if (rc == kColresNA) rc = launch_colblk<double2>(s, ncols, v, out, dg);
    if (rc != kColresNA) return rc;
    var = 1;  // does not apply (DIRECT mode, column larger than shared memory): generic kernel
  }
  if ((var == 4 || var == 5) && ctx().mode == CDMFT_B200_SPARSE && s.pkell && s.nblocks > 0 &&
      (size_t)s.max_block * 128 + 2048 <= 232448) {
    Ctx &c = ctx();
    // variant 5: only the in-block terms here; the caller's row pass adds the off-block ones (UPOFF)
    const bool inonly = var == 5 && dg.enabled && !(c.spmd || c.sim || c.opt.force_sharded);
    const size_t smem = ((size_t)s.max_block * 8 + 128) * sizeof(double2);
    auto kern = c.real_h ? (inonly ? k_colpass_rot<true, true> : k_colpass_rot<true, false>)
                         : (inonly ? k_colpass_rot<false, true> : k_colpass_rot<false, false>);
    static size_t configured[4] = {0, 0, 0, 0};
    const int slot = (c.real_h ? 2 : 0) + (inonly ? 1 : 0);
    if (smem > configured[slot]) {
      CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured[slot] = smem;
    }
    const int64_t nct = ((ncols + 7) / 8) * s.nblocks;
    if (nct > 0x7fffffffLL) return fail("colpass_rot: grid too large");
    const int niter = (s.max_block + 575) / 576;
    int threads = ((s.max_block + niter - 1) / niter + 31) / 32 * 32;
    threads = std::max(64, std::min(576, threads));
    kern<<<(unsigned)nct, threads, smem, c.stream>>>(s.n, ncols, v, out, s.blocks, s.nblocks, s.pkell, s.rowlen, s.rowsplit, s.coef, dg);
    c.launches++;
    return 0;
  }
  if (var != 1 && var != 4 && var != 5 && s.nblocks > 0 && (size_t)s.max_block * 128 + 2048 <= 232448) {
    if (var != 2 && ctx().mode == CDMFT_B200_SPARSE && s.pk_in && s.pk_swizzled)
      return ctx().real_h ? launch_tile_pk<true, true>(s, s.n, ncols, v, out, dg) : launch_tile_pk<false, true>(s, s.n, ncols, v, out, dg);
    return launch_colpass_tile(s, ncols, v, out, dg);
  }
  switch (ctx().opt.col_batch) {
    case 1: return launch_colpass<1>(s, ncols, v, out, dg);
    case 2: return launch_colpass<2>(s, ncols, v, out, dg);
    case 8: return launch_colpass<8>(s, ncols, v, out, dg);
    default: return launch_colpass<4>(s, ncols, v, out, dg);
  }
}

template <bool REALH, bool DIRECT>
static int launch_rowpass_tile_t(const SpinOp &s, int64_t nrows, const double2 *v, double2 *out) {
  Ctx &c = ctx();
  const size_t smem = (size_t)s.max_block * 8 * sizeof(double2);
  static size_t configured = 0;
  if (smem > configured) {
    CB_CUDA(cudaFuncSetAttribute(k_rowpass_tile<REALH, DIRECT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const int64_t nct = ((nrows + 7) / 8) * s.nblocks;
  if (nct > 0x7fffffffLL) return fail("rowpass_tile: grid too large");
  int threads = (int)std::min<int64_t>(1024, std::max<int64_t>(256, ((int64_t)s.max_block * 8 + 255) / 256 * 256));
  k_rowpass_tile<REALH, DIRECT><<<(unsigned)nct, threads, smem, c.stream>>>(nrows, v, out, s.blocks, s.nblocks, s.rowptr, s.col,
                                                                             s.val, op_args(s));
  c.launches++;
  return 0;
}


// ------------------------------------------------------------------------------------
// Row pass, block-resident (SPARSE mode, rowpass_variant 4; tested, NOT default -- measured 3.9 ms vs 2.7 ms for
// the generic L2-slab kernel at K3: latency-bound, and the off-block gathers cost twice an in-block one):
//   out(i,c) += sum_k Hd(c,j_k) v(i,j_k).
// One CTA owns (block of columns sharing their top bits) x (8 consecutive rows i0..i0+7).  The tile
// v(i0..i0+7, block) -- one 128-byte line per column -- is brought into shared memory by TMA bulk copies
// (cp.async.bulk + mbarrier; no registers, no L1 wavefronts).  A warp task is 4 columns, one per 8-lane group;
// the 8 lanes of a group are the 8 rows.  Entries whose source column lies inside the block read the tile
// (one conflict-free 128-byte line per group and step); entries that change the top bits read global memory /
// L2 (one line per group).  The generic kernel fetches every v element ~15 times from L2; here the in-block
// share (60-75 %) is fetched once per tile.  Operator words are warp-task streams built on the host
// (RowRes, ctx.h / build_rowres, sector.cu), shared by all row strips of a block.
// MODE: 0 = coefficient table, complex; 1 = coefficient table, real; 2 = sign/class bits (fast).
// ------------------------------------------------------------------------------------
struct RowResArgs {
  const int2 *blocks;
  const int32_t *tbase;
  const uint4 *task;
  const int32_t *task_col;
  const uint4 *win;
  const uint32_t *woff;
  const double2 *coef;
  double m0, m1;
  int nblocks;
};

template <int MODE>
__device__ __forceinline__ void rowres_fma(double2 &acc, uint32_t w, double2 x, const char *coef_b, double m0, double m1) {
  if (MODE == 2) rfma(acc, colres_signed((w & 1u) ? m1 : m0, w & 0x80000000u), x);
  else if (MODE == 1) rfma(acc, *(const double *)(coef_b + ((w & 127u) << 4)), x);
  else cfma(acc, *(const double2 *)(coef_b + ((w & 127u) << 4)), x);
}

template <int MODE>
__device__ __forceinline__ bool rowres_on(uint32_t w) { return MODE == 2 ? (w != 0xFFFFFFFFu) : ((w & 127u) != 0u); }
template <int MODE>
__device__ __forceinline__ int64_t rowres_col(uint32_t w) { return MODE == 2 ? ((w & 0x7FFFFFFFu) >> 1) : (w >> 7); }

template <int MODE>
__global__ void __launch_bounds__(384, 3) k_rowres(int64_t n /*rows = DimUp*/, const double2 *__restrict__ v,
                                                    double2 *__restrict__ out, RowResArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t *bar = (uint64_t *)smem_raw;
  double2 *coef = (double2 *)(smem_raw + 128);
  double2 *tile = (double2 *)(smem_raw + 128 + 2048);  // [ng+1][8], the last line stays zero
  const int blk = blockIdx.x % a.nblocks;
  const int64_t i0 = (int64_t)(blockIdx.x / a.nblocks) * 8;
  const int2 b = a.blocks[blk];
  const int g0 = b.x, ng = b.y;
  const int nb = (int)min((int64_t)8, n - i0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int r = lane & 7, grp = lane >> 3;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_expect_tx(bar, (uint32_t)ng * (uint32_t)nb * 16u);
  }
  if (threadIdx.x < 128) coef[threadIdx.x] = MODE == 2 ? make_double2(0.0, 0.0) : a.coef[threadIdx.x];
  if (threadIdx.x < 8) tile[(size_t)ng * 8 + threadIdx.x] = make_double2(0.0, 0.0);
  if (nb < 8)  // ragged last strip: rows past the end read as zero
    for (int k = threadIdx.x; k < ng * 8; k += blockDim.x)
      if ((k & 7) >= nb) tile[k] = make_double2(0.0, 0.0);
  __syncthreads();
  for (int cidx = threadIdx.x; cidx < ng; cidx += blockDim.x)
    bulk_g2s(tile + (size_t)cidx * 8, v + (int64_t)(g0 + cidx) * n + i0, (uint32_t)nb * 16u, bar);
  const int t0 = __ldg(a.tbase + blk), t1 = __ldg(a.tbase + blk + 1);
  const char *tile_b = (const char *)tile + r * 16;
  const char *coef_b = (const char *)coef;
  const int rc = min(r, nb - 1);  // clamped on the ragged last strip (never stored)
  const double2 *vrow = v + i0 + rc;
  bool waited = false;
  for (int t = t0 + warp; t < t1; t += nwarps) {
    const uint4 tk = __ldg(a.task + t);
    const int cl = __ldg(a.task_col + t * 4 + grp);
    double2 *o = out + (int64_t)(g0 + max(cl, 0)) * n + i0 + rc;
    const double2 y = *o;  // read-modify-write operand: requested first, needed last
    double2 acc = make_double2(0.0, 0.0);
    {  // ---- sources outside the block (long latency): global memory / L2, one 128-byte line per group
      const uint32_t *wo = a.woff + (int64_t)tk.z * 4 + grp;
      const int noff = (int)tk.w;
      int k = 0;
      for (; k + 2 <= noff; k += 2) {
        const uint32_t w0 = __ldg(wo + k * 4), w1 = __ldg(wo + k * 4 + 4);
        const bool on0 = rowres_on<MODE>(w0), on1 = rowres_on<MODE>(w1);
        double2 x0 = make_double2(0.0, 0.0), x1 = make_double2(0.0, 0.0);
        if (on0) x0 = ldg2(vrow + rowres_col<MODE>(w0) * n);
        if (on1) x1 = ldg2(vrow + rowres_col<MODE>(w1) * n);
        if (on0) rowres_fma<MODE>(acc, w0, x0, coef_b, a.m0, a.m1);
        if (on1) rowres_fma<MODE>(acc, w1, x1, coef_b, a.m0, a.m1);
      }
      if (k < noff) {
        const uint32_t w0 = __ldg(wo + k * 4);
        if (rowres_on<MODE>(w0)) rowres_fma<MODE>(acc, w0, ldg2(vrow + rowres_col<MODE>(w0) * n), coef_b, a.m0, a.m1);
      }
    }
    if (!waited) { mbar_wait(bar, 0); waited = true; }
    {  // ---- sources inside the block: shared memory, four steps per operator load
      const uint4 *wi = a.win + (int64_t)tk.x * 4 + grp;
      const int nq = (int)tk.y;
      uint4 wn = __ldg(wi);  // slack behind the stream: always readable
      for (int q = 0; q < nq; q++) {
        const uint4 w = wn;
        wn = __ldg(wi + (q + 1) * 4);
        rowres_fma<MODE>(acc, w.x, *(const double2 *)(tile_b + (w.x & 0x7FFFFF80u)), coef_b, a.m0, a.m1);
        rowres_fma<MODE>(acc, w.y, *(const double2 *)(tile_b + (w.y & 0x7FFFFF80u)), coef_b, a.m0, a.m1);
        rowres_fma<MODE>(acc, w.z, *(const double2 *)(tile_b + (w.z & 0x7FFFFF80u)), coef_b, a.m0, a.m1);
        rowres_fma<MODE>(acc, w.w, *(const double2 *)(tile_b + (w.w & 0x7FFFFF80u)), coef_b, a.m0, a.m1);
      }
    }
    if (cl >= 0 && r < nb) *o = make_double2(y.x + acc.x, y.y + acc.y);
  }
  if (!waited) mbar_wait(bar, 0);  // a warp without tasks still must not leave before the copies have landed
}

static int launch_rowres(const SpinOp &s, int64_t nrows, const double2 *v, double2 *out) {
  Ctx &c = ctx();
  const RowRes &rr = s.rr;
  if (c.mode != CDMFT_B200_SPARSE || !rr.win || rr.ntask <= 0) return kColresNA;
  const size_t smem = 128 + 2048 + ((size_t)rr.max_block + 1) * 128;
  if (smem > 232448) return kColresNA;
  RowResArgs a{};
  a.blocks = rr.blocks; a.tbase = rr.tbase; a.task = rr.task; a.task_col = rr.task_col;
  a.win = (const uint4 *)rr.win; a.woff = rr.woff; a.coef = s.coef; a.m0 = s.sc_mag[0]; a.m1 = s.sc_mag[1];
  a.nblocks = rr.nblocks;
  void (*kern)(int64_t, const double2 *, double2 *, RowResArgs) =
      rr.fmt == 1 ? k_rowres<2> : (c.real_h ? k_rowres<1> : k_rowres<0>);
  CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t nct = ((nrows + 7) / 8) * rr.nblocks;
  if (nct > 0x7fffffffLL) return fail("rowres: grid too large");
  // 384 threads x 3 CTAs per SM at K3 (59 KB tiles): the tile load of one CTA overlaps the compute of the others
  const int threads = (int)std::min<int64_t>(384, std::max<int64_t>(64, ((int64_t)rr.max_block / 4 / 4 + 1) * 32));
  kern<<<(unsigned)nct, threads, smem, c.stream>>>(nrows, v, out, a);
  c.launches++;
  return 0;
}

static int rowpass_impl(const SpinOp &s, int64_t nrows, const double2 *v, double2 *out);
static int rowpass(const SpinOp &s, int64_t nrows, const double2 *v, double2 *out) {
  prof_begin(1);
  int rc = rowpass_impl(s, nrows, v, out);
  prof_end();
  return rc;
}
static int rowpass_impl(const SpinOp &s, int64_t nrows, const double2 *v, double2 *out) {
  Ctx &c = ctx();
  if (s.n <= 0 || nrows <= 0) return 0;
  // colpass_variant 5 leaves the off-block Hup terms to the generic row-pass kernel
  int64_t rv = c.opt.colpass_variant == 5 ? 1 : c.opt.rowpass_variant;
  if (rv == 4) {
    const int rc = launch_rowres(s, nrows, v, out);
    if (rc != kColresNA) return rc;
    rv = 1;  // DIRECT mode / no operator streams: generic kernel
  }
  if (rv == 3 && c.mode == CDMFT_B200_SPARSE && s.nblocks_l1 > 0) {
    static bool configured = false;
    if (!configured) {  // all of the unified L1/shared array as L1
      cudaFuncSetAttribute(k_rowpass_l1<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
      cudaFuncSetAttribute(k_rowpass_l1<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
      configured = true;
    }
    const int64_t nct = ((nrows + 31) / 32) * s.nblocks_l1;
    if (nct > 0x7fffffffLL) return fail("rowpass_l1: grid too large");
    if (c.real_h) k_rowpass_l1<true><<<(unsigned)nct, 1024, 0, c.stream>>>(nrows, v, out, s.blocks_l1, s.nblocks_l1, s.rowptr, s.col, s.val);
    else k_rowpass_l1<false><<<(unsigned)nct, 1024, 0, c.stream>>>(nrows, v, out, s.blocks_l1, s.nblocks_l1, s.rowptr, s.col, s.val);
    c.launches++;
    return 0;
  }
  if (rv != 1 && s.nblocks > 0 && (size_t)s.max_block * 128 + 2048 <= 232448) {
    const bool direct = c.mode == CDMFT_B200_DIRECT;
    if (rv != 2 && !direct && s.pk_in && !s.pk_swizzled) {
      DiagArgs nodiag{};
      return c.real_h ? launch_tile_pk<true, false>(s, nrows, nrows, v, out, nodiag) : launch_tile_pk<false, false>(s, nrows, nrows, v, out, nodiag);
    }
    if (c.real_h) return direct ? launch_rowpass_tile_t<true, true>(s, nrows, v, out) : launch_rowpass_tile_t<true, false>(s, nrows, v, out);
    return direct ? launch_rowpass_tile_t<false, true>(s, nrows, v, out) : launch_rowpass_tile_t<false, false>(s, nrows, v, out);
  }
  dim3 grid((unsigned)s.n, (unsigned)((nrows + 255) / 256));
  if (grid.y > 65535) return fail("rowpass: too many row chunks");
  OpArgs op = op_args(s);
  const bool direct = c.mode == CDMFT_B200_DIRECT;
  UpOffArgs uo{};
  const SpinOp &u = c.up;
  const bool upoff = c.opt.colpass_variant == 5 && !direct && u.pkell && u.rowsplit && u.nblocks > 0 &&
                     (size_t)u.max_block * 128 + 2048 <= 232448 && &s == &c.dw;
  if (upoff) { uo.pkell = u.pkell; uo.rowlen = u.rowlen; uo.rowsplit = u.rowsplit; uo.coef = u.coef; }
  if (!direct && !upoff && (c.opt.row_rb > 1 || c.opt.row_slab != 256)) {
    // RB row chunks per thread: grid.y covers 256*RB rows per CTA
    // the slab of rows one grid.y index sweeps (threads*RB rows x all columns) must stay L2-resident:
    // 256 rows x 12870 columns x 16 B = 53 MB at K3 -> threads = row_slab / RB
    const int rb = c.opt.row_rb >= 4 ? 4 : (c.opt.row_rb >= 2 ? 2 : 1);
    const int slab = (int)std::max<int64_t>(64, std::min<int64_t>(1024, c.opt.row_slab));
    const int thr = std::max(32, slab / rb / 32 * 32);
    dim3 g2((unsigned)s.n, (unsigned)((nrows + (int64_t)thr * rb - 1) / ((int64_t)thr * rb)));
    // fused Lanczos dot (lanczos.cu sets dot_request): one partial per CTA
    double *dp = nullptr;
    if (c.dot_request) {
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
    if (direct) k_rowpass<true, true, false><<<grid, 256, 0, c.stream>>>(nrows, s.n, v, out, s.rowptr, s.col, s.val, op, uo);
    else if (upoff) k_rowpass<true, false, true><<<grid, 256, 0, c.stream>>>(nrows, s.n, v, out, s.rowptr, s.col, s.val, op, uo);
    else k_rowpass<true, false, false><<<grid, 256, 0, c.stream>>>(nrows, s.n, v, out, s.rowptr, s.col, s.val, op, uo);
  } else {
    if (direct) k_rowpass<false, true, false><<<grid, 256, 0, c.stream>>>(nrows, s.n, v, out, s.rowptr, s.col, s.val, op, uo);
    else if (upoff) k_rowpass<false, false, true><<<grid, 256, 0, c.stream>>>(nrows, s.n, v, out, s.rowptr, s.col, s.val, op, uo);
    else k_rowpass<false, false, false><<<grid, 256, 0, c.stream>>>(nrows, s.n, v, out, s.rowptr, s.col, s.val, op, uo);
  }
  c.launches++;
  return 0;
}

// Row pass of Hdw on a REAL vector with an even number of rows: adjacent rows (i, i+1) of the real vector are
// one double2 of a "complex" vector with nrows/2 rows, and a real coefficient acts on both halves alike, so the
// complex kernels run unchanged on half as many rows: every gather moves 16 bytes per lane instead of 8.
int rowpass_real_as_pairs(int64_t nrows, const double *v, double *out) {
  Ctx &c = ctx();
  if ((nrows & 1) || !c.real_h) return fail("rowpass_real_as_pairs: needs a real H and an even DimUp");
  return rowpass(c.dw, nrows / 2, (const double2 *)v, (double2 *)out);
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
int colpass_real(const SpinOp &s, int64_t ncols, const double *v, double *out, const DiagArgs &dg);  // hxv_real.cu

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
    // allgather_vector_MPI (ED_SETUP.f90:672-708): every rank sends its shard to every rank
    if (c.rk.empty()) return 0;
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
    vfull = c.vfull;
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
    // one rank: diag + up (column pass), then dw on the strided index (row pass)
    CB_CHECK(colpass(c.up, c.dimdw, v, hv, diag_args(0)));
    CB_CHECK(rowpass(c.dw, c.dimup, v, hv));
    return 0;
  }
  // sharded: diag -> UP -> transpose -> DW on vt -> transpose back -> add  (spMatVec_mpi_main order)
  int64_t off = 0;
  std::vector<int64_t> offs;
  const int P = c.p_eff;
  // measured on B200 (K3): peer-memory transposes win at P=2 (8.4 vs 8.7 ms), the overlapped NCCL path wins
  // at P=4 (4.6 vs 4.9) and P=8 (2.5 vs 2.8) -> use_ipc 1 = auto (P<=2), 2 = always, 0 = never
  const bool use_ipc = c.spmd && P > 1 && !c.rk.empty() && c.ipc_ready &&
                       (c.opt.use_ipc == 2 || (c.opt.use_ipc == 1 && P <= 2));
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
    if (pairs) CB_CHECK(colpass_real(c.up, r.dw.q, (const double *)(v + off), (double *)(hv + off), diag_args(r.dw.off)));
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
