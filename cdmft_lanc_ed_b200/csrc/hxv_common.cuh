// hxv_common.cuh -- device helpers shared by the H x v translation units (hxv.cu, hxv_real.cu)
#pragma once
#include <algorithm>
#include <map>
#include <type_traits>
#include <utility>

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
// Column-resident column pass (SPARSE mode; the default when a column fits in shared memory).
//   out(i,c) = [diag] d(i,c) v(i,c) + sum_k H(i,j_k) v(j_k,c)
// A persistent CTA per SM walks over columns c = blockIdx.x, +gridDim.x, ...: the whole column
// (DimUp x sizeof(T): 206 KB for complex(8) at Ns=16) is brought into shared memory by TMA bulk copies
// (cp.async.bulk + mbarrier, no registers, no L1 wavefronts), so v is read from HBM exactly once and
// EVERY gather of the operator is a shared-memory read -- 128 B/clk/SM instead of the 64 B/clk of the
// L1 path that bounds the generic kernels.  The gathers follow the edge-coloured schedule (Sched,
// ctx.h): at every step the G lanes of a shared-memory phase read G different banks, and a warp's
// four (two) row groups have about the same number of steps.  Operator words stream from L2,
// coalesced, two steps per 8-byte load, prefetched two loads ahead.
// T = double2 (G = 8) or double (G = 16, real Krylov mode).  FAST: real H with <= 2 |coefficients|,
// decoded with selects; otherwise a 128-entry coefficient table in shared memory.
// ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)),
               "l"((uint64_t)__cvta_generic_to_global(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
  } while (!done);
}

struct ColResArgs {
  const int32_t *tbase, *qbase;  // [nwarps+1]
  const uint4 *meta;             // [ntask*32]
  const uint32_t *words;
  const double2 *coef;  // [128] (table decode)
  double m0, m1;        // fast decode (m0..m3 for the four-class variant)
  double m2, m3;
  int accum;            // 1: out += (the row pass wrote out first), 0: out =
  double *dot_partial;  // != nullptr: Re<v,out> of the final out, one partial per CTA (fused Lanczos alpha)
};
__device__ __forceinline__ double colres_dotre(double2 x, double2 y) { return fma(x.x, y.x, x.y * y.y); }
__device__ __forceinline__ double colres_dotre(double x, double y) { return x * y; }

__device__ __forceinline__ void colres_fma(double2 &acc, double h, double2 x) { rfma(acc, h, x); }
__device__ __forceinline__ void colres_fma(double &acc, double h, double x) { acc = fma(h, x, acc); }
__device__ __forceinline__ void colres_cfma(double2 &acc, double2 h, double2 x) { cfma(acc, h, x); }
__device__ __forceinline__ void colres_cfma(double &acc, double2 h, double x) { acc = fma(h.x, x, acc); }
__device__ __forceinline__ void colres_zero(double2 &a) { a = make_double2(0.0, 0.0); }
__device__ __forceinline__ void colres_zero(double &a) { a = 0.0; }
__device__ __forceinline__ double2 operator+(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 colres_scale(double d, double2 x) { return make_double2(d * x.x, d * x.y); }
__device__ __forceinline__ double colres_scale(double d, double x) { return d * x; }

// decode modes: 0 = coefficient table, complex values; 1 = coefficient table, real values;
// 2 = fast 32-bit words; 3 = fast 16-bit words; 4 = fast4: purely real or purely imaginary coefficients of up
// to four magnitudes (formats in sector.cu, build_schedule_host)
__device__ __forceinline__ double2 colres_times_i(double2 x) { return make_double2(-x.y, x.x); }
__device__ __forceinline__ double colres_times_i(double x) { return x; }  // real vectors never carry the phase bit
__device__ __forceinline__ double colres_signed(double m, uint32_t signbit31) {
  return __hiloint2double(__double2hiint(m) ^ (int)signbit31, __double2loint(m));
}
template <typename T, int MODE>
__device__ __forceinline__ void colres_step(T &acc, uint32_t w, const char *xs, const char *coef_b, double m0, double m1,
                                            double m2 = 0.0, double m3 = 0.0) {
  if (MODE == 4) {
    // w = (negative << 31) | byte offset | imaginary << 2 | class : h = +-m or +-i*m
    T x = *(const T *)(xs + (w & 0x7FFFFFF8u & ~(uint32_t)(sizeof(T) - 1)));
    const double ma = (w & 1u) ? m1 : m0, mb = (w & 1u) ? m3 : m2;
    if (w & 4u) x = colres_times_i(x);
    colres_fma(acc, colres_signed((w & 2u) ? mb : ma, w & 0x80000000u), x);
  } else if (MODE == 2) {
    // w = (negative << 31) | byte offset | class : branch-free, idle lanes read a zero element
    const T x = *(const T *)(xs + (w & 0x7FFFFFF8u & ~(uint32_t)(sizeof(T) - 1)));
    colres_fma(acc, colres_signed((w & 1u) ? m1 : m0, w & 0x80000000u), x);
  } else {
    const T x = *(const T *)(xs + (size_t)(w >> 7) * sizeof(T));
    if (MODE == 1) colres_fma(acc, *(const double *)(coef_b + ((w & 127u) << 4)), x);
    else colres_cfma(acc, *(const double2 *)(coef_b + ((w & 127u) << 4)), x);
  }
}
// two 16-bit words (negative << 15 | class << 14 | source row) in one register
template <typename T>
__device__ __forceinline__ void colres_step16x2(T &acc, uint32_t w, const char *xs, double m0, double m1) {
  constexpr int SH = sizeof(T) == 16 ? 4 : 3;
  constexpr uint32_t AM = 0x3FFFu << SH;
  const T x0 = *(const T *)(xs + ((w << SH) & AM));
  const T x1 = *(const T *)(xs + ((w >> (16 - SH)) & AM));
  colres_fma(acc, colres_signed((w & 0x4000u) ? m1 : m0, (w << 16) & 0x80000000u), x0);
  colres_fma(acc, colres_signed((w & 0x40000000u) ? m1 : m0, w & 0x80000000u), x1);
}

template <typename T, int MODE>
__global__ void __launch_bounds__(1024, 1) k_colres(int64_t n, int64_t ncols, const T *__restrict__ v, T *__restrict__ out,
                                                     ColResArgs a, DiagArgs dg) {
  constexpr int G = sizeof(T) == 16 ? 8 : 16;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // layout: [0,8) mbarrier | [128, 128+2048) coefficient table | dtab [2^nimp doubles] | column + G zero elements
  uint64_t *bar = (uint64_t *)smem_raw;
  double2 *coef = (double2 *)(smem_raw + 128);
  double *dtab = (double *)(smem_raw + 128 + 2048);
  const int ndt = dg.enabled ? (1 << dg.nimp) : 0;
  T *xs = (T *)(smem_raw + 128 + 2048 + (((size_t)ndt * 8 + 127) & ~(size_t)127));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t npad = (n + G - 1) / G * G;
  if (threadIdx.x == 0) mbar_init(bar, 1);
  if (threadIdx.x < 128) coef[threadIdx.x] = MODE >= 2 ? make_double2(0.0, 0.0) : a.coef[threadIdx.x];  // table decode only
  for (int64_t k = n + threadIdx.x; k < npad + G; k += blockDim.x) colres_zero(xs[k]);  // idle lanes gather these
  // this warp's stream: tasks [t0,t1) and the quads from q0 on (the same for every column)
  const int t0 = __ldg(a.tbase + warp), t1 = __ldg(a.tbase + warp + 1);
  const int64_t q0 = __ldg(a.qbase + warp);
  const uint4 *mp = a.meta + (int64_t)t0 * 32 + lane;
  using WQ = typename std::conditional<MODE == 3, uint32_t, uint4>::type;  // one unit of a lane: 2 or 4 steps
  const WQ *wbase = (const WQ *)a.words + q0 * 32 + lane;
  __syncthreads();
  const uint32_t bytes = (uint32_t)(n * sizeof(T));
  const char *xs_b = (const char *)xs;
  const char *coef_b = (const char *)coef;
  uint32_t phase = 0;
  double dsum = 0.0;
  for (int64_t c = blockIdx.x; c < ncols; c += gridDim.x) {
    if (threadIdx.x == 0) {
      mbar_expect_tx(bar, bytes);
      const char *src = (const char *)(v + c * n);
      for (uint32_t off = 0; off < bytes; off += 32768u) bulk_g2s((char *)xs + off, src + off, min(32768u, bytes - off), bar);
    }
    for (int mu = threadIdx.x; mu < ndt; mu += blockDim.x) {  // diagonal of this column per impurity configuration of the row
      const int64_t cg = dg.coloff + c;
      double val = __ldg(dg.f_col + cg);
      uint32_t md = (uint32_t)__ldg(dg.map_col + cg) & ((1u << dg.nimp) - 1u);
      while (md) {
        const int b = __ffs(md) - 1;
        md &= md - 1;
        val += __ldg(dg.cross_tab + ((int64_t)b << dg.nimp) + mu);
      }
      dtab[mu] = val;
    }
    // operator stream: nothing of it depends on the column, so the first loads fly while the column arrives
    const WQ *wp = wbase;
    WQ wa = __ldg(wp), wb = __ldg(wp + 32);  // four units of slack behind every stream
    WQ wc = wa, wd = wa;
    if constexpr (MODE == 3) { wc = __ldg(wp + 64); wd = __ldg(wp + 96); }
    uint4 m = t0 < t1 ? __ldg(mp) : make_uint4(0u, 0u, 0xFFFFFFFFu, 0u);
    mbar_wait(bar, phase);
    phase ^= 1u;
    __syncthreads();
    T *oc = out + c * n;
    for (int t = t0; t < t1; t++) {
      uint4 mnext = make_uint4(0u, 0u, 0xFFFFFFFFu, 0u);
      if (t + 1 < t1) mnext = __ldg(mp + (int64_t)(t + 1 - t0) * 32);
      const int nquad = (int)(m.w >> 16);
      const bool valid = m.z != 0xFFFFFFFFu;
      T acc, yold;
      colres_zero(acc);
      colres_zero(yold);
      if (a.accum && valid) yold = oc[m.z];  // requested first, needed last
      if (dg.enabled && valid)
        acc = colres_scale(__hiloint2double((int)m.y, (int)m.x) + dtab[m.w & 0xFFFFu], xs[m.z]);
      for (int kq = 0; kq < nquad; kq++) {
        const WQ w = wa;
        wp += 32;
        if constexpr (MODE == 3) {
          // one 4-byte load (a single 128-byte line per warp) per two steps, four loads in flight
          wa = wb; wb = wc; wc = wd;
          wd = __ldg(wp + 96);
          colres_step16x2<T>(acc, w, xs_b, a.m0, a.m1);
        } else {
          wa = wb;
          wb = __ldg(wp + 32);
          colres_step<T, MODE>(acc, w.x, xs_b, coef_b, a.m0, a.m1, a.m2, a.m3);
          colres_step<T, MODE>(acc, w.y, xs_b, coef_b, a.m0, a.m1, a.m2, a.m3);
          colres_step<T, MODE>(acc, w.z, xs_b, coef_b, a.m0, a.m1, a.m2, a.m3);
          colres_step<T, MODE>(acc, w.w, xs_b, coef_b, a.m0, a.m1, a.m2, a.m3);
        }
      }
      if (valid) {
        acc = acc + yold;
        oc[m.z] = acc;
        if (a.dot_partial) dsum += colres_dotre(xs[m.z], acc);
      }
      m = mnext;
    }
    __syncthreads();  // every gather of this column is done before the next bulk copy lands
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

// ------------------------------------------------------------------------------------
// Block-split column-resident kernel (ColBlk, ctx.h): columns larger than shared memory (Ns = 18).
// Work item = (column, row block); items of one column are adjacent, so its blocks run on neighbouring SMs at
// the same time and the off-block gathers hit L2.  In-block entries: the block's edge-coloured schedule against
// shared memory, exactly as in k_colres; off-block entries: lane-parallel stream, scattered 16-byte (8-byte)
// gathers from the column in global memory.  MODE as in k_colres (16-bit words are not used here).
// ------------------------------------------------------------------------------------
struct ColBlkArgs {
  const int4 *blk;
  const int32_t *tbase, *qbase;
  const uint4 *meta;
  const uint32_t *words;
  const uint2 *toff;
  const uint32_t *woff;
  const double2 *coef;
  double m0, m1, m2, m3;
  int nblk, nwarps;
};

template <int MODE>
__device__ __forceinline__ bool colblk_on(uint32_t w) { return MODE >= 2 ? (w != 0xFFFFFFFFu) : ((w & 127u) != 0u); }
template <typename T, int MODE>
__device__ __forceinline__ T colblk_off_load(uint32_t w, const T *__restrict__ vc) {
  T x;
  colres_zero(x);
  if (colblk_on<MODE>(w)) x = __ldg(vc + (MODE >= 2 ? ((w & 0x7FFFFFFFu) >> 3) : (w >> 7)));
  return x;
}
template <typename T, int MODE>
__device__ __forceinline__ void colblk_off_apply(T &acc, uint32_t w, T x, const char *coef_b, double m0, double m1, double m2,
                                                 double m3) {
  if (!colblk_on<MODE>(w)) return;
  if (MODE >= 2) {
    const double ma = (w & 1u) ? m1 : m0, mb = (w & 1u) ? m3 : m2;
    if (MODE == 4 && (w & 4u)) x = colres_times_i(x);
    colres_fma(acc, colres_signed((MODE == 4 && (w & 2u)) ? mb : ma, w & 0x80000000u), x);
  } else if (MODE == 1) {
    colres_fma(acc, *(const double *)(coef_b + ((w & 127u) << 4)), x);
  } else {
    colres_cfma(acc, *(const double2 *)(coef_b + ((w & 127u) << 4)), x);
  }
}

#ifndef COLBLK_KP
#define COLBLK_KP 2
#endif
constexpr int kColblkKP = COLBLK_KP;  // off-block steps prefetched per task (the rest, if any, is loaded in place)
template <typename T, int MODE>
__global__ void __launch_bounds__(kColblkWarps * 32, 1) k_colblk(int64_t n, int64_t ncols, const T *__restrict__ v, T *__restrict__ out,
                                                                  ColBlkArgs a, DiagArgs dg) {
  constexpr int G = sizeof(T) == 16 ? 8 : 16;
  constexpr int KP = kColblkKP;
  constexpr uint32_t NONE = MODE >= 2 ? 0xFFFFFFFFu : 0u;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t *bar = (uint64_t *)smem_raw;
  double2 *coef = (double2 *)(smem_raw + 128);
  double *dtab = (double *)(smem_raw + 128 + 2048);
  const int ndt = dg.enabled ? (1 << dg.nimp) : 0;
  T *xs = (T *)(smem_raw + 128 + 2048 + (((size_t)ndt * 8 + 127) & ~(size_t)127));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) mbar_init(bar, 1);
  if (threadIdx.x < 128) coef[threadIdx.x] = MODE >= 2 ? make_double2(0.0, 0.0) : a.coef[threadIdx.x];
  __syncthreads();
  const char *xs_b = (const char *)xs;
  const char *coef_b = (const char *)coef;
  uint32_t phase = 0;
  const int64_t nitems = ncols * a.nblk;
  for (int64_t item = blockIdx.x; item < nitems; item += gridDim.x) {
    const int64_t c = item / a.nblk;
    const int b = (int)(item - c * a.nblk);
    const int4 bd = __ldg(a.blk + b);
    const int g0 = bd.x, ng = bd.y;
    const int npad = (ng + G - 1) / G * G;
    const T *vc = v + c * n;
    if (threadIdx.x == 0) {
      const uint32_t bytes = (uint32_t)ng * (uint32_t)sizeof(T);
      mbar_expect_tx(bar, bytes);
      const char *src = (const char *)(vc + g0);
      for (uint32_t off = 0; off < bytes; off += 32768u) bulk_g2s((char *)xs + off, src + off, min(32768u, bytes - off), bar);
    }
    for (int k = ng + threadIdx.x; k < npad + G; k += blockDim.x) colres_zero(xs[k]);  // idle lanes gather these
    for (int mu = threadIdx.x; mu < ndt; mu += blockDim.x) {
      const int64_t cg = dg.coloff + c;
      double val = __ldg(dg.f_col + cg);
      uint32_t md = (uint32_t)__ldg(dg.map_col + cg) & ((1u << dg.nimp) - 1u);
      while (md) {
        const int bb = __ffs(md) - 1;
        md &= md - 1;
        val += __ldg(dg.cross_tab + ((int64_t)bb << dg.nimp) + mu);
      }
      dtab[mu] = val;
    }
    const int t0 = bd.z + __ldg(a.tbase + b * (a.nwarps + 1) + warp), t1 = bd.z + __ldg(a.tbase + b * (a.nwarps + 1) + warp + 1);
    using WQ = typename std::conditional<MODE == 3, uint32_t, uint4>::type;  // one unit of a lane: 2 or 4 steps
    const WQ *wp = (const WQ *)a.words + ((int64_t)bd.w + __ldg(a.qbase + b * (a.nwarps + 1) + warp)) * 32 + lane;
    const uint4 *mp = a.meta + (int64_t)t0 * 32 + lane;
    WQ wa = __ldg(wp), wb = __ldg(wp + 32);
    WQ wc = wa, wd = wa;
    if constexpr (MODE == 3) { wc = __ldg(wp + 64); wd = __ldg(wp + 96); }
    uint4 m = t0 < t1 ? __ldg(mp) : make_uint4(0u, 0u, 0xFFFFFFFFu, 0u);
    // Off-block sources (hops that change the top bits) are scattered gathers from the column in global memory / L2.
    // Nothing of them depends on the staged block, and a load that waits on a load would stall the warp twice: the
    // operator words of task t+1 are requested during task t, the gathers of task t are issued at its top -- for the
    // first task while the block is still in flight -- and consumed after its shared-memory loop.
    uint2 to = t0 < t1 ? __ldg(a.toff + t0) : make_uint2(0u, 0u);
    uint32_t pw[KP];
#pragma unroll
    for (int k = 0; k < KP; k++) pw[k] = k < (int)to.y ? __ldg(a.woff + ((int64_t)to.x + k) * 32 + lane) : NONE;
    mbar_wait(bar, phase);
    phase ^= 1u;
    __syncthreads();
    T *oc = out + c * n + g0;
    for (int t = t0; t < t1; t++) {
      uint4 mnext = make_uint4(0u, 0u, 0xFFFFFFFFu, 0u);
      uint2 tonext = make_uint2(0u, 0u);
      if (t + 1 < t1) {
        mnext = __ldg(mp + (int64_t)(t + 1 - t0) * 32);
        tonext = __ldg(a.toff + t + 1);
      }
      const int nquad = (int)(m.w >> 16);
      const bool valid = m.z != 0xFFFFFFFFu;
      const int noff = (int)to.y;
      T px[KP];
      uint32_t pc[KP];  // the words of this task (their low bits decode the coefficient)
#pragma unroll
      for (int k = 0; k < KP; k++) {
        pc[k] = pw[k];
        px[k] = colblk_off_load<T, MODE>(pc[k], vc);
      }
#pragma unroll
      for (int k = 0; k < KP; k++) pw[k] = k < (int)tonext.y ? __ldg(a.woff + ((int64_t)tonext.x + k) * 32 + lane) : NONE;
      T acc;
      colres_zero(acc);
      if (dg.enabled && valid)
        acc = colres_scale(__hiloint2double((int)m.y, (int)m.x) + dtab[m.w & 0xFFFFu], xs[m.z]);
      if constexpr (MODE == 3) {
        // 16-bit words: one 4-byte load per two steps, four loads in flight (as in k_colres)
        for (int kq = 0; kq < nquad; kq++) {
          const uint32_t w = wa;
          wp += 32;
          wa = wb; wb = wc; wc = wd;
          wd = __ldg(wp + 96);
          colres_step16x2<T>(acc, w, xs_b, a.m0, a.m1);
        }
      } else {
        // two operator loads in flight in registers that are never moved (a move would wait for the load it copies)
        auto quad = [&](const uint4 w) {
          colres_step<T, MODE>(acc, w.x, xs_b, coef_b, a.m0, a.m1, a.m2, a.m3);
          colres_step<T, MODE>(acc, w.y, xs_b, coef_b, a.m0, a.m1, a.m2, a.m3);
          colres_step<T, MODE>(acc, w.z, xs_b, coef_b, a.m0, a.m1, a.m2, a.m3);
          colres_step<T, MODE>(acc, w.w, xs_b, coef_b, a.m0, a.m1, a.m2, a.m3);
        };
        int kq = 0;
        for (; kq + 2 <= nquad; kq += 2) {
          const uint4 w0 = wa;
          wa = __ldg(wp + 64);
          quad(w0);
          const uint4 w1 = wb;
          wb = __ldg(wp + 96);
          wp += 64;
          quad(w1);
        }
        if (kq < nquad) {
          const uint4 w0 = wa;
          wa = wb;
          wb = __ldg(wp + 64);
          wp += 32;
          quad(w0);
        }
      }
#pragma unroll
      for (int k = 0; k < KP; k++) colblk_off_apply<T, MODE>(acc, pc[k], px[k], coef_b, a.m0, a.m1, a.m2, a.m3);
      for (int k = KP; k < noff; k++) {  // more off-block steps than prefetch slots
        const uint32_t w0 = __ldg(a.woff + ((int64_t)to.x + k) * 32 + lane);
        colblk_off_apply<T, MODE>(acc, w0, colblk_off_load<T, MODE>(w0, vc), coef_b, a.m0, a.m1, a.m2, a.m3);
      }
      if (valid) oc[m.z] = acc;
      m = mnext;
      to = tonext;
    }
    __syncthreads();  // every gather of this block is done before the next bulk copy lands
  }
}

// shared memory the kernel needs for a column of n elements of elem bytes
inline size_t colres_smem(int64_t n, int elem, int nimp_diag) {
  const size_t ndt = nimp_diag >= 0 ? ((size_t)1 << nimp_diag) : 0;
  return 128 + 2048 + ((ndt * 8 + 127) & ~(size_t)127) + ((size_t)n + 32) * elem;  // + padding and zero elements
}

// launch on the context's stream; returns kColresNA when the kernel does not apply (the caller then uses the
// generic kernel), 0 on success, the usual non-zero rc on a CUDA error
constexpr int kColresNA = -7;
// does the whole-column kernel apply to (s, T, dg)?  (it is the only column pass that can accumulate)
template <typename T>
inline bool colres_applicable(const SpinOp &s, const DiagArgs &dg) {
  Ctx &c = ctx();
  const Sched &sc = sizeof(T) == 16 ? s.sc8 : s.sc16;
  if (c.opt.colpass_variant != 6 || c.opt.colres_rows > 0) return false;
  if (!use_tables() || !sc.words || sc.ntask <= 0) return false;
  if (sizeof(T) == 8 && ((s.n & 1) || !c.real_h)) return false;  // bulk copies need 16-byte aligned columns
  if (colres_smem(s.n, (int)sizeof(T), dg.enabled ? dg.nimp : -1) > 232448) return false;
  if (dg.enabled && dg.f_row != s.f) return false;  // the schedule carries the operator's own row diagonal
  return true;
}

// accum: out += instead of out = ; final: out is the finished H x v after this launch, so a pending Lanczos
// dot request (Ctx::dot_request) is served here
template <typename T>
inline int launch_colres(const SpinOp &s, int64_t ncols, const T *v, T *out, const DiagArgs &dg, bool accum = false,
                         bool final = false) {
  Ctx &c = ctx();
  const Sched &sc = sizeof(T) == 16 ? s.sc8 : s.sc16;
  if (!use_tables() || !sc.words || sc.ntask <= 0) return kColresNA;
  if (sizeof(T) == 8 && ((s.n & 1) || !c.real_h)) return kColresNA;  // bulk copies need 16-byte aligned columns
  const size_t smem = colres_smem(s.n, (int)sizeof(T), dg.enabled ? dg.nimp : -1);
  if (smem > 232448) return kColresNA;
  if (dg.enabled && dg.f_row != s.f) return kColresNA;  // the schedule carries the operator's own row diagonal
  ColResArgs a{};
  a.accum = accum ? 1 : 0;
  a.tbase = sc.tbase; a.qbase = sc.qbase; a.meta = (const uint4 *)sc.meta; a.words = sc.words;
  a.coef = s.coef; a.m0 = s.sc_mag[0]; a.m1 = s.sc_mag[1]; a.m2 = s.sc_mag[2]; a.m3 = s.sc_mag[3];
  void (*kern)(int64_t, int64_t, const T *, T *, ColResArgs, DiagArgs) =
      sc.fmt == 3 ? k_colres<T, 4>
                  : (sc.fmt == 2 ? k_colres<T, 3> : (sc.fmt == 1 ? k_colres<T, 2> : (c.real_h ? k_colres<T, 1> : k_colres<T, 0>)));
  const int threads = sc.nwarps * 32;  // the streams were dealt for exactly this many warps
  // attribute + occupancy query once per (kernel, shared-memory size)
  static std::map<const void *, size_t> max_smem;  // the attribute is only ever raised
  static std::map<std::pair<const void *, size_t>, int> configured;
  size_t &ms = max_smem[(const void *)kern];
  if (smem > ms) {
    CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ms = smem;
  }
  auto key = std::make_pair((const void *)kern, smem * 4096 + (size_t)threads);
  auto it = configured.find(key);
  if (it == configured.end()) {
    int q = 0;
    CB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, kern, threads, smem));
    it = configured.emplace(key, q).first;
  }
  const int per_sm = it->second;
  if (per_sm < 1) return kColresNA;
  const int64_t grid = std::min<int64_t>(ncols, (int64_t)c.sm_count * per_sm);
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
  kern<<<(unsigned)grid, threads, smem, c.stream>>>(s.n, ncols, v, out, a, dg);
  c.launches++;
  return 0;
}


template <typename T>
inline int launch_colblk(const SpinOp &s, int64_t ncols, const T *v, T *out, const DiagArgs &dg) {
  Ctx &c = ctx();
  const ColBlk &cb = sizeof(T) == 16 ? s.cb8 : s.cb16;
  if (!use_tables() || !cb.words || cb.nblk <= 0) return kColresNA;
  if (sizeof(T) == 8 && ((s.n & 1) || !c.real_h || !cb.even_blocks)) return kColresNA;  // 16-byte aligned bulk copies
  if (dg.enabled && dg.f_row != s.f) return kColresNA;
  const size_t smem = colres_smem(cb.max_rows, (int)sizeof(T), dg.enabled ? dg.nimp : -1);
  if (smem > 232448) return kColresNA;
  ColBlkArgs a{};
  a.blk = cb.blk; a.tbase = cb.tbase; a.qbase = cb.qbase; a.meta = (const uint4 *)cb.meta; a.words = cb.words;
  a.toff = cb.toff; a.woff = cb.woff; a.coef = s.coef;
  a.m0 = s.sc_mag[0]; a.m1 = s.sc_mag[1]; a.m2 = s.sc_mag[2]; a.m3 = s.sc_mag[3];
  a.nblk = cb.nblk; a.nwarps = cb.nwarps;
  void (*kern)(int64_t, int64_t, const T *, T *, ColBlkArgs, DiagArgs) =
      cb.fmt == 3 ? k_colblk<T, 4> : (cb.fmt == 2 ? k_colblk<T, 3> : (cb.fmt == 1 ? k_colblk<T, 2> : (c.real_h ? k_colblk<T, 1> : k_colblk<T, 0>)));
  static std::map<const void *, size_t> max_smem;
  size_t &ms = max_smem[(const void *)kern];
  if (smem > ms) {
    CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ms = smem;
  }
  if (cb.nwarps != kColblkWarps) return fail("internal: block-split schedule dealt for %d warps", cb.nwarps);
  const int threads = cb.nwarps * 32;
  const int64_t grid = std::min<int64_t>(ncols * cb.nblk, (int64_t)c.sm_count);
  kern<<<(unsigned)grid, threads, smem, c.stream>>>(s.n, ncols, v, out, a, dg);
  c.launches++;
  return 0;
}

}  // namespace cb
