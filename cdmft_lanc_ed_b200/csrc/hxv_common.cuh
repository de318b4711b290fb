// hxv_common.cuh -- device helpers shared by the H x v translation units (hxv.cu, hxv_real.cu)
#pragma once
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


}  // namespace cb
