// rowtile.cu -- tile-resident row pass:  out(i,c) (=|+=) sum_k Hd(c,j_k) v(i,j_k)
//
// The 1 (x) Hdw term of spMatVec_main (ED_HAMILTONIAN_SPARSE_HxV.f90:196-207, sparse/H_dw.f90:1-89) acts on
// the STRIDED index of v(iup + idw*DimUp).  The reference (and the generic kernel k_rowpass_rb) fetch every
// v element once per non-zero of Hdw that references it (~15 times at Ns=16); on B200 that pass is bound by
// the L2 -> SM fill bandwidth (64 B/clk/SM), not by HBM.  Here the re-use happens in shared memory:
//
//   * columns (dw states) are cut into BLOCKS = runs of states sharing their leading bits, split recursively
//     until a block fits the tile (build_rowtile); hops that leave those bits alone stay inside a block
//     (K3: 22 blocks of 462-792 columns, ~70 % of the entries are in-block);
//   * a work item is (strip of 8 consecutive rows) x (block): the tile v(i0..i0+7, block) -- one 128-byte
//     line per column -- is brought in by 2-D TMA tensor copies (cp.async.bulk.tensor.2d, 32 columns per
//     instruction, out-of-range rows zero-filled) into a DOUBLE-BUFFERED ring: a producer warp runs one item
//     ahead of 16 consumer warps (full/empty mbarriers, no CTA-wide barrier in the steady state);
//   * a warp task is 4 columns (one per 8-lane group, the 8 lanes = the 8 rows): in-block entries read the
//     tile -- each group one full 128-byte line, conflict-free by construction -- at 128 B/clk/SM; the entries
//     that change the leading bits read global memory / L2 (one line per group) and are PREFETCHED one task
//     ahead into registers, so their latency hides behind the previous task's shared-memory work;
//   * persistent CTAs, one per SM, walk the items slab by slab (16 strips x all blocks), so the off-block
//     lines of a slab are L2 hits and v is read from HBM once.
// Operator words are per-block streams built on the host at sector build (RowRes, ctx.h); the decode is
// sign/class/phase bits for purely real-or-imaginary coefficients of <= 4 magnitudes (every replica-bath
// Hubbard and BHZ model), a 128-entry coefficient table otherwise.
#include <cuda.h>

#include <algorithm>
#include <vector>

#include "ctx.h"
#include "hxv_common.cuh"

namespace cb {

static int64_t binom64r(int n, int k) {
  if (k < 0 || k > n) return 0;
  int64_t r = 1;
  for (int i = 1; i <= k; i++) r = r * (n - k + i) / i;
  return r;
}

// blocks: the states with `m` particles in the low `t` bits (all higher bits fixed) are contiguous in rank
// order; split on bit t-1 (bit clear first, then bit set) until the run fits `cap`.
static void split_blocks(int t, int m, int64_t start, int64_t cap, std::vector<int2> &out) {
  const int64_t sz = binom64r(t, m);
  if (sz <= 0) return;
  if (sz <= cap || t == 0) {
    out.push_back(make_int2((int)start, (int)sz));
    return;
  }
  split_blocks(t - 1, m, start, cap, out);
  split_blocks(t - 1, m - 1, start + binom64r(t - 1, m), cap, out);
}

// fmt 1 (fast4): code = negative | class << 1 | imaginary << 3 ; fmt 0: code = coefficient-table id (0 = 0.0)
//   in-block word   fmt 1: negative << 31 | (rel + 1) << 7 | imaginary << 2 | class      idle = 0 (the zero line)
//                   fmt 0: (rel + 1) << 7 | id
//   off-block word  fmt 1: negative << 31 | column << 3 | imaginary << 2 | class         none = 0xFFFFFFFF
//                   fmt 0: column << 7 | id                                                none = 0
int build_rowtile(SpinOp &op, int ns, const std::vector<int32_t> &rowptr, const std::vector<int32_t> &col,
                  const std::vector<uint8_t> &code, int fmt, int64_t cap) {
  Ctx &c = ctx();
  RowRes &rr = op.rr;
  std::vector<int2> blk;
  split_blocks(ns, op.npart, 0, cap, blk);
  int64_t covered = 0, mx = 0;
  for (auto &b : blk) {
    if (b.x != covered) return fail("internal: rowtile blocks are not contiguous");
    covered += b.y;
    mx = std::max<int64_t>(mx, b.y);
  }
  if (covered != op.n) return fail("internal: rowtile blocks do not cover the sector");
  std::vector<int32_t> tbase(blk.size() + 1, 0), task_col;
  std::vector<uint4> task;
  std::vector<uint32_t> win, woff;
  const bool fast = fmt == 1;
  const uint32_t NONE = fast ? 0xFFFFFFFFu : 0u;
  auto word_in = [&](int64_t rel, uint32_t cd) -> uint32_t {
    const uint32_t off = (uint32_t)(rel + 1) << 7;
    return fast ? ((cd & 1u) << 31) | off | ((cd >> 1) & 3u) | (((cd >> 3) & 1u) << 2) : off | cd;
  };
  auto word_off = [&](int64_t j, uint32_t cd) -> uint32_t {
    return fast ? ((cd & 1u) << 31) | ((uint32_t)j << 3) | ((cd >> 1) & 3u) | (((cd >> 3) & 1u) << 2) : ((uint32_t)j << 7) | cd;
  };
  int64_t nin_tot = 0;
  for (size_t b = 0; b < blk.size(); b++) {
    const int g0 = blk[b].x, ng = blk[b].y;
    tbase[b] = (int32_t)task.size();
    std::vector<int32_t> order(ng), nin(ng, 0);
    for (int k = 0; k < ng; k++) {
      order[k] = k;
      for (int32_t p = rowptr[g0 + k]; p < rowptr[g0 + k + 1]; p++) nin[k] += (col[p] >= g0 && col[p] < g0 + ng);
      nin_tot += nin[k];
    }
    // columns with the same amount of shared-memory work share a task (the 4 lane groups run in lockstep)
    std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) { return nin[x] > nin[y]; });
    for (int k0 = 0; k0 < ng; k0 += 4) {
      int cols[4], kin = 0, koff = 0;
      for (int q = 0; q < 4; q++) {
        cols[q] = k0 + q < ng ? order[k0 + q] : -1;
        if (cols[q] >= 0) {
          const int len = rowptr[g0 + cols[q] + 1] - rowptr[g0 + cols[q]];
          kin = std::max(kin, nin[cols[q]]);
          koff = std::max(koff, len - nin[cols[q]]);
        }
        task_col.push_back(cols[q]);
      }
      const int nq = (kin + 3) / 4;
      const size_t qb = win.size() / 16, ob = woff.size() / 4;
      task.push_back(make_uint4((uint32_t)qb, (uint32_t)nq, (uint32_t)ob, (uint32_t)koff));
      win.resize(win.size() + (size_t)nq * 16, 0u);  // idle steps read the zero line in front of the tile
      woff.resize(woff.size() + (size_t)koff * 4, NONE);
      for (int q = 0; q < 4; q++) {
        if (cols[q] < 0) continue;
        int ki = 0, ko = 0;
        for (int32_t p = rowptr[g0 + cols[q]]; p < rowptr[g0 + cols[q] + 1]; p++) {
          const int32_t j = col[p];
          if (j >= g0 && j < g0 + ng) {
            win[(qb + ki / 4) * 16 + q * 4 + (ki & 3)] = word_in(j - g0, code[p]);  // uint4 (4 steps) per quad and group
            ki++;
          } else {
            woff[(ob + ko) * 4 + q] = word_off(j, code[p]);
            ko++;
          }
        }
      }
    }
  }
  tbase[blk.size()] = (int32_t)task.size();
  rr.nblocks = (int32_t)blk.size();
  rr.max_block = (int32_t)mx;
  rr.ntask = (int32_t)task.size();
  rr.fmt = fmt;
  rr.in_frac = rowptr[op.n] > 0 ? (double)nin_tot / (double)rowptr[op.n] : 1.0;
  win.resize(win.size() + 64, 0u);  // slack for the one-quad-ahead prefetch
  woff.resize(woff.size() + 16, NONE);
  CB_CHECK(dev_alloc(&rr.blocks, (int64_t)blk.size()));
  CB_CHECK(dev_alloc(&rr.tbase, (int64_t)tbase.size()));
  CB_CHECK(dev_alloc(&rr.task, (int64_t)task.size()));
  CB_CHECK(dev_alloc(&rr.task_col, (int64_t)task_col.size()));
  CB_CHECK(dev_alloc(&rr.win, (int64_t)win.size()));
  CB_CHECK(dev_alloc(&rr.woff, (int64_t)woff.size()));
  CB_CUDA(cudaMemcpyAsync(rr.blocks, blk.data(), blk.size() * sizeof(int2), cudaMemcpyHostToDevice, c.stream));
  CB_CUDA(cudaMemcpyAsync(rr.tbase, tbase.data(), tbase.size() * 4, cudaMemcpyHostToDevice, c.stream));
  CB_CUDA(cudaMemcpyAsync(rr.task, task.data(), task.size() * sizeof(uint4), cudaMemcpyHostToDevice, c.stream));
  CB_CUDA(cudaMemcpyAsync(rr.task_col, task_col.data(), task_col.size() * 4, cudaMemcpyHostToDevice, c.stream));
  CB_CUDA(cudaMemcpyAsync(rr.win, win.data(), win.size() * 4, cudaMemcpyHostToDevice, c.stream));
  CB_CUDA(cudaMemcpyAsync(rr.woff, woff.data(), woff.size() * 4, cudaMemcpyHostToDevice, c.stream));
  CB_CUDA(cudaStreamSynchronize(c.stream));
  return 0;
}

// ------------------------------------------------------------------------------------
constexpr int RT_NCW = 16;   // consumer warps (+ 1 producer warp)
constexpr int RT_KP = 8;     // off-block entries prefetched per task and lane
constexpr int RT_BOX = 32;   // columns per TMA tensor copy (32 x 128 B = 4 KB)

struct RowTileArgs {
  const int2 *blocks;
  const int32_t *tbase;
  const uint4 *task;
  const int32_t *task_col;
  const uint4 *win;
  const uint32_t *woff;
  const double2 *coef;
  double m0, m1, m2, m3;
  int nblocks;
  int slab;      // strips per slab (slab x all blocks is the L2 working set)
  int bufcols;   // 128-byte lines per tile buffer: zero line + columns rounded up to RT_BOX
  int tma2d;     // 1 = 2-D tensor copies, 0 = one 128-byte bulk copy per column
  int64_t nstrips;
};

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *tm, int x, int y, uint64_t *bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(
                   smem_u32(dst)),
               "l"((uint64_t)tm), "r"(x), "r"(y), "r"(smem_u32(bar))
               : "memory");
}

// MODE: 0 = coefficient table, complex values; 1 = coefficient table, real values; 2 = sign/class/phase bits
template <int MODE>
__device__ __forceinline__ void rt_fma(double2 &acc, uint32_t w, double2 x, const char *coef_b, double m0, double m1, double m2,
                                       double m3) {
  if (MODE == 2) {
    const double ma = (w & 1u) ? m1 : m0, mb = (w & 1u) ? m3 : m2;
    if (w & 4u) x = make_double2(-x.y, x.x);  // times i
    rfma(acc, colres_signed((w & 2u) ? mb : ma, w & 0x80000000u), x);
  } else if (MODE == 1) {
    rfma(acc, *(const double *)(coef_b + ((w & 127u) << 4)), x);
  } else {
    cfma(acc, *(const double2 *)(coef_b + ((w & 127u) << 4)), x);
  }
}
template <int MODE>
__device__ __forceinline__ bool rt_on(uint32_t w) { return MODE == 2 ? (w != 0xFFFFFFFFu) : ((w & 127u) != 0u); }
template <int MODE>
__device__ __forceinline__ int64_t rt_col(uint32_t w) { return MODE == 2 ? (int64_t)((w & 0x7FFFFFFFu) >> 3) : (int64_t)(w >> 7); }
template <int MODE>
__device__ __forceinline__ uint32_t rt_tileoff(uint32_t w) { return MODE == 2 ? (w & 0x7FFFFF80u) : (w & 0xFFFFFF80u); }

__device__ __forceinline__ void rt_item(const RowTileArgs &a, int64_t item, int &b, int64_t &strip) {
  const int64_t per_slab = (int64_t)a.slab * a.nblocks;
  const int64_t nslabs = (a.nstrips + a.slab - 1) / a.slab;
  int64_t slab = item / per_slab;
  if (slab > nslabs - 1) slab = nslabs - 1;
  const int64_t r = item - slab * per_slab;
  const int64_t ns_in = min((int64_t)a.slab, a.nstrips - slab * a.slab);
  b = (int)(r / ns_in);
  strip = slab * a.slab + r % ns_in;
}

template <int MODE, bool ACCUM>
__global__ void __launch_bounds__((RT_NCW + 1) * 32, 1)
    k_rowtile(const __grid_constant__ CUtensorMap tm, int64_t n /*rows*/, const double2 *__restrict__ v, double2 *out, RowTileArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // layout: [0,16) full[2] | [16,32) empty[2] | [128, 128+2048) coefficient table | two tile buffers, each
  // bufcols lines of 128 B: line 0 stays zero (idle steps), column rel of the block is line rel+1
  uint64_t *full = (uint64_t *)smem_raw;
  uint64_t *empty = full + 2;
  double2 *coef = (double2 *)(smem_raw + 128);
  unsigned char *buf0 = smem_raw + 128 + 2048;
  const size_t bufbytes = (size_t)a.bufcols * 128;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    mbar_init(&empty[0], RT_NCW);
    mbar_init(&empty[1], RT_NCW);
  }
  if (threadIdx.x < 128) coef[threadIdx.x] = MODE == 2 ? make_double2(0.0, 0.0) : a.coef[threadIdx.x];
  if (threadIdx.x < 16) ((double2 *)(buf0 + (threadIdx.x >> 3) * bufbytes))[threadIdx.x & 7] = make_double2(0.0, 0.0);
  __syncthreads();
  const int64_t nitems = a.nstrips * a.nblocks;

  if (warp == RT_NCW) {
    // ===== producer: tile of item it -> buffer it & 1, one item ahead of the consumers =====
    int it = 0;
    for (int64_t item = blockIdx.x; item < nitems; item += gridDim.x, it++) {
      const int bf = it & 1;
      if (it >= 2) mbar_wait(&empty[bf], (uint32_t)(((it >> 1) - 1) & 1));  // every consumer warp released use it-2
      int b;
      int64_t strip;
      rt_item(a, item, b, strip);
      const int2 bd = __ldg(a.blocks + b);
      const int g0 = bd.x, ng = bd.y;
      const int64_t i0 = strip * 8;
      unsigned char *dst = buf0 + bf * bufbytes + 128;
      if (a.tma2d) {
        if (lane == 0) {
          const int nbox = (ng + RT_BOX - 1) / RT_BOX;
          mbar_expect_tx(&full[bf], (uint32_t)nbox * (uint32_t)(RT_BOX * 128));  // zero-filled parts count as well
          for (int k = 0; k < nbox; k++) tma_load_2d(dst + (size_t)k * (RT_BOX * 128), &tm, (int)(2 * i0), g0 + k * RT_BOX, &full[bf]);
        }
      } else {
        const int nb = (int)min((int64_t)8, n - i0);
        if (lane == 0) mbar_expect_tx(&full[bf], (uint32_t)ng * (uint32_t)nb * 16u);
        __syncwarp();
        for (int cidx = lane; cidx < ng; cidx += 32)
          bulk_g2s(dst + (size_t)cidx * 128, v + (int64_t)(g0 + cidx) * n + i0, (uint32_t)nb * 16u, &full[bf]);
      }
      __syncwarp();
    }
    return;
  }

  // ===== consumers =====
  const int r = lane & 7, grp = lane >> 3;
  const char *coef_b = (const char *)coef;
  int it = 0;
  for (int64_t item = blockIdx.x; item < nitems; item += gridDim.x, it++) {
    const int bf = it & 1;
    int b;
    int64_t strip;
    rt_item(a, item, b, strip);
    const int2 bd = __ldg(a.blocks + b);
    const int g0 = bd.x;
    const int64_t i0 = strip * 8;
    const int nb = (int)min((int64_t)8, n - i0);
    const int rc = min(r, nb - 1);  // ragged last strip: clamp the loads, skip the stores
    const double2 *vrow = v + i0 + rc;
    double2 *orow = out + i0 + rc;
    const char *tile_b = (const char *)(buf0 + bf * bufbytes) + r * 16;
    const int t0 = __ldg(a.tbase + b), t1 = __ldg(a.tbase + b + 1);
    // the warp <-> task assignment rotates with the item so that the odd task of a block moves around
    int t = t0 + (warp + it) % RT_NCW;
    uint4 tk = make_uint4(0u, 0u, 0u, 0u);
    int cl = -1;
    uint32_t pw[RT_KP];
    double2 px[RT_KP];
    double2 py = make_double2(0.0, 0.0);
    auto prefetch = [&](int tt) {
      tk = __ldg(a.task + tt);
      cl = __ldg(a.task_col + tt * 4 + grp);
      const uint32_t *wo = a.woff + (int64_t)tk.z * 4 + grp;
#pragma unroll
      for (int k = 0; k < RT_KP; k++) {
        uint32_t w = MODE == 2 ? 0xFFFFFFFFu : 0u;
        if (k < (int)tk.w) w = __ldg(wo + k * 4);
        pw[k] = w;
      }
#pragma unroll
      for (int k = 0; k < RT_KP; k++) {
        px[k] = make_double2(0.0, 0.0);
        if (rt_on<MODE>(pw[k])) px[k] = ldg2(vrow + rt_col<MODE>(pw[k]) * n);
      }
      if (ACCUM) {
        py = make_double2(0.0, 0.0);
        if (cl >= 0) py = orow[(int64_t)(g0 + cl) * n];
      }
    };
    if (t < t1) prefetch(t);
    mbar_wait(&full[bf], (uint32_t)((it >> 1) & 1));
    for (; t < t1; t += RT_NCW) {
      double2 acc = ACCUM ? py : make_double2(0.0, 0.0);
      // ---- sources outside the block: prefetched while the previous task ran
#pragma unroll
      for (int k = 0; k < RT_KP; k++)
        if (rt_on<MODE>(pw[k])) rt_fma<MODE>(acc, pw[k], px[k], coef_b, a.m0, a.m1, a.m2, a.m3);
      if ((int)tk.w > RT_KP) {  // rare: more off-block entries than prefetch slots
        const uint32_t *wo = a.woff + (int64_t)tk.z * 4 + grp;
        for (int k = RT_KP; k < (int)tk.w; k++) {
          const uint32_t w = __ldg(wo + k * 4);
          if (rt_on<MODE>(w)) rt_fma<MODE>(acc, w, ldg2(vrow + rt_col<MODE>(w) * n), coef_b, a.m0, a.m1, a.m2, a.m3);
        }
      }
      const uint4 *wi = a.win + (int64_t)tk.x * 4 + grp;
      const int nq = (int)tk.y;
      const int cl_cur = cl;
      uint4 wn = __ldg(wi);  // slack behind the stream: always readable
      if (t + RT_NCW < t1) prefetch(t + RT_NCW);
      // ---- sources inside the block: shared memory, four steps per operator load
      for (int q = 0; q < nq; q++) {
        const uint4 w = wn;
        wn = __ldg(wi + (q + 1) * 4);
        rt_fma<MODE>(acc, w.x, *(const double2 *)(tile_b + rt_tileoff<MODE>(w.x)), coef_b, a.m0, a.m1, a.m2, a.m3);
        rt_fma<MODE>(acc, w.y, *(const double2 *)(tile_b + rt_tileoff<MODE>(w.y)), coef_b, a.m0, a.m1, a.m2, a.m3);
        rt_fma<MODE>(acc, w.z, *(const double2 *)(tile_b + rt_tileoff<MODE>(w.z)), coef_b, a.m0, a.m1, a.m2, a.m3);
        rt_fma<MODE>(acc, w.w, *(const double2 *)(tile_b + rt_tileoff<MODE>(w.w)), coef_b, a.m0, a.m1, a.m2, a.m3);
      }
      if (cl_cur >= 0 && r < nb) orow[(int64_t)(g0 + cl_cur) * n] = acc;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[bf]);
  }
}

// ------------------------------------------------------------------------------------
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                        const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_tmapEncodeTiled tmap_encoder() {
  static PFN_tmapEncodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (PFN_tmapEncodeTiled)p;
    else
      cudaGetLastError();
  }
  return fn;
}

size_t rowtile_smem(int64_t max_block) {
  const int64_t bufcols = 1 + (max_block + RT_BOX - 1) / RT_BOX * RT_BOX;
  return 128 + 2048 + 2 * (size_t)bufcols * 128;
}
// largest block (columns) two tile buffers leave room for
int64_t rowtile_cap() { return ((232448 - 128 - 2048) / 2 / 128 - 1) / RT_BOX * RT_BOX; }

bool rowtile_applicable(const SpinOp &s) {
  Ctx &c = ctx();
  return c.mode == CDMFT_B200_SPARSE && c.opt.rowpass_variant == 4 && s.rr.win && s.rr.ntask > 0 &&
         rowtile_smem(s.rr.max_block) <= 232448;
}

// out(i,c) (=|+=) sum_k Hd(c,j_k) v(i,j_k) for nrows rows of 16-byte elements; kColresNA when the kernel does not apply
int launch_rowtile(const SpinOp &s, int64_t nrows, const double2 *v, double2 *out, bool accum) {
  Ctx &c = ctx();
  if (!rowtile_applicable(s)) return kColresNA;
  const RowRes &rr = s.rr;
  if (nrows <= 0 || s.n <= 0) return 0;
  RowTileArgs a{};
  a.blocks = rr.blocks; a.tbase = rr.tbase; a.task = rr.task; a.task_col = rr.task_col;
  a.win = (const uint4 *)rr.win; a.woff = rr.woff; a.coef = s.coef;
  a.m0 = s.sc_mag[0]; a.m1 = s.sc_mag[1]; a.m2 = s.sc_mag[2]; a.m3 = s.sc_mag[3];
  a.nblocks = rr.nblocks;
  a.nstrips = (nrows + 7) / 8;
  a.slab = (int)std::max<int64_t>(1, std::min<int64_t>(a.nstrips, c.opt.row_slab / 8));
  a.bufcols = (int)(1 + ((int64_t)rr.max_block + RT_BOX - 1) / RT_BOX * RT_BOX);
  a.tma2d = c.opt.tma2d ? 1 : 0;
  CUtensorMap tm;
  memset(&tm, 0, sizeof tm);
  if (a.tma2d) {
    PFN_tmapEncodeTiled enc = tmap_encoder();
    if (!enc) a.tma2d = 0;
    else {
      // v as a 2-D tensor of doubles: dim0 = 2*nrows (contiguous), dim1 = columns, stride nrows*16 bytes
      const cuuint64_t gdim[2] = {(cuuint64_t)(2 * nrows), (cuuint64_t)s.n};
      const cuuint64_t gstr[1] = {(cuuint64_t)nrows * 16};
      const cuuint32_t box[2] = {16, (cuuint32_t)RT_BOX};
      const cuuint32_t estr[2] = {1, 1};
      const CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void *)v, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) a.tma2d = 0;  // e.g. a vector that is not 16-byte aligned: per-column bulk copies
    }
  }
  const size_t smem = rowtile_smem(rr.max_block);
  const int mode = rr.fmt == 1 ? 2 : (c.real_h ? 1 : 0);
  void (*kern)(const CUtensorMap, int64_t, const double2 *, double2 *, RowTileArgs) = nullptr;
  switch (mode * 2 + (accum ? 1 : 0)) {
    case 0: kern = k_rowtile<0, false>; break;
    case 1: kern = k_rowtile<0, true>; break;
    case 2: kern = k_rowtile<1, false>; break;
    case 3: kern = k_rowtile<1, true>; break;
    case 4: kern = k_rowtile<2, false>; break;
    default: kern = k_rowtile<2, true>; break;
  }
  static std::map<const void *, size_t> max_smem;
  size_t &ms = max_smem[(const void *)kern];
  if (smem > ms) {
    CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ms = smem;
  }
  const int64_t nitems = a.nstrips * a.nblocks;
  const int64_t grid = std::min<int64_t>(nitems, c.sm_count);
  kern<<<(unsigned)grid, (RT_NCW + 1) * 32, smem, c.stream>>>(tm, nrows, v, out, a);
  c.launches++;
  return 0;
}

}  // namespace cb
