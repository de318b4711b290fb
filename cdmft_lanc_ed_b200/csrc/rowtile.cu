// rowtile.cu -- tile-resident row pass:  out(i,c) (=|+=) sum_k Hd(c,j_k) v(i,j_k)
//
// The 1 (x) Hdw term of spMatVec_main (ED_HAMILTONIAN_SPARSE_HxV.f90:196-207, sparse/H_dw.f90:1-89) acts on
// the STRIDED index of v(iup + idw*DimUp).  The reference (and the generic kernel k_rowpass_rb) fetch every
// v element once per non-zero of Hdw that references it (~15 times at Ns=16); on B200 that pass is bound by
// the L2 -> SM fill bandwidth (64 B/clk/SM), not by HBM.  Here the re-use happens in shared memory:
//
//   * columns (dw states) are cut into BLOCKS = runs of states sharing their leading bits, split recursively
//     until a block fits the tile (build_rowtile); hops that leave those bits alone stay inside a block
//     (K3: 22 blocks of 462-792 columns, 67 % of the entries are in-block);
//   * a work item is (strip of 8 consecutive rows) x (block): the tile v(i0..i0+7, block) -- one 128-byte
//     line per column -- is brought in by 2-D TMA tensor copies (cp.async.bulk.tensor.2d, 32 columns per
//     instruction, out-of-range rows zero-filled).  Two 320-thread CTAs per SM, each with ONE tile: while one
//     waits for its tile the other computes, no producer / consumer hand-shake;
//   * a warp task is 4 columns (one per 8-lane group, the 8 lanes = the 8 rows): in-block entries read the
//     tile -- each group one full 128-byte line, conflict-free by construction -- at 128 B/clk/SM; the entries
//     that change the leading bits read global memory / L2 (one line per group) and are PREFETCHED one task
//     ahead into registers: their operator words are requested at the top of a task, the gathers are issued
//     after the task's first shared-memory quad (by then the words are there), so no load waits on a load;
//   * every (block, warp) owns one contiguous stream of task headers / in-block quads / off-block rows, built on
//     the host at sector build and read from L2 with a fixed look-ahead (RowRes, ctx.h);
//   * items are ordered blocks-fastest, so the CTAs that run at the same time cover all blocks of a few strips:
//     the off-block lines are L2 hits and v is read from HBM once.
// The decode is sign/class/phase bits for purely real-or-imaginary coefficients of <= 4 magnitudes (every
// replica-bath Hubbard and BHZ model), a 128-entry coefficient table otherwise.
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

// fmt 1, 2 (sign/class[/phase] bits): code = negative | class << 1 | imaginary << 3 ; fmt 0: code = coefficient-table id (0 = 0.0)
//   in-block word   fmt 1: negative << 31 | (rel + 1) << 7 | imaginary << 2 | class      idle = 0 (the zero line)
//                   fmt 0: (rel + 1) << 7 | id
//   off-block word  fmt 1: negative << 31 | column << 3 | imaginary << 2 | class         none = 0xFFFFFFFF
//                   fmt 0: column << 7 | id                                                none = 0
#ifndef RT_NW_
#define RT_NW_ 10
#endif
#ifndef RT_KP_
#define RT_KP_ 6
#endif
constexpr int RT_NW = RT_NW_;  // warps per CTA (two CTAs per SM)
constexpr int RT_KP = RT_KP_;  // off-block entries prefetched per task and lane
constexpr int RT_BOX = 32;   // columns per TMA tensor copy (32 x 128 B = 4 KB)

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
  if (mx >= 1023) return fail("internal: rowtile block too large for 10-bit line ids");
  const bool fast = fmt >= 1;  // 1 = sign/class bits (real, two magnitudes), 2 = sign/class/phase bits
  const uint32_t NONE = fast ? 0xFFFFFFFFu : 0u;
  auto word_in = [&](int64_t rel, uint32_t cd) -> uint32_t {
    const uint32_t off = (uint32_t)(rel + 1) << 7;
    return fast ? ((cd & 1u) << 31) | off | ((cd >> 1) & 3u) | (((cd >> 3) & 1u) << 2) : off | cd;
  };
  auto word_off = [&](int64_t j, uint32_t cd) -> uint32_t {
    return fast ? ((cd & 1u) << 31) | ((uint32_t)j << 3) | ((cd >> 1) & 3u) | (((cd >> 3) & 1u) << 2) : ((uint32_t)j << 7) | cd;
  };
  // in-block words: 16 bits (negative << 15 | imaginary << 12 | class << 10 | line) for the bit-decoded formats, one
  // uint4 = 8 steps per lane group; 32 bits for the coefficient-table format, one uint4 = 4 steps
  const int steps_per_unit = fast ? 8 : 4;
  auto word_in16 = [&](int64_t rel, uint32_t cd) -> uint32_t {
    return ((cd & 1u) << 15) | (((cd >> 3) & 1u) << 12) | (((cd >> 1) & 3u) << 10) | (uint32_t)(rel + 1);
  };
  std::vector<int4> wbase(blk.size() * RT_NW);
  std::vector<uint32_t> thdr;
  std::vector<uint4> win, woff;
  int64_t nin_tot = 0;
  struct Task { int cols[4]; int nq, koff; };
  for (size_t b = 0; b < blk.size(); b++) {
    const int g0 = blk[b].x, ng = blk[b].y;
    std::vector<int32_t> order(ng), nin(ng, 0), nof(ng, 0);
    for (int k = 0; k < ng; k++) {
      order[k] = k;
      for (int32_t p = rowptr[g0 + k]; p < rowptr[g0 + k + 1]; p++) nin[k] += (col[p] >= g0 && col[p] < g0 + ng);
      nof[k] = rowptr[g0 + k + 1] - rowptr[g0 + k] - nin[k];
      nin_tot += nin[k];
    }
    // columns with the same amount of shared-memory and L2 work share a task (the 4 lane groups run in lockstep)
    std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) {
      const int qx = (nin[x] + 3) / 4, qy = (nin[y] + 3) / 4;
      if (qx != qy) return qx > qy;
      if (nof[x] != nof[y]) return nof[x] > nof[y];
      return nin[x] > nin[y];
    });
    std::vector<Task> tasks;
    for (int k0 = 0; k0 < ng; k0 += 4) {
      Task t{};
      int kin = 0;
      for (int q = 0; q < 4; q++) {
        t.cols[q] = k0 + q < ng ? order[k0 + q] : -1;
        if (t.cols[q] >= 0) {
          kin = std::max(kin, nin[t.cols[q]]);
          t.koff = std::max(t.koff, nof[t.cols[q]]);
        }
      }
      t.nq = (kin + 3) / 4;
      if (t.nq > 255 || t.koff > 255) return fail("internal: rowtile task too long");
      tasks.push_back(t);
    }
    // tasks are in descending cost order: dealing them round-robin balances the warps
    for (int w = 0; w < RT_NW; w++) {
      int4 wb = make_int4((int)(thdr.size() / 4), 0, (int)(win.size() / 4), (int)woff.size());
      for (size_t ti = w; ti < tasks.size(); ti += RT_NW) {
        const Task &t = tasks[ti];
        for (int q = 0; q < 4; q++)  // per lane group: column | quads << 16 | off-block steps << 24
          thdr.push_back((t.cols[q] >= 0 ? (uint32_t)t.cols[q] : 0xFFFFu) | ((uint32_t)t.nq << 16) | ((uint32_t)t.koff << 24));
        const int nu = std::max(2, ((t.nq * 4 + steps_per_unit - 1) / steps_per_unit + 1) / 2 * 2);  // pairs of units
        const size_t qb = win.size(), ob = woff.size();
        win.resize(qb + (size_t)nu * 4, make_uint4(0u, 0u, 0u, 0u));  // idle steps read the zero line in front of the tile
        woff.resize(ob + (size_t)t.koff, make_uint4(NONE, NONE, NONE, NONE));
        for (int q = 0; q < 4; q++) {
          if (t.cols[q] < 0) continue;
          int ki = 0, ko = 0;
          for (int32_t p = rowptr[g0 + t.cols[q]]; p < rowptr[g0 + t.cols[q] + 1]; p++) {
            const int32_t j = col[p];
            if (j >= g0 && j < g0 + ng) {
              if (fast) {
                uint16_t *u = (uint16_t *)&win[qb + (size_t)(ki / 8) * 4 + q];
                u[ki & 7] = (uint16_t)word_in16(j - g0, code[p]);
              } else {
                uint32_t *u = (uint32_t *)&win[qb + (size_t)(ki / 4) * 4 + q];
                u[ki & 3] = word_in(j - g0, code[p]);
              }
              ki++;
            } else {
              ((uint32_t *)&woff[ob + ko])[q] = word_off(j, code[p]);  // uint4 (4 groups) per step
              ko++;
            }
          }
        }
        wb.y++;
      }
      wbase[b * RT_NW + w] = wb;
    }
  }
  rr.nblocks = (int32_t)blk.size();
  rr.max_block = (int32_t)mx;
  rr.ntask = (int32_t)(thdr.size() / 4);
  rr.nwarps = RT_NW;
  rr.fmt = fmt;
  rr.in_frac = rowptr[op.n] > 0 ? (double)nin_tot / (double)rowptr[op.n] : 1.0;
  // slack behind the streams for the fixed look-ahead of the kernel
  thdr.resize(thdr.size() + 16, 0xFFFFu);
  win.resize(win.size() + 16, make_uint4(0u, 0u, 0u, 0u));
  woff.resize(woff.size() + RT_KP + 1, make_uint4(NONE, NONE, NONE, NONE));
  CB_CHECK(dev_alloc(&rr.blocks, (int64_t)blk.size()));
  CB_CHECK(dev_alloc(&rr.wbase, (int64_t)wbase.size()));
  CB_CHECK(dev_alloc(&rr.thdr, (int64_t)thdr.size()));
  CB_CHECK(dev_alloc(&rr.win, (int64_t)win.size()));
  CB_CHECK(dev_alloc(&rr.woff, (int64_t)woff.size()));
  CB_CUDA(cudaMemcpyAsync(rr.blocks, blk.data(), blk.size() * sizeof(int2), cudaMemcpyHostToDevice, c.stream));
  CB_CUDA(cudaMemcpyAsync(rr.wbase, wbase.data(), wbase.size() * sizeof(int4), cudaMemcpyHostToDevice, c.stream));
  CB_CUDA(cudaMemcpyAsync(rr.thdr, thdr.data(), thdr.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, c.stream));
  CB_CUDA(cudaMemcpyAsync(rr.win, win.data(), win.size() * sizeof(uint4), cudaMemcpyHostToDevice, c.stream));
  CB_CUDA(cudaMemcpyAsync(rr.woff, woff.data(), woff.size() * sizeof(uint4), cudaMemcpyHostToDevice, c.stream));
  CB_CUDA(cudaStreamSynchronize(c.stream));
  return 0;
}

// ------------------------------------------------------------------------------------
struct RowTileArgs {
  const int2 *blocks;
  const int4 *wbase;
  const uint32_t *thdr;
  const uint4 *win;
  const uint4 *woff;
  const double2 *coef;
  double m0, m1, m2, m3;
  int nblocks;
  int bufcols;   // 128-byte lines of the tile buffer: zero line + columns rounded up to RT_BOX
  int tma2d;     // 1 = 2-D tensor copies, 0 = one 128-byte bulk copy per column
  int64_t nitems;
  unsigned int *queue;  // [2] {next item, CTAs done}: items are handed out dynamically so that the running CTAs stay
                        // within a few strips of each other (the off-block lines must still be in L2)
};

__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *tm, int x, int y, uint64_t *bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(
                   smem_u32(dst)),
               "l"((uint64_t)tm), "r"(x), "r"(y), "r"(smem_u32(bar))
               : "memory");
}

// MODE: 0 = coefficient table, complex values; 1 = coefficient table, real values; 2 = sign/class bits (real H,
// two magnitudes); 3 = sign/class/phase bits (purely real or purely imaginary coefficients, four magnitudes)
template <int MODE>
__device__ __forceinline__ void rt_fma(double2 &acc, uint32_t w, double2 x, const char *coef_b, double m0, double m1, double m2,
                                       double m3) {
  if (MODE == 2) {
    rfma(acc, colres_signed((w & 1u) ? m1 : m0, w & 0x80000000u), x);
  } else if (MODE == 3) {
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
__device__ __forceinline__ bool rt_on(uint32_t w) { return MODE >= 2 ? (w != 0xFFFFFFFFu) : ((w & 127u) != 0u); }
template <int MODE>
__device__ __forceinline__ int64_t rt_col(uint32_t w) { return MODE >= 2 ? (int64_t)((w & 0x7FFFFFFFu) >> 3) : (int64_t)(w >> 7); }
template <int MODE>
__device__ __forceinline__ uint32_t rt_tileoff(uint32_t w) { return MODE >= 2 ? (w & 0x7FFFFF80u) : (w & 0xFFFFFF80u); }

// one step from the low (HALF = 0) or high half of a register holding two 16-bit in-block words
// (negative << 15 | imaginary << 12 | class << 10 | line)
template <int MODE, int HALF>
__device__ __forceinline__ void rt_step16(double2 &acc, uint32_t w, const char *tile_l, double m0, double m1, double m2, double m3) {
  const uint32_t off = HALF ? ((w >> 9) & 0x1FF80u) : ((w << 7) & 0x1FF80u);
  double2 x = *(const double2 *)(tile_l + off);
  const uint32_t sg = HALF ? (w & 0x80000000u) : ((w << 16) & 0x80000000u);
  const uint32_t c0 = HALF ? 0x4000000u : 0x400u;
  if (MODE == 2) {
    rfma(acc, colres_signed((w & c0) ? m1 : m0, sg), x);
  } else {
    const double ma = (w & c0) ? m1 : m0, mb = (w & c0) ? m3 : m2;
    if (w & (c0 << 2)) x = make_double2(-x.y, x.x);  // times i
    rfma(acc, colres_signed((w & (c0 << 1)) ? mb : ma, sg), x);
  }
}

template <int MODE, bool ACCUM>
__global__ void __launch_bounds__(RT_NW * 32, 2)
    k_rowtile(const __grid_constant__ CUtensorMap tm, int64_t n /*rows*/, const double2 *__restrict__ v, double2 *out, RowTileArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // layout: [0,8) mbarrier | [128, 128+2048) coefficient table | tile: bufcols lines of 128 B, line 0 stays zero
  // (idle steps), column rel of the block is line rel+1
  uint64_t *full = (uint64_t *)smem_raw;
  volatile int64_t *s_next = (volatile int64_t *)(smem_raw + 16);  // [2]
  double2 *coef = (double2 *)(smem_raw + 128);
  unsigned char *tile = smem_raw + 128 + 2048;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = lane & 7, grp = lane >> 3;
  if (threadIdx.x == 0) mbar_init(full, 1);
  if (threadIdx.x < 128) coef[threadIdx.x] = MODE >= 2 ? make_double2(0.0, 0.0) : a.coef[threadIdx.x];
  if (threadIdx.x < 8) ((double2 *)tile)[threadIdx.x] = make_double2(0.0, 0.0);
  __syncthreads();
  const char *coef_b = (const char *)coef;
  const char *tile_l = (const char *)tile + r * 16;
  constexpr uint32_t NONE = MODE >= 2 ? 0xFFFFFFFFu : 0u;
  uint32_t phase = 0;
  int it = 0;
  for (int64_t item = blockIdx.x; item < a.nitems; it++) {
    if (threadIdx.x == 0) s_next[it & 1] = (int64_t)atomicAdd(a.queue, 1u) + gridDim.x;
    const int64_t strip = item / a.nblocks;
    const int b = (int)(item - strip * a.nblocks);
    const int2 bd = __ldg(a.blocks + b);
    const int g0 = bd.x, ng = bd.y;
    const int64_t i0 = strip * 8;
    const int nb = (int)min((int64_t)8, n - i0);
    // the tile is free: every warp passed the barrier at the end of the previous item
    if (a.tma2d) {
      if (threadIdx.x == 0) {
        const int nbox = (ng + RT_BOX - 1) / RT_BOX;
        mbar_expect_tx(full, (uint32_t)nbox * (uint32_t)(RT_BOX * 128));  // zero-filled parts count as well
        for (int k = 0; k < nbox; k++) tma_load_2d(tile + 128 + (size_t)k * (RT_BOX * 128), &tm, (int)(2 * i0), g0 + k * RT_BOX, full);
      }
    } else if (warp == 0) {
      if (lane == 0) mbar_expect_tx(full, (uint32_t)ng * (uint32_t)nb * 16u);
      __syncwarp();
      for (int cidx = lane; cidx < ng; cidx += 32)
        bulk_g2s(tile + 128 + (size_t)cidx * 128, v + (int64_t)(g0 + cidx) * n + i0, (uint32_t)nb * 16u, full);
    }
    const bool rvalid = r < nb;   // ragged last strip: clamp the loads, skip the stores
    const double2 *vrow = v + i0 + (rvalid ? r : nb - 1);
    double2 *orow = out + i0 + (rvalid ? r : nb - 1) + (int64_t)g0 * n;
    // this warp's streams for the block
    const int4 wb = __ldg(a.wbase + b * RT_NW + warp);
    // 32-bit stream positions against the (uniform) array bases: fewer live registers than three pointers
    const uint32_t *__restrict__ th = a.thdr;
    const uint4 *__restrict__ wi = a.win;
    const uint32_t *__restrict__ wo = (const uint32_t *)a.woff;
    uint32_t ith = (uint32_t)wb.x * 4u + grp;   // task k: th[ith + k * 4]
    const int ntask = wb.y;
    uint32_t iwi = (uint32_t)wb.z * 4u + grp;   // unit u: wi[iwi + u * 4]
    uint32_t iwo = (uint32_t)wb.w * 4u + grp;   // off-block step k: wo[iwo + k * 4]
    uint32_t H0 = __ldg(th + ith), H1 = __ldg(th + ith + 4);          // slack behind the array
    uint4 wa = __ldg(wi + iwi), wq = __ldg(wi + iwi + 4);
    uint32_t pw[RT_KP];
    double2 px[RT_KP];
    double2 py = make_double2(0.0, 0.0);
    auto load_words = [&](uint32_t H) {
      const int noff = (int)(H >> 24);
#pragma unroll
      for (int k = 0; k < RT_KP; k++) {
        uint32_t w = NONE;
        if (k < noff) w = __ldg(wo + iwo + k * 4);
        pw[k] = w;
      }
    };
    auto gather = [&](uint32_t H) {
#pragma unroll
      for (int k = 0; k < RT_KP; k++) {
        px[k] = make_double2(0.0, 0.0);
        if (rt_on<MODE>(pw[k])) px[k] = ldg2(vrow + rt_col<MODE>(pw[k]) * n);
      }
      if (ACCUM) {
        const uint32_t cn = H & 0xFFFFu;
        py = make_double2(0.0, 0.0);
        if (cn != 0xFFFFu) py = orow[(int64_t)cn * n];
      }
    };
    if (ntask > 0) {
      load_words(H0);
      gather(H0);
    }
    mbar_wait(full, phase);
    phase ^= 1u;
    for (int t = 0; t < ntask; t++) {
      const uint32_t H = H0;
      const int nq = (int)((H >> 16) & 0xFFu), noff = (int)(H >> 24);
      const uint32_t cl = H & 0xFFFFu;
      double2 acc = ACCUM ? py : make_double2(0.0, 0.0);
      // ---- sources outside the block: prefetched while the previous task ran
#pragma unroll
      for (int k = 0; k < RT_KP; k++)
        rt_fma<MODE>(acc, pw[k], px[k], coef_b, a.m0, a.m1, a.m2, a.m3);  // unused slot: x = 0
      for (int k = RT_KP; k < noff; k++) {  // rare: more off-block entries than prefetch slots
        const uint32_t w = __ldg(wo + iwo + k * 4);
        if (rt_on<MODE>(w)) rt_fma<MODE>(acc, w, ldg2(vrow + rt_col<MODE>(w) * n), coef_b, a.m0, a.m1, a.m2, a.m3);
      }
      iwo += (uint32_t)noff * 4u;
      H0 = H1;
      ith += 4u;
      H1 = __ldg(th + ith + 4);
      const bool more = t + 1 < ntask;
      if (more) load_words(H0);  // operator words of the next task: requested now, used after the first unit
      // ---- sources inside the block: shared memory.  One 16-byte operator load per lane and unit (8 steps of
      // 16-bit words, or 4 steps of 32-bit words), two units in flight in registers that are never moved
      constexpr int QU = MODE >= 2 ? 2 : 1;  // quads per unit
      const int nu = (nq + QU - 1) / QU;
      auto unit = [&](const uint4 w, bool second) {
        if (MODE >= 2) {
          rt_step16<MODE, 0>(acc, w.x, tile_l, a.m0, a.m1, a.m2, a.m3);
          rt_step16<MODE, 1>(acc, w.x, tile_l, a.m0, a.m1, a.m2, a.m3);
          rt_step16<MODE, 0>(acc, w.y, tile_l, a.m0, a.m1, a.m2, a.m3);
          rt_step16<MODE, 1>(acc, w.y, tile_l, a.m0, a.m1, a.m2, a.m3);
          if (second) {
            rt_step16<MODE, 0>(acc, w.z, tile_l, a.m0, a.m1, a.m2, a.m3);
            rt_step16<MODE, 1>(acc, w.z, tile_l, a.m0, a.m1, a.m2, a.m3);
            rt_step16<MODE, 0>(acc, w.w, tile_l, a.m0, a.m1, a.m2, a.m3);
            rt_step16<MODE, 1>(acc, w.w, tile_l, a.m0, a.m1, a.m2, a.m3);
          }
        } else {
          rt_fma<MODE>(acc, w.x, *(const double2 *)(tile_l + rt_tileoff<MODE>(w.x)), coef_b, a.m0, a.m1, a.m2, a.m3);
          rt_fma<MODE>(acc, w.y, *(const double2 *)(tile_l + rt_tileoff<MODE>(w.y)), coef_b, a.m0, a.m1, a.m2, a.m3);
          rt_fma<MODE>(acc, w.z, *(const double2 *)(tile_l + rt_tileoff<MODE>(w.z)), coef_b, a.m0, a.m1, a.m2, a.m3);
          rt_fma<MODE>(acc, w.w, *(const double2 *)(tile_l + rt_tileoff<MODE>(w.w)), coef_b, a.m0, a.m1, a.m2, a.m3);
        }
      };
      // units are loaded in pairs (the stream pads every task to an even number of units), so the two register
      // sets alternate without moves and every load has a full pair (16 or 8 steps) to arrive
      (void)nu;
      int qleft = nq;  // quads still to run
      bool first = true;
      while (qleft > 0 || first) {
        const uint4 w0 = wa;
        wa = __ldg(wi + iwi + 8);
        if (qleft > 0) unit(w0, qleft >= QU);
        qleft -= QU;
        if (first) {  // the next task's L2 gathers fly behind the rest of this task
          if (more) gather(H0);
          first = false;
        }
        const uint4 w1 = wq;
        wq = __ldg(wi + iwi + 12);
        iwi += 8u;
        if (qleft > 0) unit(w1, qleft >= QU);
        qleft -= QU;
      }
      if (cl != 0xFFFFu && rvalid) orow[(int64_t)cl * n] = acc;
    }
    __syncthreads();  // every read of the tile is done before the next copy lands
    item = s_next[it & 1];
  }
  // the last CTA to leave re-arms the queue for the next launch
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(a.queue + 1, 1u) == gridDim.x - 1) {
      a.queue[0] = 0u;
      a.queue[1] = 0u;
    }
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

// two CTAs per SM: each gets half of the 228 KB minus the 1 KB the system reserves per CTA
constexpr size_t kRowTileSmemMax = (233472 - 2 * 1024) / 2;
size_t rowtile_smem(int64_t max_block) {
  const int64_t bufcols = 1 + (max_block + RT_BOX - 1) / RT_BOX * RT_BOX;
  return 128 + 2048 + (size_t)bufcols * 128;
}
// largest block (columns) a tile buffer has room for
int64_t rowtile_cap() { return ((int64_t)(kRowTileSmemMax - 128 - 2048) / 128 - 1) / RT_BOX * RT_BOX; }

bool rowtile_applicable(const SpinOp &s) {
  Ctx &c = ctx();
  return use_tables() && c.opt.rowpass_variant == 4 && s.rr.win && s.rr.ntask > 0 &&
         rowtile_smem(s.rr.max_block) <= kRowTileSmemMax;
}

// out(i,c) (=|+=) sum_k Hd(c,j_k) v(i,j_k) for nrows rows of 16-byte elements; kColresNA when the kernel does not apply
int launch_rowtile(const SpinOp &s, int64_t nrows, const double2 *v, double2 *out, bool accum) {
  Ctx &c = ctx();
  if (!rowtile_applicable(s)) return kColresNA;
  const RowRes &rr = s.rr;
  if (nrows <= 0 || s.n <= 0) return 0;
  RowTileArgs a{};
  a.blocks = rr.blocks; a.wbase = rr.wbase; a.thdr = rr.thdr; a.win = rr.win; a.woff = rr.woff; a.coef = s.coef;
  a.m0 = s.sc_mag[0]; a.m1 = s.sc_mag[1]; a.m2 = s.sc_mag[2]; a.m3 = s.sc_mag[3];
  a.nblocks = rr.nblocks;
  a.nitems = (nrows + 7) / 8 * rr.nblocks;
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
  const int mode = rr.fmt == 1 ? 2 : (rr.fmt == 2 ? 3 : (c.real_h ? 1 : 0));
  void (*kern)(const CUtensorMap, int64_t, const double2 *, double2 *, RowTileArgs) = nullptr;
  switch (mode * 2 + (accum ? 1 : 0)) {
    case 0: kern = k_rowtile<0, false>; break;
    case 1: kern = k_rowtile<0, true>; break;
    case 2: kern = k_rowtile<1, false>; break;
    case 3: kern = k_rowtile<1, true>; break;
    case 4: kern = k_rowtile<2, false>; break;
    case 5: kern = k_rowtile<2, true>; break;
    case 6: kern = k_rowtile<3, false>; break;
    default: kern = k_rowtile<3, true>; break;
  }
  if (!c.rt_queue) {
    CB_CHECK(dev_alloc(&c.rt_queue, 2));
    CB_CUDA(cudaMemsetAsync(c.rt_queue, 0, 2 * sizeof(unsigned int), c.stream));
  }
  a.queue = c.rt_queue;
  static std::map<const void *, size_t> max_smem;
  static std::map<std::pair<const void *, size_t>, int> occ;
  size_t &ms = max_smem[(const void *)kern];
  if (smem > ms) {
    CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ms = smem;
  }
  auto key = std::make_pair((const void *)kern, smem);
  auto it = occ.find(key);
  if (it == occ.end()) {
    int q = 0;
    CB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, kern, RT_NW * 32, smem));
    it = occ.emplace(key, q).first;
  }
  if (it->second < 1) return kColresNA;
  const int64_t grid = std::min<int64_t>(a.nitems, (int64_t)c.sm_count * it->second);
  kern<<<(unsigned)grid, RT_NW * 32, smem, c.stream>>>(tm, nrows, v, out, a);
  c.launches++;
  return 0;
}

}  // namespace cb
