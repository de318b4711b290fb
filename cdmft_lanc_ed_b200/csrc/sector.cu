// sector.cu -- sector Fock maps, per-spin CSR (ED_SPARSE_MATRIX), ED_SPARSE_MAP and the
// build_Hv_sector / delete_Hv_sector lifecycle as sm_100a kernels.
//
// Reference semantics reproduced bit-exactly for the integer objects:
//   Hs(s)%map   = ascending Ns-bit integers with popcount N      ED_SETUP.f90:749-769
//   spH0ups/dws = row-wise lists, columns ascending (insertion order of the jup-outermost
//                 loop of sparse/H_up.f90:1-89 with sp_insert_element, ED_SPARSE_MATRIX.f90:254-284)
//   sparse_map  = per impurity configuration the bath configurations and sector indices in
//                 ascending-state order                          ED_SETUP.f90:757-759
// The scan over 2^Ns integers + recursive binary_search (ED_SETUP.f90:1044-1061) of the
// reference is replaced by ranking with two Lin tables: rank(s) = hi[s >> L] + lo[s & (2^L-1)].
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "ctx.h"

namespace cb {

__device__ __forceinline__ int32_t lin_rank(const int32_t *__restrict__ lin_lo, const int32_t *__restrict__ lin_hi,
                                            int lbits, uint32_t s) {
  return __ldg(lin_hi + (s >> lbits)) + __ldg(lin_lo + (s & ((1u << lbits) - 1u)));
}

// one thread per Ns-bit integer: the members of the sector write themselves at their rank
__global__ void k_sector_map(int ns, int npart, const int32_t *__restrict__ lin_lo, const int32_t *__restrict__ lin_hi,
                             int lbits, int32_t *__restrict__ map) {
  uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= (1u << ns)) return;
  if (__popc(s) != npart) return;
  map[lin_rank(lin_lo, lin_hi, lbits, s)] = (int32_t)s;
}

// spin-local diagonal: f(i) = const + sum_p e[p] n_p + sum_{a<b} S[a][b] n_a n_b   (sparse/H_local.f90)
__global__ void k_spin_diag(int64_t n, const int32_t *__restrict__ map, int ns, int nimp, const double *__restrict__ e,
                            const double *__restrict__ spair, double const_add, double *__restrict__ f) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t s = (uint32_t)map[i];
  double acc = const_add;
  for (int p = 0; p < ns; p++)
    if ((s >> p) & 1u) acc += e[p];
  for (int a = 0; a < nimp; a++)
    if ((s >> a) & 1u)
      for (int b = a + 1; b < nimp; b++)
        if ((s >> b) & 1u) acc += spair[a * nimp + b];
  f[i] = acc;
}

// fermionic sign of c^+_a c_b on a state: (-1)^(#occupied orbitals strictly between a and b)
// == sg1*sg2 of c(b,m,k1,sg1); cdg(a,k1,k2,sg2)  (ED_SETUP.f90:807-833)
__device__ __forceinline__ double hop_sign(uint32_t s, int a, int b) {
  int lo = min(a, b), hi = max(a, b);
  uint32_t between = ((1u << hi) - 1u) & ~((2u << lo) - 1u);
  return (__popc(s & between) & 1) ? -1.0 : 1.0;
}

// row lengths: entry (row k2, col m) exists for term (a,b) iff k2 has a occupied and b empty
__global__ void k_csr_count(int64_t n, const int32_t *__restrict__ map, const Term *__restrict__ terms, int nterms,
                            int32_t *__restrict__ rowlen) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t s = (uint32_t)map[i];
  int cnt = 0;
  for (int t = 0; t < nterms; t++) {
    int a = terms[t].a, b = terms[t].b;
    cnt += (int)(((s >> a) & 1u) & (~(s >> b) & 1u));
  }
  rowlen[i] = cnt;
}

// exclusive scan of row lengths, single block (n <= ~1e6 rows; runs once per sector)
__global__ void k_exclusive_scan(int64_t n, const int32_t *__restrict__ in, int32_t *__restrict__ out) {
  __shared__ int64_t carry;
  __shared__ int32_t buf[1024];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < n; base += 1024) {
    int64_t i = base + threadIdx.x;
    int32_t x = i < n ? in[i] : 0;
    buf[threadIdx.x] = x;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
      int32_t y = threadIdx.x >= off ? buf[threadIdx.x - off] : 0;
      __syncthreads();
      buf[threadIdx.x] += y;
      __syncthreads();
    }
    if (i < n) out[i] = (int32_t)(carry + buf[threadIdx.x] - x);
    __syncthreads();
    if (threadIdx.x == 1023) carry += buf[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[n] = (int32_t)carry;
}

// fill rows, then sort each row by column (rows are short: <= nterms entries)
__global__ void k_csr_fill(int64_t n, const int32_t *__restrict__ map, const int32_t *__restrict__ lin_lo,
                           const int32_t *__restrict__ lin_hi, int lbits, const Term *__restrict__ terms, int nterms,
                           const int32_t *__restrict__ rowptr, int32_t *__restrict__ col, double2 *__restrict__ val) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t s = (uint32_t)map[i];
  int32_t p0 = rowptr[i], p = p0;
  for (int t = 0; t < nterms; t++) {
    int a = terms[t].a, b = terms[t].b;
    if (((s >> a) & 1u) && !((s >> b) & 1u)) {
      uint32_t m = (s & ~(1u << a)) | (1u << b);  // source state: k2 = c^+_a c_b m
      double sg = hop_sign(s, a, b);
      int32_t j = lin_rank(lin_lo, lin_hi, lbits, m);
      double2 h = make_double2(terms[t].re * sg, terms[t].im * sg);
      int32_t q = p;  // insertion sort by column
      while (q > p0 && col[q - 1] > j) {
        col[q] = col[q - 1];
        val[q] = val[q - 1];
        q--;
      }
      col[q] = j;
      val[q] = h;
      p++;
    }
  }
}

__global__ void k_csr_to_ell(int64_t n, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                             const double2 *__restrict__ val, int ell_w, int32_t *__restrict__ ell_col,
                             double2 *__restrict__ ell_val) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int32_t p0 = rowptr[i], len = rowptr[i + 1] - p0;
  for (int k = 0; k < ell_w; k++) {
    bool in = k < len;
    ell_col[(int64_t)k * n + i] = in ? col[p0 + k] : (int32_t)i;
    ell_val[(int64_t)k * n + i] = in ? val[p0 + k] : make_double2(0.0, 0.0);
  }
}

// ED_SPARSE_MAP (ED_SPARSE_MAP.f90:101-121 via ED_SETUP.f90:757-759): counting pass + ordered fill.
// States of one impurity configuration appear in ascending sector index, so the position inside
// the row is the number of earlier sector states with the same impurity bits.
__global__ void k_spmap_count(int64_t n, const int32_t *__restrict__ map, int nimp, unsigned long long *__restrict__ cnt) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t s = (uint32_t)map[i];
  atomicAdd(&cnt[s & ((1u << nimp) - 1u)], 1ull);
}
__global__ void k_spmap_fill(int64_t n, const int32_t *__restrict__ map, int ns, int nimp, int npart,
                             const int64_t *__restrict__ rowptr, const int32_t *__restrict__ binom /*[31*31]*/,
                             int32_t *__restrict__ bath_state, int32_t *__restrict__ sector_indx) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t s = (uint32_t)map[i];
  uint32_t iimp = s & ((1u << nimp) - 1u), ibath = s >> nimp;
  // position = rank of ibath among (Ns-Nimp)-bit strings with the same popcount (ascending)
  int k = 0, r = 0;
  for (int p = 0; p < ns - nimp; p++)
    if ((ibath >> p) & 1u) { k++; r += binom[p * 31 + k]; }
  (void)npart;
  int64_t q = rowptr[iimp] + r;
  bath_state[q] = (int32_t)ibath;
  sector_indx[q] = (int32_t)(i + 1);
}

// ------------------------------------------------------------------------------------
static int64_t binom64(int n, int k) {
  if (k < 0 || k > n) return 0;
  int64_t r = 1;
  for (int i = 1; i <= k; i++) r = r * (n - k + i) / i;
  return r;
}

#define LAUNCH_1D(kernel, n, ...)                                                  \
  do {                                                                             \
    int64_t nb__ = ((n) + 255) / 256;                                              \
    if (nb__ > 0) {                                                                \
      kernel<<<(unsigned)nb__, 256, 0, c.stream>>>(__VA_ARGS__);                   \
      c.launches++;                                                                \
    }                                                                              \
  } while (0)

// ------------------------------------------------------------------------------------
// Conflict-free schedule for the column-resident kernels (see Sched in ctx.h).
// One group = G consecutive rows.  Its entries are the edges of a bipartite multigraph rows x banks
// (bank = source row mod G = the shared-memory bank group the gather hits); a proper edge colouring
// with Delta = max(longest row, busiest bank) colours (Koenig) makes every colour class a step in
// which the G lanes read G different banks.  Alternating-path recolouring, O(E * G) per group.
// `natural` skips the colouring (step k = k-th entry of the row) -- kept for comparison runs.
// ------------------------------------------------------------------------------------
struct SchedHost {
  int G = 0, fmt = 0, nwarps = 0;
  int32_t ntask = 0;
  int64_t nquads = 0;
  std::vector<int32_t> tbase, qbase;  // [nwarps+1] first task / first quad of each warp's stream
  std::vector<uint32_t> meta;         // [ntask*32] uint4 per task and lane: f_row (8 B), row (-1 = none), mu | units << 16
  std::vector<uint32_t> words;        // [(nquads+4)*32] per unit and lane: uint4 = 4 steps (fmt 0,1,3) or uint32 = 2 steps (fmt 2)
};

static void build_schedule_host(int64_t n, const std::vector<int32_t> &rowptr, const std::vector<int32_t> &col,
                                const std::vector<uint8_t> &code, int G, int fmt, bool natural, int nwarps,
                                const double *f_row, const uint32_t *mu_row, SchedHost &out) {
  const bool fast = fmt == 1 || fmt == 2, w16 = fmt == 2, f4 = fmt == 3;
  const int64_t ngroups = (n + G - 1) / G;
  const int64_t npad = ngroups * G;  // G zero elements follow the column in shared memory: [npad, npad+G)
  const int esz = G == 8 ? 16 : 8;   // bytes per vector element
  // word of an entry (source row j, code) / of an idle lane parked on the zero element of bank b.
  // fmt 0 (general): (j << 7) | coefficient id, id 0 = 0.0       four steps per uint4
  // fmt 1 (fast32):  (negative << 31) | (j * esz) | class         four steps per uint4, address one AND away
  // fmt 2 (fast16):  (negative << 15) | (class << 14) | j         two steps per uint32 (needs npad+G <= 2^14)
  // fmt 3 (fast4):   (negative << 31) | (j * esz) | imaginary << 2 | class (2 bits)   four steps per uint4
  auto mkword = [&](int64_t j, uint32_t cd) -> uint32_t {
    if (f4) return ((cd & 1u) << 31) | (uint32_t)(j * esz) | ((cd >> 1) & 3u) | (((cd >> 3) & 1u) << 2);
    if (w16) return ((cd & 1u) << 15) | (((cd >> 1) & 1u) << 14) | (uint32_t)j;
    if (fast) return ((cd & 1u) << 31) | (uint32_t)(j * esz) | ((cd >> 1) & 1u);
    return ((uint32_t)j << 7) | cd;
  };
  auto idle = [&](int b) -> uint32_t {
    if (w16) return (uint32_t)(npad + b);
    return (fast || f4) ? (uint32_t)((npad + b) * esz) : ((uint32_t)(npad + b) << 7);
  };
  std::vector<std::vector<uint32_t>> steps(ngroups);  // [K_g * G] words of each group
  std::vector<int32_t> K(ngroups, 0);
  struct Edge { int r, b, c; uint32_t w; };
  std::vector<Edge> ed;
  std::vector<int> at_row, at_bank, path;
  for (int64_t g = 0; g < ngroups; g++) {
    const int64_t r0 = g * G;
    const int nr = (int)std::min<int64_t>(G, n - r0);
    ed.clear();
    int deg_r[16] = {0}, deg_b[16] = {0};
    for (int r = 0; r < nr; r++)
      for (int32_t p = rowptr[r0 + r]; p < rowptr[r0 + r + 1]; p++) {
        const int b = col[p] % G;
        ed.push_back({r, b, natural ? deg_r[r] : -1, mkword(col[p], code[p])});
        deg_r[r]++; deg_b[b]++;
      }
    int delta = 0;
    for (int x = 0; x < G; x++) delta = std::max(delta, natural ? deg_r[x] : std::max(deg_r[x], deg_b[x]));
    if (!natural && delta > 0) {
      at_row.assign((size_t)G * delta, -1);
      at_bank.assign((size_t)G * delta, -1);
      for (int e = 0; e < (int)ed.size(); e++) {
        const int r = ed[e].r, b = ed[e].b;
        int a = 0, bb = 0;
        while (at_row[(size_t)r * delta + a] >= 0) a++;
        while (at_bank[(size_t)b * delta + bb] >= 0) bb++;
        if (a != bb) {
          // free colour a at bank b: flip a <-> bb along the alternating path that leaves b with colour a
          path.clear();
          int cur = b, want = a;
          bool at_bank_side = true;
          for (;;) {
            const int e2 = at_bank_side ? at_bank[(size_t)cur * delta + want] : at_row[(size_t)cur * delta + want];
            if (e2 < 0) break;
            path.push_back(e2);
            cur = at_bank_side ? ed[e2].r : ed[e2].b;
            at_bank_side = !at_bank_side;
            want = want == a ? bb : a;
          }
          for (int e2 : path) {
            at_row[(size_t)ed[e2].r * delta + ed[e2].c] = -1;
            at_bank[(size_t)ed[e2].b * delta + ed[e2].c] = -1;
          }
          for (int e2 : path) {
            ed[e2].c = ed[e2].c == a ? bb : a;
            at_row[(size_t)ed[e2].r * delta + ed[e2].c] = e2;
            at_bank[(size_t)ed[e2].b * delta + ed[e2].c] = e2;
          }
        }
        ed[e].c = a;
        at_row[(size_t)r * delta + a] = e;
        at_bank[(size_t)b * delta + a] = e;
      }
    }
    K[g] = delta;
    // idle lanes of a step are parked on the zero elements of the banks the step leaves free
    std::vector<int8_t> lane_bank((size_t)delta * G, -1);
    steps[g].assign((size_t)delta * G, 0u);
    for (auto &e : ed) { steps[g][(size_t)e.c * G + e.r] = e.w; lane_bank[(size_t)e.c * G + e.r] = (int8_t)e.b; }
    for (int k = 0; k < delta; k++) {
      bool used[16] = {false};
      for (int r = 0; r < G; r++) if (lane_bank[(size_t)k * G + r] >= 0) used[lane_bank[(size_t)k * G + r]] = true;
      int nb = 0;
      for (int r = 0; r < G; r++)
        if (lane_bank[(size_t)k * G + r] < 0) {
          while (nb < G - 1 && used[nb]) nb++;  // natural order may leave fewer free banks than idle lanes
          steps[g][(size_t)k * G + r] = idle(nb);
          used[nb] = true;
        }
    }
  }
  // warp tasks: groups sorted by step count (descending, stable), 32/G per task; tasks dealt round-robin to
  // the nwarps warps of the CTA and stored warp-major, so that a warp's words form one contiguous stream
  const int per = 32 / G;
  std::vector<int32_t> order(ngroups);
  for (int64_t g = 0; g < ngroups; g++) order[g] = (int32_t)g;
  std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) { return K[x] > K[y]; });
  const int64_t ntask = (ngroups + per - 1) / per;
  out.G = G; out.fmt = fmt; out.nwarps = nwarps; out.ntask = (int32_t)ntask;
  out.tbase.assign(nwarps + 1, 0);
  out.qbase.assign(nwarps + 1, 0);
  out.meta.assign((size_t)ntask * 32 * 4, 0u);
  std::vector<int64_t> sorted_of_task(ntask);  // warp-major task -> index in the sorted list
  {
    int64_t t = 0;
    for (int w = 0; w < nwarps; w++) {
      out.tbase[w] = (int32_t)t;
      for (int64_t st = w; st < ntask; st += nwarps) sorted_of_task[t++] = st;
    }
    out.tbase[nwarps] = (int32_t)t;
  }
  const int spu = w16 ? 2 : 4;  // steps per unit
  std::vector<int32_t> qoff(ntask + 1, 0);
  for (int64_t t = 0; t < ntask; t++) {
    int kt = 0;
    for (int q = 0; q < per; q++) {
      const int64_t idx = sorted_of_task[t] * per + q;
      if (idx < ngroups) kt = std::max(kt, K[order[idx]]);
    }
    qoff[t + 1] = qoff[t] + (kt + spu - 1) / spu;  // one load per unit: uint4 = 4 steps of 32-bit words, uint32 = 2 steps of 16-bit words
  }
  for (int w = 0; w <= nwarps; w++) out.qbase[w] = qoff[out.tbase[w]];
  out.nquads = qoff[ntask];
  const size_t wpq = w16 ? 1 : 4;  // 32-bit registers per lane and unit
  out.words.assign(((size_t)out.nquads + 4) * 32 * wpq, 0u);  // four units of slack: the prefetch runs ahead unguarded
  for (int64_t t = 0; t < ntask; t++) {
    const int nq = qoff[t + 1] - qoff[t];
    for (int q = 0; q < per; q++) {
      const int64_t idx = sorted_of_task[t] * per + q;
      const int32_t g = idx < ngroups ? order[idx] : -1;
      for (int r = 0; r < G; r++) {
        const int lane = q * G + r;
        const int64_t i = g >= 0 ? (int64_t)g * G + r : -1;
        uint32_t *m = &out.meta[((size_t)t * 32 + lane) * 4];
        const bool valid = i >= 0 && i < n;
        double f = valid && f_row ? f_row[i] : 0.0;
        memcpy(m, &f, 8);
        m[2] = valid ? (uint32_t)i : 0xFFFFFFFFu;
        m[3] = (valid && mu_row ? (mu_row[i] & 0xFFFFu) : 0u) | ((uint32_t)nq << 16);
        for (int k = 0; k < nq * spu; k++) {
          const uint32_t w = (g >= 0 && k < K[g]) ? steps[g][(size_t)k * G + r] : idle(r);
          const size_t base = (((size_t)qoff[t] + k / spu) * 32 + lane) * wpq;
          if (w16) out.words[base] |= w << ((k & 1) * 16);
          else out.words[base + (k & 3)] = w;
        }
      }
    }
  }
}

static int upload_schedule(Sched &sc, const SchedHost &h) {
  Ctx &c = ctx();
  sc.G = h.G; sc.fmt = h.fmt; sc.nwarps = h.nwarps; sc.ntask = h.ntask; sc.nquads = h.nquads;
  CB_CHECK(dev_alloc(&sc.tbase, (int64_t)h.tbase.size()));
  CB_CHECK(dev_alloc(&sc.qbase, (int64_t)h.qbase.size()));
  CB_CHECK(dev_alloc(&sc.meta, (int64_t)h.meta.size()));
  CB_CHECK(dev_alloc(&sc.words, (int64_t)h.words.size()));
  CB_CUDA(cudaMemcpyAsync(sc.tbase, h.tbase.data(), h.tbase.size() * 4, cudaMemcpyHostToDevice, c.stream));
  CB_CUDA(cudaMemcpyAsync(sc.qbase, h.qbase.data(), h.qbase.size() * 4, cudaMemcpyHostToDevice, c.stream));
  if (!h.meta.empty()) CB_CUDA(cudaMemcpyAsync(sc.meta, h.meta.data(), h.meta.size() * 4, cudaMemcpyHostToDevice, c.stream));
  CB_CUDA(cudaMemcpyAsync(sc.words, h.words.data(), h.words.size() * 4, cudaMemcpyHostToDevice, c.stream));
  CB_CUDA(cudaStreamSynchronize(c.stream));
  return 0;
}

// ------------------------------------------------------------------------------------
// Block-split schedules (ColBlk in ctx.h) for columns larger than shared memory: one build_schedule_host per
// row block on the in-block part of the CSR (relative indices) plus, per warp task, the off-block entries of
// its 32 rows as a lane-parallel stream.  code3 of an off-block fast word: class (2 bits) | imaginary << 2.
// ------------------------------------------------------------------------------------
struct ColBlkHost {  // host image of ColBlk (what build_colblk uploads); filled instead of uploading for the CPU tests
  std::vector<int4> blk;
  std::vector<int32_t> tbase, qbase;
  std::vector<uint32_t> meta, words, woff;
  std::vector<uint2> toff;
  int nwarps = 0;
  int64_t max_rows = 0;
};

static int build_colblk(SpinOp &op, ColBlk &cb, int ns, const std::vector<int32_t> &rowptr, const std::vector<int32_t> &col,
                        const std::vector<uint8_t> &code, int G, int fmt, bool natural, int64_t cap_rows,
                        const double *f_row, const uint32_t *mu_row, ColBlkHost *host_out = nullptr) {
  Ctx &c = ctx();
  int t = 0;
  for (; t <= ns; t++) {
    int64_t mx = 0;
    for (int k = 0; k <= t; k++) mx = std::max(mx, binom64(ns - t, op.npart - k));
    if (mx <= cap_rows) break;
  }
  std::vector<int4> blk;
  std::vector<int32_t> tbase, qbase;
  std::vector<uint32_t> meta, words, woff;
  std::vector<uint2> toff;
  const int nwarps = kColblkWarps;
  const bool fastw = fmt == 1 || fmt == 3;
  // in-block words: 16 bits (sign | class | row relative to the block) when the two-magnitude decode applies and the
  // largest block fits 14-bit row numbers -- half the operator traffic of the 32-bit words; the off-block words stay 32 bits
  int fmt_in = fmt;
  {
    int64_t mx = 0;
    for (int k = 0; k <= t; k++) mx = std::max(mx, binom64(ns - t, op.npart - k));
    if (fmt == 1 && G == 8 && mx + 32 <= (1 << 14) && !host_out) fmt_in = 2;
  }
  const uint32_t NONE = fastw ? 0xFFFFFFFFu : 0u;
  std::vector<int32_t> bstart;  // first row of every block + the end
  {
    int64_t st = 0;
    for (uint32_t P = 0; P < (1u << t); P++) {
      const int64_t ng = binom64(ns - t, op.npart - __builtin_popcount(P));
      if (ng <= 0) continue;
      bstart.push_back((int32_t)st);
      st += ng;
    }
    bstart.push_back((int32_t)st);
  }
  int64_t start = 0, mxrows = 0, task0 = 0, unit0 = 0;
  for (uint32_t P = 0; P < (1u << t); P++) {
    const int64_t ng = binom64(ns - t, op.npart - __builtin_popcount(P));
    if (ng <= 0) continue;
    const int64_t g0 = start;
    start += ng;
    mxrows = std::max(mxrows, ng);
    // in-block sub-matrix with relative indices
    std::vector<int32_t> rp(ng + 1, 0), cl;
    std::vector<uint8_t> cd;
    for (int64_t i = 0; i < ng; i++) {
      for (int32_t p = rowptr[g0 + i]; p < rowptr[g0 + i + 1]; p++)
        if (col[p] >= g0 && col[p] < g0 + ng) { cl.push_back((int32_t)(col[p] - g0)); cd.push_back(code[p]); }
      rp[i + 1] = (int32_t)cl.size();
    }
    SchedHost sh;
    build_schedule_host(ng, rp, cl, cd, G, fmt_in, natural, nwarps, f_row ? f_row + g0 : nullptr, mu_row ? mu_row + g0 : nullptr, sh);
    blk.push_back(make_int4((int)g0, (int)ng, (int)task0, (int)unit0));
    tbase.insert(tbase.end(), sh.tbase.begin(), sh.tbase.end());
    qbase.insert(qbase.end(), sh.qbase.begin(), sh.qbase.end());
    meta.insert(meta.end(), sh.meta.begin(), sh.meta.end());
    words.insert(words.end(), sh.words.begin(), sh.words.end());
    // off-block streams, lane-parallel per task.  Steps are aligned by SOURCE BLOCK: all lanes of a step gather
    // from the same block, entries in ascending order, so that neighbouring rows taking the same hop read
    // neighbouring sources (a hop between two top bits keeps the in-block index: perfectly coalesced).
    auto block_of = [&](int32_t j) -> int {
      int lo = 0, hi = (int)bstart.size() - 2;
      while (lo < hi) { const int mid = (lo + hi + 1) / 2; if (bstart[mid] <= j) lo = mid; else hi = mid - 1; }
      return lo;
    };
    const int nbtot = (int)bstart.size() - 1;
    std::vector<int> cnt((size_t)32 * nbtot), base(nbtot);
    for (int32_t tk = 0; tk < sh.ntask; tk++) {
      std::fill(cnt.begin(), cnt.end(), 0);
      for (int lane = 0; lane < 32; lane++) {
        const uint32_t rel = sh.meta[((size_t)tk * 32 + lane) * 4 + 2];
        if (rel == 0xFFFFFFFFu) continue;
        for (int32_t p = rowptr[g0 + rel]; p < rowptr[g0 + rel + 1]; p++)
          if (!(col[p] >= g0 && col[p] < g0 + ng)) cnt[(size_t)lane * nbtot + block_of(col[p])]++;
      }
      int noff = 0;
      for (int bb = 0; bb < nbtot; bb++) {
        int mxc = 0;
        for (int lane = 0; lane < 32; lane++) mxc = std::max(mxc, cnt[(size_t)lane * nbtot + bb]);
        base[bb] = noff;
        noff += mxc;
      }
      const size_t ob = woff.size() / 32;
      toff.push_back(make_uint2((uint32_t)ob, (uint32_t)noff));
      woff.resize(woff.size() + (size_t)noff * 32, NONE);
      for (int lane = 0; lane < 32; lane++) {
        const uint32_t rel = sh.meta[((size_t)tk * 32 + lane) * 4 + 2];
        if (rel == 0xFFFFFFFFu) continue;
        std::vector<int> fill(nbtot, 0);
        for (int32_t p = rowptr[g0 + rel]; p < rowptr[g0 + rel + 1]; p++) {
          if (col[p] >= g0 && col[p] < g0 + ng) continue;
          const int bb = block_of(col[p]);
          const uint32_t cdp = code[p];
          uint32_t w;
          if (fmt == 3) w = ((cdp & 1u) << 31) | ((uint32_t)col[p] << 3) | ((cdp >> 1) & 3u) | (((cdp >> 3) & 1u) << 2);
          else if (fmt == 1) w = ((cdp & 1u) << 31) | ((uint32_t)col[p] << 3) | ((cdp >> 1) & 1u);
          else w = ((uint32_t)col[p] << 7) | cdp;
          woff[(ob + base[bb] + fill[bb]) * 32 + lane] = w;
          fill[bb]++;
        }
      }
    }
    task0 += sh.ntask;
    unit0 += (int64_t)(sh.words.size() / (32 * (fmt_in == 2 ? 1 : 4)));  // one uint4 (uint32 for 16-bit words) per unit and lane, slack included
  }
  if (start != op.n) return fail("internal: column blocks do not cover the sector");
  woff.resize(woff.size() + 64, NONE);
  if (host_out) {  // CPU inspection (cdmft_b200_colblk_host): no device involved
    host_out->blk = blk; host_out->tbase = tbase; host_out->qbase = qbase; host_out->meta = meta; host_out->words = words;
    host_out->woff = woff; host_out->toff = toff; host_out->nwarps = nwarps; host_out->max_rows = mxrows;
    return 0;
  }
  cb.nblk = (int32_t)blk.size(); cb.nwarps = nwarps; cb.fmt = fmt_in; cb.G = G; cb.max_rows = mxrows;
  cb.even_blocks = true;
  for (auto &q : blk) cb.even_blocks = cb.even_blocks && !(q.x & 1) && !(q.y & 1);
  CB_CHECK(dev_alloc(&cb.blk, (int64_t)blk.size()));
  CB_CHECK(dev_alloc(&cb.tbase, (int64_t)tbase.size()));
  CB_CHECK(dev_alloc(&cb.qbase, (int64_t)qbase.size()));
  CB_CHECK(dev_alloc(&cb.meta, (int64_t)meta.size()));
  CB_CHECK(dev_alloc(&cb.words, (int64_t)words.size()));
  CB_CHECK(dev_alloc(&cb.toff, (int64_t)toff.size()));
  CB_CHECK(dev_alloc(&cb.woff, (int64_t)woff.size()));
  CB_CUDA(cudaMemcpyAsync(cb.blk, blk.data(), blk.size() * sizeof(int4), cudaMemcpyHostToDevice, c.stream));
  CB_CUDA(cudaMemcpyAsync(cb.tbase, tbase.data(), tbase.size() * 4, cudaMemcpyHostToDevice, c.stream));
  CB_CUDA(cudaMemcpyAsync(cb.qbase, qbase.data(), qbase.size() * 4, cudaMemcpyHostToDevice, c.stream));
  CB_CUDA(cudaMemcpyAsync(cb.meta, meta.data(), meta.size() * 4, cudaMemcpyHostToDevice, c.stream));
  CB_CUDA(cudaMemcpyAsync(cb.words, words.data(), words.size() * 4, cudaMemcpyHostToDevice, c.stream));
  CB_CUDA(cudaMemcpyAsync(cb.toff, toff.data(), toff.size() * sizeof(uint2), cudaMemcpyHostToDevice, c.stream));
  CB_CUDA(cudaMemcpyAsync(cb.woff, woff.data(), woff.size() * 4, cudaMemcpyHostToDevice, c.stream));
  CB_CUDA(cudaStreamSynchronize(c.stream));
  return 0;
}

int build_spin_op(SpinOp &op, int32_t npart, const std::vector<Term> &terms, const std::vector<double> &e,
                  double const_add, bool want_csr) {
  Ctx &c = ctx();
  const int ns = c.ns;
  op.npart = npart;
  op.n = binom64(ns, npart);
  op.real_h = c.real_h;
  const int lbits = ns / 2, hbits = ns - lbits;
  // Lin tables (host, tiny)
  std::vector<int32_t> lo((size_t)1 << lbits), hi((size_t)1 << hbits);
  {
    std::vector<int32_t> cnt(lbits + 1, 0);
    for (uint32_t x = 0; x < (1u << lbits); x++) lo[x] = cnt[__builtin_popcount(x)]++;
    int64_t off = 0;
    for (uint32_t h = 0; h < (1u << hbits); h++) {
      hi[h] = (int32_t)off;
      int need = npart - __builtin_popcount(h);
      if (need >= 0 && need <= lbits) off += binom64(lbits, need);
    }
    if (off != op.n) return fail("internal: Lin table size mismatch");
  }
  CB_CHECK(dev_alloc(&op.lin_lo, (int64_t)lo.size()));
  CB_CHECK(dev_alloc(&op.lin_hi, (int64_t)hi.size()));
  CB_CUDA(cudaMemcpyAsync(op.lin_lo, lo.data(), lo.size() * 4, cudaMemcpyHostToDevice, c.stream));
  CB_CUDA(cudaMemcpyAsync(op.lin_hi, hi.data(), hi.size() * 4, cudaMemcpyHostToDevice, c.stream));
  CB_CHECK(dev_alloc(&op.map, op.n));
  LAUNCH_1D(k_sector_map, (int64_t)1 << ns, ns, npart, op.lin_lo, op.lin_hi, lbits, op.map);
  // terms + diagonal coefficients
  op.nterms = (int32_t)terms.size();
  CB_CHECK(dev_alloc(&op.terms, (int64_t)terms.size()));
  if (!terms.empty())
    CB_CUDA(cudaMemcpyAsync(op.terms, terms.data(), terms.size() * sizeof(Term), cudaMemcpyHostToDevice, c.stream));
  double *d_e = nullptr, *d_sp = nullptr;
  CB_CHECK(dev_alloc(&d_e, ns));
  CB_CHECK(dev_alloc(&d_sp, (int64_t)c.nimp * c.nimp));
  CB_CUDA(cudaMemcpyAsync(d_e, e.data(), ns * sizeof(double), cudaMemcpyHostToDevice, c.stream));
  CB_CUDA(cudaMemcpyAsync(d_sp, c.spair.data(), c.spair.size() * sizeof(double), cudaMemcpyHostToDevice, c.stream));
  CB_CHECK(dev_alloc(&op.f, op.n));
  LAUNCH_1D(k_spin_diag, op.n, op.n, op.map, ns, c.nimp, d_e, d_sp, const_add, op.f);
  // row lengths are needed by both modes (ELL width / statistics)
  CB_CHECK(dev_alloc(&op.rowlen, op.n));
  LAUNCH_1D(k_csr_count, op.n, op.n, op.map, op.terms, op.nterms, op.rowlen);
  if (want_csr) {
    CB_CHECK(dev_alloc(&op.rowptr, op.n + 1));
    k_exclusive_scan<<<1, 1024, 0, c.stream>>>(op.n, op.rowlen, op.rowptr);
    c.launches++;
    int32_t nnz32 = 0;
    CB_CUDA(cudaMemcpyAsync(&nnz32, op.rowptr + op.n, 4, cudaMemcpyDeviceToHost, c.stream));
    CB_CUDA(cudaStreamSynchronize(c.stream));
    op.nnz = nnz32;
    CB_CHECK(dev_alloc(&op.col, op.nnz));
    CB_CHECK(dev_alloc(&op.val, op.nnz));
    LAUNCH_1D(k_csr_fill, op.n, op.n, op.map, op.lin_lo, op.lin_hi, lbits, op.terms, op.nterms, op.rowptr, op.col, op.val);
    // ELL width = longest row
    std::vector<int32_t> rl(op.n);
    CB_CUDA(cudaMemcpyAsync(rl.data(), op.rowlen, op.n * 4, cudaMemcpyDeviceToHost, c.stream));
    CB_CUDA(cudaStreamSynchronize(c.stream));
    op.ell_w = 0;
    for (auto x : rl) op.ell_w = std::max(op.ell_w, x);
    CB_CHECK(dev_alloc(&op.ell_col, (int64_t)op.ell_w * op.n));
    CB_CHECK(dev_alloc(&op.ell_val, (int64_t)op.ell_w * op.n));
    LAUNCH_1D(k_csr_to_ell, op.n, op.n, op.rowptr, op.col, op.val, op.ell_w, op.ell_col, op.ell_val);
  }
  // distinct signed coefficients -> 7-bit ids (host; the per-spin matrices are small)
  if (want_csr && op.nnz > 0 && op.n < (1 << 21)) {
    std::vector<double2> hval(op.nnz);
    CB_CUDA(cudaMemcpyAsync(hval.data(), op.val, op.nnz * sizeof(double2), cudaMemcpyDeviceToHost, c.stream));
    CB_CUDA(cudaStreamSynchronize(c.stream));
    std::vector<double2> table;
    table.push_back(make_double2(0.0, 0.0));
    std::vector<uint8_t> ids(op.nnz);
    bool ok = true;
    for (int64_t k = 0; k < op.nnz && ok; k++) {
      int found = -1;
      for (size_t t = 1; t < table.size(); t++)
        if (table[t].x == hval[k].x && table[t].y == hval[k].y) { found = (int)t; break; }
      if (found < 0) {
        if (table.size() >= 128) { ok = false; break; }
        table.push_back(hval[k]);
        found = (int)table.size() - 1;
      }
      ids[k] = (uint8_t)found;
    }
    if (ok) {
      op.ncoef = (int32_t)table.size();
      CB_CHECK(dev_alloc(&op.coef, 128));
      table.resize(128, make_double2(0.0, 0.0));
      CB_CUDA(cudaMemcpyAsync(op.coef, table.data(), 128 * sizeof(double2), cudaMemcpyHostToDevice, c.stream));
      CB_CUDA(cudaStreamSynchronize(c.stream));
      std::vector<int32_t> hrp, hcol;
      if (op.n < (1 << 24)) {
        hrp.resize(op.n + 1); hcol.resize(op.nnz);
        CB_CUDA(cudaMemcpy(hrp.data(), op.rowptr, (op.n + 1) * 4, cudaMemcpyDeviceToHost));
        CB_CUDA(cudaMemcpy(hcol.data(), op.col, op.nnz * 4, cudaMemcpyDeviceToHost));
      }
      // fast decode: real H with at most two distinct |coefficient| -> code = class<<1 | negative
      std::vector<double> mags;
      bool fast = op.real_h;
      for (int t = 1; t < op.ncoef && fast; t++) {
        if (table[t].y != 0.0) { fast = false; break; }
        const double m = std::fabs(table[t].x);
        if (std::find(mags.begin(), mags.end(), m) == mags.end()) mags.push_back(m);
        if (mags.size() > 2) fast = false;
      }
      std::vector<uint8_t> code(ids);
      op.sc_fast = fast;
      if (fast) {
        op.sc_mag[0] = mags.size() > 0 ? mags[0] : 0.0;
        op.sc_mag[1] = mags.size() > 1 ? mags[1] : 0.0;
        for (int64_t k = 0; k < op.nnz; k++) {
          const double x = table[ids[k]].x;
          const int cls = (mags.size() > 1 && std::fabs(x) == mags[1]) ? 1 : 0;
          code[k] = (uint8_t)((cls << 1) | (std::signbit(x) ? 1 : 0));
        }
      }
      // second level: every coefficient purely real or purely imaginary, at most four distinct magnitudes
      // (complex hoppings such as the BHZ model's -ts*sigma_z + i*lambda/2*sigma_x): code4 = negative | class<<1 | imaginary<<3
      std::vector<uint8_t> code4;
      bool fast4 = !fast;
      {
        std::vector<double> mags4;
        for (int t = 1; t < op.ncoef && fast4; t++) {
          if (table[t].x != 0.0 && table[t].y != 0.0) { fast4 = false; break; }
          const double m = std::fabs(table[t].y != 0.0 ? table[t].y : table[t].x);
          if (std::find(mags4.begin(), mags4.end(), m) == mags4.end()) mags4.push_back(m);
          if (mags4.size() > 4) fast4 = false;
        }
        if (fast4) {
          for (int q = 0; q < 4; q++) op.sc_mag[q] = q < (int)mags4.size() ? mags4[q] : 0.0;
          code4.resize(op.nnz);
          for (int64_t k = 0; k < op.nnz; k++) {
            const double2 h = table[ids[k]];
            const bool im = h.y != 0.0;
            const double x = im ? h.y : h.x;
            const int cls = (int)(std::find(mags4.begin(), mags4.end(), std::fabs(x)) - mags4.begin());
            code4[k] = (uint8_t)((std::signbit(x) ? 1 : 0) | (cls << 1) | (im ? 8 : 0));
          }
        }
      }
      // operator streams of the tile-resident row pass (rowtile.cu)
      if (!hrp.empty()) {
        const bool f4 = (fast || fast4) && c.opt.fast4;
        const int64_t cap = c.opt.rowres_cols > 0 ? std::min<int64_t>(c.opt.rowres_cols, rowtile_cap()) : rowtile_cap();
        CB_CHECK(build_rowtile(op, ns, hrp, hcol, f4 ? (fast ? code : code4) : ids, f4 ? (fast ? 1 : 2) : 0, std::max<int64_t>(4, cap)));
      }
      // schedules of the column-resident kernels (only when a column can live in shared memory)
      if ((size_t)op.n * 8 + 4096 <= 232448 && op.n < (1 << 24)) {
        const bool natural = c.opt.sched == 0;
        const int fmt = fast4 && c.opt.fast4 ? 3 : (!fast ? 0 : (op.n + 32 <= (1 << 14) ? 2 : 1));
        if (fmt == 3) code = code4;
        // per-row part of the diagonal and the row's impurity configuration travel with the schedule
        std::vector<double> hf(op.n);
        std::vector<int32_t> hmap(op.n);
        std::vector<uint32_t> hmu(op.n);
        CB_CUDA(cudaMemcpy(hf.data(), op.f, op.n * 8, cudaMemcpyDeviceToHost));
        CB_CUDA(cudaMemcpy(hmap.data(), op.map, op.n * 4, cudaMemcpyDeviceToHost));
        for (int64_t i = 0; i < op.n; i++) hmu[i] = (uint32_t)hmap[i] & ((1u << c.nimp) - 1u);
        SchedHost sh;
        if ((size_t)op.n * 16 + 4096 <= 232448) {
          build_schedule_host(op.n, hrp, hcol, code, 8, fmt, natural, 32, hf.data(), hmu.data(), sh);
          CB_CHECK(upload_schedule(op.sc8, sh));
        }
        if (op.real_h) {
          // 8-byte elements: two 512-thread CTAs per SM when two columns fit, else one of 1024
          const int nw = (size_t)op.n * 8 * 2 + 8192 <= 232448 ? 16 : 32;
          build_schedule_host(op.n, hrp, hcol, code, 16, fmt, natural, nw, hf.data(), hmu.data(), sh);
          CB_CHECK(upload_schedule(op.sc16, sh));
          if ((fmt == 1 || fmt == 2) && !(op.n & 1) && (size_t)(op.n + 32) * 16 + 4096 <= 232448) {  // two columns side by side
            if (nw != 32) build_schedule_host(op.n, hrp, hcol, code, 16, fmt, natural, 32, hf.data(), hmu.data(), sh);
            CB_CHECK(upload_schedule(op.sc16x2, sh));
          }
        }
      }
      // block-split schedules: columns larger than shared memory (or forced small blocks for the tests)
      {
        const int64_t forced = c.opt.colres_rows;
        const int fmtb = fast4 && c.opt.fast4 ? 3 : (fast ? 1 : 0);
        const std::vector<uint8_t> &codeb = fmtb == 3 ? code4 : (fast ? code : ids);
        std::vector<double> hf;
        std::vector<uint32_t> hmu;
        auto need_rowdata = [&]() -> int {
          if (!hf.empty()) return 0;
          hf.resize(op.n);
          std::vector<int32_t> hmap(op.n);
          hmu.resize(op.n);
          CB_CUDA(cudaMemcpy(hf.data(), op.f, op.n * 8, cudaMemcpyDeviceToHost));
          CB_CUDA(cudaMemcpy(hmap.data(), op.map, op.n * 4, cudaMemcpyDeviceToHost));
          for (int64_t i = 0; i < op.n; i++) hmu[i] = (uint32_t)hmap[i] & ((1u << c.nimp) - 1u);
          return 0;
        };
        const int64_t room = 232448 - 128 - 2048 - (((int64_t)8 << c.nimp) + 127) / 128 * 128;
        if (!hrp.empty() && (forced > 0 || (int64_t)(op.n + 32) * 16 > room)) {
          CB_CHECK(need_rowdata());
          const int64_t cap = forced > 0 ? std::max<int64_t>(8, forced) : room / 16 - 32;
          CB_CHECK(build_colblk(op, op.cb8, ns, hrp, hcol, codeb, 8, fmtb, c.opt.sched == 0, cap, hf.data(), hmu.data()));
        }
        if (!hrp.empty() && op.real_h && (forced > 0 || (int64_t)(op.n + 32) * 8 > room)) {
          CB_CHECK(need_rowdata());
          const int64_t cap = forced > 0 ? std::max<int64_t>(16, forced) : room / 8 - 32;
          CB_CHECK(build_colblk(op, op.cb16, ns, hrp, hcol, codeb, 16, fmtb, c.opt.sched == 0, cap, hf.data(), hmu.data()));
        }
      }
    }
  }
  CB_CUDA(cudaStreamSynchronize(c.stream));
  cudaFree(d_e);
  cudaFree(d_sp);
  CB_CUDA(cudaGetLastError());
  return 0;
}

void free_spin_op(SpinOp &op) {
  dev_free(op.map); dev_free(op.lin_lo); dev_free(op.lin_hi); dev_free(op.f); dev_free(op.terms);
  dev_free(op.rowptr); dev_free(op.col); dev_free(op.val); dev_free(op.ell_col); dev_free(op.ell_val);
  dev_free(op.rowlen); dev_free(op.coef);
  dev_free(op.rr.blocks); dev_free(op.rr.wbase); dev_free(op.rr.thdr); dev_free(op.rr.win); dev_free(op.rr.woff);
  for (ColBlk *cb : {&op.cb8, &op.cb16}) {
    dev_free(cb->blk); dev_free(cb->tbase); dev_free(cb->qbase); dev_free(cb->meta); dev_free(cb->words); dev_free(cb->toff); dev_free(cb->woff);
  }
  for (Sched *sc : {&op.sc8, &op.sc16, &op.sc16x2}) { dev_free(sc->tbase); dev_free(sc->qbase); dev_free(sc->meta); dev_free(sc->words); }
  op = SpinOp();
}

int ensure_stage(int64_t n) {
  Ctx &c = ctx();
  if (c.stage_n >= n) return 0;
  dev_free(c.stage_v); dev_free(c.stage_hv);
  CB_CHECK(dev_alloc(&c.stage_v, n));
  CB_CHECK(dev_alloc(&c.stage_hv, n));
  c.stage_n = n;
  return 0;
}

static int sector_np(int32_t isector, int32_t *nup, int32_t *ndw) {
  Ctx &c = ctx();
  if (!c.have_model) return fail("no model set (call cdmft_b200_set_model)");
  int nsec = (c.ns + 1) * (c.ns + 1);
  if (isector < 1 || isector > nsec) return fail("isector %d out of range [1,%d]", isector, nsec);
  *ndw = (isector - 1) % (c.ns + 1);  // get_Ndw / get_Nup, ED_SETUP.f90:476-500
  *nup = (isector - 1) / (c.ns + 1);
  return 0;
}

}  // namespace cb

using namespace cb;

extern "C" {

int cdmft_b200_schedule_host(int64_t n, const int32_t *rowptr, const int32_t *col, const uint8_t *code, int32_t g,
                             int32_t natural, int32_t nwarps, int32_t *ntask, int64_t *nquads, int32_t *tbase,
                             int32_t *qbase, uint32_t *meta, uint32_t *words) {
  if (g != 8 && g != 16) return fail("schedule_host: g must be 8 or 16");
  if (n <= 0 || n >= (1 << 25)) return fail("schedule_host: n out of range");
  if (nwarps < 1 || nwarps > 32) return fail("schedule_host: nwarps must be in [1,32]");
  std::vector<int32_t> rp(rowptr, rowptr + n + 1), cl(col, col + rowptr[n]);
  std::vector<uint8_t> cd(code, code + rowptr[n]);
  SchedHost sh;
  build_schedule_host(n, rp, cl, cd, g, 0, natural != 0, nwarps, nullptr, nullptr, sh);
  *ntask = sh.ntask;
  *nquads = sh.nquads;
  if (words) {
    std::copy(sh.tbase.begin(), sh.tbase.end(), tbase);
    std::copy(sh.qbase.begin(), sh.qbase.end(), qbase);
    std::copy(sh.meta.begin(), sh.meta.end(), meta);
    std::copy(sh.words.begin(), sh.words.begin() + (size_t)sh.nquads * 32 * 4, words);
  }
  return 0;
}

int cdmft_b200_colblk_host(int32_t ns, int32_t npart, const int32_t *rowptr, const int32_t *col, const uint8_t *code, int32_t g,
                            int32_t natural, int64_t cap_rows, int64_t *sizes, int32_t *blk, int32_t *tbase, int32_t *qbase,
                            uint32_t *meta, uint32_t *words, uint32_t *toff, uint32_t *woff) {
  if (g != 8 && g != 16) return fail("colblk_host: g must be 8 or 16");
  if (ns < 1 || ns > 30 || npart < 0 || npart > ns) return fail("colblk_host: bad ns / npart");
  const int64_t n = binom64(ns, npart);
  if (n <= 0 || n >= (1 << 24)) return fail("colblk_host: sector too large");
  SpinOp op;
  op.n = n;
  op.npart = npart;
  ColBlk dummy;
  ColBlkHost h;
  std::vector<int32_t> rp(rowptr, rowptr + n + 1), cl(col, col + rowptr[n]);
  std::vector<uint8_t> cd(code, code + rowptr[n]);
  CB_CHECK(build_colblk(op, dummy, ns, rp, cl, cd, g, 0, natural != 0, std::max<int64_t>(g, cap_rows), nullptr, nullptr, &h));
  sizes[0] = (int64_t)h.blk.size(); sizes[1] = (int64_t)h.tbase.size(); sizes[2] = (int64_t)h.meta.size();
  sizes[3] = (int64_t)h.words.size(); sizes[4] = (int64_t)h.toff.size(); sizes[5] = (int64_t)h.woff.size();
  sizes[6] = h.nwarps; sizes[7] = h.max_rows;
  if (blk) {
    memcpy(blk, h.blk.data(), h.blk.size() * sizeof(int4));
    std::copy(h.tbase.begin(), h.tbase.end(), tbase);
    std::copy(h.qbase.begin(), h.qbase.end(), qbase);
    std::copy(h.meta.begin(), h.meta.end(), meta);
    std::copy(h.words.begin(), h.words.end(), words);
    memcpy(toff, h.toff.data(), h.toff.size() * sizeof(uint2));
    std::copy(h.woff.begin(), h.woff.end(), woff);
  }
  return 0;
}

int cdmft_b200_get_sector_dims(int32_t isector, int64_t *dimup, int64_t *dimdw, int64_t *dim) {
  int32_t nup, ndw;
  CB_CHECK(sector_np(isector, &nup, &ndw));
  int64_t du = binom64(ctx().ns, nup), dd = binom64(ctx().ns, ndw);
  if (dimup) *dimup = du;
  if (dimdw) *dimdw = dd;
  if (dim) *dim = du * dd;
  return 0;
}

int cdmft_b200_vecdim_hv_sector(int32_t isector, int64_t *vecdim) {
  CB_REQUIRE_INIT();
  Ctx &c = ctx();
  int64_t du, dd, d;
  CB_CHECK(cdmft_b200_get_sector_dims(isector, &du, &dd, &d));
  if (!c.spmd) { *vecdim = d; return 0; }
  int p_eff = (int)std::min<int64_t>(c.nranks, dd);
  // ranks outside the shrunk communicator hold nothing (ED_HAMILTONIAN.f90:62-90)
  *vecdim = c.rank < p_eff ? du * split_of(dd, p_eff, c.rank).q : 0;
  return 0;
}

int cdmft_b200_build_hv_sector(int32_t isector, int32_t mode, int64_t *nloc) {
  CB_REQUIRE_INIT();
  Ctx &c = ctx();
  if (c.hstatus) return fail("build_hv_sector: sector %d is still active (sp_init_matrix: already allocated)", c.hsector);
  if (mode != CDMFT_B200_SPARSE && mode != CDMFT_B200_DIRECT) return fail("build_hv_sector: bad mode %d", mode);
  int32_t nup, ndw;
  CB_CHECK(sector_np(isector, &nup, &ndw));
  CB_CUDA(cudaSetDevice(c.device));
  c.hsector = isector;
  c.mode = mode;
  // ed_sparse_H = F (DIRECT): the reference recomputes every hop instead of storing spH0d / spH0nd / spH0ups / spH0dws.
  // Nothing O(Dim) is ever stored here in either mode; the per-spin operator tables (O(Dim_sigma), < 1 MB at Ns = 16) are
  // built in DIRECT mode too unless option direct_tables = 0 asks for the matrix-free (on-the-fly bit-hopping) kernels.
  c.tables = mode == CDMFT_B200_SPARSE || c.opt.direct_tables != 0;
  CB_CHECK(build_spin_op(c.up, nup, c.terms_up, c.e_up, c.const0, c.tables));
  CB_CHECK(build_spin_op(c.dw, ndw, c.terms_dw, c.e_dw, 0.0, c.tables));
  c.dimup = c.up.n; c.dimdw = c.dw.n; c.dim = c.dimup * c.dimdw;
  c.p_eff = (int)std::min<int64_t>(c.nranks, c.dimdw);
  c.rk.clear();
  const bool sharded = c.spmd || c.sim || c.opt.force_sharded;
  if (c.spmd) {
    if (c.rank < c.p_eff) { RankState r; r.rank = c.rank; c.rk.push_back(r); }
  } else {
    for (int r = 0; r < c.p_eff; r++) { RankState s; s.rank = r; c.rk.push_back(s); }
  }
  int64_t total = 0;
  for (auto &r : c.rk) {
    r.dw = split_of(c.dimdw, c.p_eff, r.rank);
    r.up = split_of(c.dimup, c.p_eff, r.rank);
    r.nloc = c.dimup * r.dw.q;
    total += r.nloc;
    if (sharded) {
      CB_CHECK(dev_alloc(&r.vt, c.dimdw * r.up.q));
      CB_CHECK(dev_alloc(&r.hvt, c.dimdw * r.up.q));
      if (c.spmd && c.p_eff > 1) {
        CB_CHECK(dev_alloc(&r.sendbuf, std::max(r.nloc, c.dimdw * r.up.q)));
        CB_CHECK(dev_alloc(&r.recvbuf, std::max(r.nloc, c.dimdw * r.up.q)));
      }
    }
  }
  c.hstatus = true;
  if (nloc) *nloc = total;
  return 0;
}

int cdmft_b200_delete_hv_sector(void) {
  Ctx &c = ctx();
  if (!c.inited) return 0;
  if (c.stream) cudaStreamSynchronize(c.stream);
  if (c.comm_stream) cudaStreamSynchronize(c.comm_stream);
  if (c.ipc_ready) {  // nobody may still be storing into the windows we are about to free
    nccl_barrier();
    cudaStreamSynchronize(c.stream);
    ipc_close_peers();
    nccl_barrier();
    cudaStreamSynchronize(c.stream);
  }
  lz_free_slots();  // Krylov vectors kept by the ground-state driver
  if (c.kin_built) { free_spin_op(c.kin_up); free_spin_op(c.kin_dw); c.kin_built = false; }
  free_spin_op(c.up);
  free_spin_op(c.dw);
  for (auto &r : c.rk) { dev_free(r.vt); dev_free(r.hvt); dev_free(r.sendbuf); dev_free(r.recvbuf); }
  c.rk.clear();
  dev_free(c.stage_v); dev_free(c.stage_hv); c.stage_n = 0;
  dev_free(c.vfull);
  for (auto &k : c.kv) dev_free(k);
  c.kv_n = 0;
  c.hsector = 0; c.hstatus = false;
  c.dim = c.dimup = c.dimdw = 0;
  return 0;
}

int cdmft_b200_active_ranks(int32_t *p) {
  if (!ctx().hstatus) return fail("active_ranks: no active sector");
  *p = ctx().p_eff;
  return 0;
}

int cdmft_b200_get_sector_map(int32_t which, int32_t *map) {
  Ctx &c = ctx();
  if (!c.hstatus) return fail("get_sector_map: no active sector");
  SpinOp &op = which == 1 ? c.up : c.dw;
  CB_CUDA(cudaMemcpy(map, op.map, op.n * 4, cudaMemcpyDeviceToHost));
  return 0;
}

int cdmft_b200_get_csr_nnz(int32_t which, int64_t *nnz) {
  Ctx &c = ctx();
  if (!c.hstatus) return fail("get_csr_nnz: no active sector");
  if (c.mode != CDMFT_B200_SPARSE) return fail("get_csr_nnz: sector was built in DIRECT mode (no stored matrix)");
  *nnz = (which == 1 ? c.up : c.dw).nnz;
  return 0;
}

int cdmft_b200_get_csr(int32_t which, int64_t *rowptr, int32_t *col, double *val) {
  Ctx &c = ctx();
  if (!c.hstatus) return fail("get_csr: no active sector");
  if (c.mode != CDMFT_B200_SPARSE) return fail("get_csr: sector was built in DIRECT mode (no stored matrix)");
  SpinOp &op = which == 1 ? c.up : c.dw;
  std::vector<int32_t> rp(op.n + 1);
  CB_CUDA(cudaMemcpy(rp.data(), op.rowptr, (op.n + 1) * 4, cudaMemcpyDeviceToHost));
  for (int64_t i = 0; i <= op.n; i++) rowptr[i] = rp[i];
  CB_CUDA(cudaMemcpy(col, op.col, op.nnz * 4, cudaMemcpyDeviceToHost));
  for (int64_t k = 0; k < op.nnz; k++) col[k] += 1;  // 1-based like the reference
  CB_CUDA(cudaMemcpy(val, op.val, op.nnz * 16, cudaMemcpyDeviceToHost));
  return 0;
}

int cdmft_b200_get_sparse_map(int32_t which, int64_t *rowptr, int32_t *bath_state, int32_t *sector_indx) {
  Ctx &c = ctx();
  if (!c.hstatus) return fail("get_sparse_map: no active sector");
  SpinOp &op = which == 1 ? c.up : c.dw;
  const int64_t nst = (int64_t)1 << c.nimp;
  unsigned long long *d_cnt = nullptr;
  CB_CHECK(dev_alloc(&d_cnt, nst));
  CB_CUDA(cudaMemsetAsync(d_cnt, 0, nst * 8, c.stream));
  LAUNCH_1D(k_spmap_count, op.n, op.n, op.map, c.nimp, d_cnt);
  std::vector<unsigned long long> cnt(nst);
  CB_CUDA(cudaMemcpyAsync(cnt.data(), d_cnt, nst * 8, cudaMemcpyDeviceToHost, c.stream));
  CB_CUDA(cudaStreamSynchronize(c.stream));
  std::vector<int64_t> rp(nst + 1, 0);
  for (int64_t k = 0; k < nst; k++) rp[k + 1] = rp[k] + (int64_t)cnt[k];
  for (int64_t k = 0; k <= nst; k++) rowptr[k] = rp[k];
  if (bath_state && sector_indx) {
    std::vector<int32_t> binom(31 * 31, 0);
    for (int n = 0; n < 31; n++)
      for (int k = 0; k < 31; k++) binom[n * 31 + k] = (int32_t)std::min<int64_t>(binom64(n, k), INT32_MAX);
    int64_t *d_rp = nullptr;
    int32_t *d_bn = nullptr, *d_bs = nullptr, *d_si = nullptr;
    CB_CHECK(dev_alloc(&d_rp, nst + 1));
    CB_CHECK(dev_alloc(&d_bn, 31 * 31));
    CB_CHECK(dev_alloc(&d_bs, op.n));
    CB_CHECK(dev_alloc(&d_si, op.n));
    CB_CUDA(cudaMemcpyAsync(d_rp, rp.data(), (nst + 1) * 8, cudaMemcpyHostToDevice, c.stream));
    CB_CUDA(cudaMemcpyAsync(d_bn, binom.data(), binom.size() * 4, cudaMemcpyHostToDevice, c.stream));
    LAUNCH_1D(k_spmap_fill, op.n, op.n, op.map, c.ns, c.nimp, op.npart, d_rp, d_bn, d_bs, d_si);
    CB_CUDA(cudaMemcpyAsync(bath_state, d_bs, op.n * 4, cudaMemcpyDeviceToHost, c.stream));
    CB_CUDA(cudaMemcpyAsync(sector_indx, d_si, op.n * 4, cudaMemcpyDeviceToHost, c.stream));
    CB_CUDA(cudaStreamSynchronize(c.stream));
    cudaFree(d_rp); cudaFree(d_bn); cudaFree(d_bs); cudaFree(d_si);
  }
  cudaFree(d_cnt);
  return 0;
}

}  // extern "C"
