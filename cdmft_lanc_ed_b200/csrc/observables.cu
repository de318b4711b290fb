// observables.cu -- impurity-configuration weights of a sector vector, the device part of the reference's
// lanc_observables (ED_OBSERVABLES.f90:94-236).
//
// The reference loops on the master over every basis state i, decodes the impurity occupations of the up and dw
// Fock states and accumulates gs_weight = |vec(i)|^2 into dens / docc / magz / s2tot / sz2 / n2 (:120-192).  Every
// one of those observables depends on the state only through the impurity bits (mu, md) of its up and dw parts, so
// the O(Dim) work is ONE reduction: W[mu + md*2^Nimp] = sum over the bath configurations of |vec|^2.  The host
// mirror (ed_hamiltonian.lanc_observables / the Fortran shim) evaluates the reference's formulas on the
// 4^Nimp-entry table.  Sharded vectors: every rank reduces its own columns and the tables are all-reduced (the
// reference broadcasts the master's result).
//
// Deterministic: rows are grouped by mu once (stable order), one warp sums one (column, mu) segment with a fixed
// lane assignment and shuffle tree, and the columns are combined in ascending order on the host.
#include <algorithm>
#include <vector>

#include "ctx.h"

namespace cb {

__global__ void __launch_bounds__(256) k_imp_weights(int64_t n, int64_t ncols, int nmu, const double2 *__restrict__ v,
                                                      const int32_t *__restrict__ perm, const int32_t *__restrict__ seg,
                                                      double *__restrict__ colhist) {
  const int64_t task = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (task >= ncols * nmu) return;
  const int64_t c = task / nmu;
  const int mu = (int)(task - c * nmu);
  const int lane = threadIdx.x & 31;
  const double2 *vc = v + c * n;
  double s = 0.0;
  for (int k = __ldg(seg + mu) + lane; k < __ldg(seg + mu + 1); k += 32) {
    const double2 x = __ldg(vc + __ldg(perm + k));
    s += x.x * x.x + x.y * x.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if (lane == 0) colhist[task] = s;
}

// ------------------------------------------------------------------------------------
// Cluster-reduced density matrix rho_IMP = Tr_BATH |vec><vec| (density_matrix_impurity, ED_OBSERVABLES.f90:465-575).
// States are ascending integers with the bath in the high bits, so the states sharing a bath configuration form a RUN
// of consecutive sector indices whose impurity parts are all configurations of one particle number, ascending.  For
// a pair of runs (up run of length LU, dw run of length LD) the amplitudes m[a + LU*b] = vec(u0+a, d0+b) are one
// column of a matrix A_class, and rho restricted to the class (ku, kd) is the Gram matrix A A^H summed over all pairs
// of runs of that class -- the reference's loops over (IimpUp,JimpUp,IimpDw,JimpDw) and their shared bath states, reordered.
// One CTA = one 32x32 tile of one class's Gram matrix over one chunk of run pairs; partial tiles are summed in a fixed
// order afterwards (deterministic).
// ------------------------------------------------------------------------------------
struct GramItem {  // one (class, tile row, tile col, chunk)
  int32_t ru0, nu, rd0, nd;  // runs of the class: starts in runs_up[ru0 .. ru0+nu), runs_dw[rd0 .. rd0+nd)
  int32_t LU, LD, ti, tj;
  int64_t p0, p1;            // run pairs [p0, p1) of the class, p = iu + nu * id
  int64_t out;               // first element of this item's 32x32 partial tile
};
constexpr int kGramBatch = 4;
__global__ void __launch_bounds__(256) k_cluster_gram(const double2 *__restrict__ v, int64_t dimup, const int32_t *__restrict__ runs_up,
                                                       const int32_t *__restrict__ runs_dw, const GramItem *__restrict__ items,
                                                       double2 *__restrict__ partial) {
  const GramItem it = items[blockIdx.x];
  __shared__ double2 mi[kGramBatch][32], mj[kGramBatch][32];
  const int x = threadIdx.x & 31, y = threadIdx.x >> 5;  // entries (I = ti*32 + x, J = tj*32 + y + 8k)
  const int L = it.LU * it.LD;
  double2 acc[4];
#pragma unroll
  for (int k = 0; k < 4; k++) acc[k] = make_double2(0.0, 0.0);
  for (int64_t p = it.p0; p < it.p1; p += kGramBatch) {
    // stage the tile rows / columns of kGramBatch amplitude columns (zero past the class size or the chunk end)
    if (threadIdx.x < kGramBatch * 64) {
      const int bq = threadIdx.x >> 6, which = (threadIdx.x >> 5) & 1, e = threadIdx.x & 31;
      const int64_t pp = p + bq;
      const int idx = (which ? it.tj : it.ti) * 32 + e;
      double2 val = make_double2(0.0, 0.0);
      if (pp < it.p1 && idx < L) {
        const int64_t u0 = __ldg(runs_up + it.ru0 + (int32_t)(pp % it.nu)), d0 = __ldg(runs_dw + it.rd0 + (int32_t)(pp / it.nu));
        val = __ldg(v + (u0 + idx % it.LU) + (d0 + idx / it.LU) * dimup);
      }
      if (which) mj[bq][e] = val; else mi[bq][e] = val;
    }
    __syncthreads();
#pragma unroll
    for (int bq = 0; bq < kGramBatch; bq++) {
      const double2 a = mi[bq][x];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const double2 b = mj[bq][y + 8 * k];  // acc += a * conj(b)
        acc[k].x = fma(a.x, b.x, fma(a.y, b.y, acc[k].x));
        acc[k].y = fma(a.y, b.x, fma(-a.x, b.y, acc[k].y));
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < 4; k++) partial[it.out + x + 32 * (y + 8 * k)] = acc[k];
}

}  // namespace cb

using namespace cb;

extern "C" int cdmft_b200_density_matrices(int64_t nloc, const void *vec, double peso, double *cdm, double *spdm) {
  CB_REQUIRE_INIT();
  Ctx &c = ctx();
  if (!c.hstatus) return fail("density_matrices: Hsector NOT set (call build_hv_sector for the sector of the vector)");
  int64_t expect = 0;
  for (auto &r : c.rk) expect += r.nloc;
  if (nloc != expect) return fail("density_matrices: Nloc=%lld != local dimension %lld", (long long)nloc, (long long)expect);
  if (c.nimp > 6) return fail("density_matrices: Nimp=%d > 6 (16^Nimp entries)", c.nimp);
  if (c.spmd && c.nranks > 1 && c.p_eff != c.nranks) return fail("density_matrices: sector too small for %d ranks (run it on one)", c.nranks);
  const int nimp = c.nimp, NI = 1 << nimp, nlat = c.m.nlat, norb = c.m.norb, nspin = c.m.nspin;
  const double2 *dv = (const double2 *)vec;
  double2 *d_tmp = nullptr;
  std::vector<void *> to_free;
  auto cleanup = [&]() { for (void *p : to_free) cudaFree(p); };
  if (nloc > 0 && !is_device_ptr(vec)) {
    if (dev_alloc(&d_tmp, nloc)) return 1;
    to_free.push_back(d_tmp);
    cudaMemcpyAsync(d_tmp, vec, (size_t)nloc * 16, cudaMemcpyHostToDevice, c.stream);
    dv = d_tmp;
  }
  int rc = 0;
  // ---- single-particle density matrix <C^+_a C_b> (:600-676): diagonal from the weight table, off-diagonal one
  // matrix-free product per pair a < b and spin block (Nspin = 1: the reference fills the spin-up block only)
  if (spdm) {
    std::vector<double> W((size_t)NI * NI);
    if ((rc = cdmft_b200_imp_weights(nloc, dv, W.data()))) { cleanup(); return rc; }
    auto spix = [&](int a, int b, int s) {  // (ilat,jlat,ispin,ispin,iorb,jorb) column-major, complex interleaved
      const int il = a / norb, io = a % norb, jl = b / norb, jo = b % norb;
      return 2 * (size_t)(il + nlat * (jl + nlat * (s + nspin * (s + nspin * (io + norb * jo)))));
    };
    for (int s = 0; s < nspin; s++)
      for (int a = 0; a < nimp; a++) {
        double d = 0.0;
        for (int mu = 0; mu < NI; mu++)
          for (int md = 0; md < NI; md++)
            if (((s == 0 ? mu : md) >> a) & 1) d += W[(size_t)mu + (size_t)md * NI];
        spdm[spix(a, a, s)] += peso * d;
        for (int b = a + 1; b < nimp; b++) {
          std::vector<Term> t{Term{a, b, 1.0, 0.0}}, none;
          double z[2];
          if ((rc = expect_terms(s == 0 ? t : none, s == 0 ? none : t, dv, z))) { cleanup(); return rc; }
          spdm[spix(a, b, s)] += peso * z[0];
          spdm[spix(a, b, s) + 1] += peso * z[1];
          spdm[spix(b, a, s)] += peso * z[0];  // <c^+_b c_a> = conjg <c^+_a c_b>
          spdm[spix(b, a, s) + 1] -= peso * z[1];
        }
      }
  }
  if (!cdm) { cleanup(); return 0; }
  // ---- cluster density matrix: needs the whole vector on this process
  const double2 *vfull = dv;
  if ((rc = allgather_full(dv, &vfull))) { cleanup(); return rc; }
  std::vector<int32_t> mapu(c.dimup), mapd(c.dimdw);
  if (cudaMemcpyAsync(mapu.data(), c.up.map, c.dimup * 4, cudaMemcpyDeviceToHost, c.stream) != cudaSuccess ||
      cudaMemcpyAsync(mapd.data(), c.dw.map, c.dimdw * 4, cudaMemcpyDeviceToHost, c.stream) != cudaSuccess ||
      cudaStreamSynchronize(c.stream) != cudaSuccess) {
    cleanup();
    return fail("density_matrices: CUDA error %s", cudaGetErrorString(cudaGetLastError()));
  }
  // runs per impurity particle number k: starts of the runs whose impurity part has k particles; impurity configurations of k particles
  std::vector<std::vector<int32_t>> cfg(nimp + 1);
  for (int m = 0; m < NI; m++) cfg[__builtin_popcount(m)].push_back(m);
  auto runs_of = [&](const std::vector<int32_t> &map, std::vector<std::vector<int32_t>> &runs) {
    runs.assign(nimp + 1, {});
    for (size_t i = 0; i < map.size();) {
      size_t j = i;
      while (j < map.size() && (map[j] >> nimp) == (map[i] >> nimp)) j++;
      runs[__builtin_popcount(map[i] & (NI - 1))].push_back((int32_t)i);  // run length = cfg[k].size() by construction
      i = j;
    }
  };
  std::vector<std::vector<int32_t>> ru, rd;
  runs_of(mapu, ru);
  runs_of(mapd, rd);
  std::vector<int32_t> flat_u, flat_d, baseu(nimp + 1), based(nimp + 1);
  for (int k = 0; k <= nimp; k++) { baseu[k] = (int32_t)flat_u.size(); flat_u.insert(flat_u.end(), ru[k].begin(), ru[k].end()); }
  for (int k = 0; k <= nimp; k++) { based[k] = (int32_t)flat_d.size(); flat_d.insert(flat_d.end(), rd[k].begin(), rd[k].end()); }
  std::vector<GramItem> items;
  struct TileRef { int ku, kd, ti, tj; size_t first_item, nchunk; };
  std::vector<TileRef> tiles;
  int64_t nout = 0;
  for (int ku = 0; ku <= nimp; ku++)
    for (int kd = 0; kd <= nimp; kd++) {
      const int64_t nu = (int64_t)ru[ku].size(), nd = (int64_t)rd[kd].size(), np = nu * nd;
      if (np == 0) continue;
      const int LU = (int)cfg[ku].size(), LD = (int)cfg[kd].size(), L = LU * LD, nt = (L + 31) / 32;
      const int64_t nchunk = std::max<int64_t>(1, std::min<int64_t>(64, np / 512));
      for (int ti = 0; ti < nt; ti++)
        for (int tj = 0; tj < nt; tj++) {
          tiles.push_back(TileRef{ku, kd, ti, tj, items.size(), (size_t)nchunk});
          for (int64_t ch = 0; ch < nchunk; ch++) {
            GramItem g{};
            g.ru0 = baseu[ku]; g.nu = (int32_t)nu; g.rd0 = based[kd]; g.nd = (int32_t)nd;
            g.LU = LU; g.LD = LD; g.ti = ti; g.tj = tj;
            g.p0 = np * ch / nchunk; g.p1 = np * (ch + 1) / nchunk;
            g.out = nout;
            nout += 1024;
            items.push_back(g);
          }
        }
    }
  int32_t *d_ru = nullptr, *d_rd = nullptr;
  GramItem *d_items = nullptr;
  double2 *d_part = nullptr;
  std::vector<double2> part((size_t)nout);
  do {
    if ((rc = dev_alloc(&d_ru, (int64_t)flat_u.size()))) break;
    to_free.push_back(d_ru);
    if ((rc = dev_alloc(&d_rd, (int64_t)flat_d.size()))) break;
    to_free.push_back(d_rd);
    if ((rc = dev_alloc(&d_items, (int64_t)items.size()))) break;
    to_free.push_back(d_items);
    if ((rc = dev_alloc(&d_part, nout))) break;
    to_free.push_back(d_part);
    cudaMemcpyAsync(d_ru, flat_u.data(), flat_u.size() * 4, cudaMemcpyHostToDevice, c.stream);
    cudaMemcpyAsync(d_rd, flat_d.data(), flat_d.size() * 4, cudaMemcpyHostToDevice, c.stream);
    cudaMemcpyAsync(d_items, items.data(), items.size() * sizeof(GramItem), cudaMemcpyHostToDevice, c.stream);
    if (!items.empty()) {
      k_cluster_gram<<<(unsigned)items.size(), 256, 0, c.stream>>>(vfull, c.dimup, d_ru, d_rd, d_items, d_part);
      c.launches++;
    }
    if (cudaMemcpyAsync(part.data(), d_part, (size_t)nout * 16, cudaMemcpyDeviceToHost, c.stream) != cudaSuccess ||
        cudaStreamSynchronize(c.stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) {
      rc = fail("density_matrices: CUDA error %s", cudaGetErrorString(cudaGetLastError()));
      break;
    }
  } while (0);
  cleanup();
  if (rc) return rc;
  // chunks summed in ascending order, entries scattered to the reference's labels io = IimpUp + 2^Nimp * IimpDw
  const size_t N2 = (size_t)NI * NI;
  for (const TileRef &t : tiles) {
    const int LU = (int)cfg[t.ku].size(), LD = (int)cfg[t.kd].size(), L = LU * LD;
    for (int y = 0; y < 32; y++)
      for (int x = 0; x < 32; x++) {
        const int I = t.ti * 32 + x, J = t.tj * 32 + y;
        if (I >= L || J >= L) continue;
        double re = 0.0, im = 0.0;
        for (size_t ch = 0; ch < t.nchunk; ch++) {
          const double2 p = part[(size_t)items[t.first_item + ch].out + x + 32 * y];
          re += p.x;
          im += p.y;
        }
        const size_t io = (size_t)cfg[t.ku][I % LU] + (size_t)NI * cfg[t.kd][I / LU];
        const size_t jo = (size_t)cfg[t.ku][J % LU] + (size_t)NI * cfg[t.kd][J / LU];
        cdm[2 * (io + jo * N2)] += peso * re;
        cdm[2 * (io + jo * N2) + 1] += peso * im;
      }
  }
  return 0;
}

extern "C" int cdmft_b200_imp_weights(int64_t nloc, const void *vec, double *w) {
  CB_REQUIRE_INIT();
  Ctx &c = ctx();
  if (!c.hstatus) return fail("imp_weights: Hsector NOT set (call build_hv_sector for the sector of the vector)");
  int64_t expect = 0;
  for (auto &r : c.rk) expect += r.nloc;
  if (nloc != expect) return fail("imp_weights: Nloc=%lld != local dimension %lld", (long long)nloc, (long long)expect);
  if (c.nimp > 8) return fail("imp_weights: Nimp=%d > 8 (table of 4^Nimp entries)", c.nimp);
  const int nmu = 1 << c.nimp;
  const int64_t n = c.dimup;
  std::vector<double> W((size_t)nmu * nmu, 0.0);
  // rows grouped by their impurity configuration (stable counting sort on the host: DimUp is small)
  std::vector<int32_t> mapu(n), mapd(c.dimdw), perm(n), seg(nmu + 1, 0);
  CB_CUDA(cudaMemcpyAsync(mapu.data(), c.up.map, n * 4, cudaMemcpyDeviceToHost, c.stream));
  CB_CUDA(cudaMemcpyAsync(mapd.data(), c.dw.map, c.dimdw * 4, cudaMemcpyDeviceToHost, c.stream));
  CB_CUDA(cudaStreamSynchronize(c.stream));
  for (int64_t i = 0; i < n; i++) seg[(mapu[i] & (nmu - 1)) + 1]++;
  for (int m = 0; m < nmu; m++) seg[m + 1] += seg[m];
  {
    std::vector<int32_t> fill(seg.begin(), seg.end() - 1);
    for (int64_t i = 0; i < n; i++) perm[fill[mapu[i] & (nmu - 1)]++] = (int32_t)i;
  }
  int32_t *d_perm = nullptr, *d_seg = nullptr;
  double *d_hist = nullptr, *d_w = nullptr;
  const double2 *d_vec = (const double2 *)vec;
  double2 *d_tmp = nullptr;
  int64_t ncols_loc = 0;
  for (auto &r : c.rk) ncols_loc += r.dw.q;
  int rc = 0;
  do {
    if ((rc = dev_alloc(&d_perm, n))) break;
    if ((rc = dev_alloc(&d_seg, (int64_t)nmu + 1))) break;
    if ((rc = dev_alloc(&d_hist, std::max<int64_t>(1, ncols_loc * nmu)))) break;
    if ((rc = dev_alloc(&d_w, (int64_t)nmu * nmu))) break;
    cudaMemcpyAsync(d_perm, perm.data(), n * 4, cudaMemcpyHostToDevice, c.stream);
    cudaMemcpyAsync(d_seg, seg.data(), (nmu + 1) * 4, cudaMemcpyHostToDevice, c.stream);
    if (!is_device_ptr(vec) && nloc > 0) {
      if ((rc = dev_alloc(&d_tmp, nloc))) break;
      cudaMemcpyAsync(d_tmp, vec, (size_t)nloc * 16, cudaMemcpyHostToDevice, c.stream);
      d_vec = d_tmp;
    }
    if (ncols_loc > 0) {
      // the local shards are contiguous column blocks (one in SPMD mode, all P in sim mode): one launch
      const int64_t ntask = ncols_loc * nmu;
      k_imp_weights<<<(unsigned)((ntask + 7) / 8), 256, 0, c.stream>>>(n, ncols_loc, nmu, d_vec, d_perm, d_seg, d_hist);
      c.launches++;
      std::vector<double> hist((size_t)ntask);
      if (cudaMemcpyAsync(hist.data(), d_hist, (size_t)ntask * 8, cudaMemcpyDeviceToHost, c.stream) != cudaSuccess ||
          cudaStreamSynchronize(c.stream) != cudaSuccess) { rc = fail("imp_weights: CUDA error %s", cudaGetErrorString(cudaGetLastError())); break; }
      int64_t lc = 0;
      for (auto &r : c.rk)
        for (int64_t q = 0; q < r.dw.q; q++, lc++) {
          const int md = mapd[r.dw.off + q] & (nmu - 1);
          for (int m = 0; m < nmu; m++) W[(size_t)m + (size_t)md * nmu] += hist[(size_t)lc * nmu + m];
        }
    }
    if (c.spmd && c.nranks > 1) {  // every rank ends up with the full table (the reference: Bcast from the master)
      cudaMemcpyAsync(d_w, W.data(), W.size() * 8, cudaMemcpyHostToDevice, c.stream);
      if ((rc = nccl_allreduce_sum(d_w, (int)W.size()))) break;
      cudaMemcpyAsync(W.data(), d_w, W.size() * 8, cudaMemcpyDeviceToHost, c.stream);
      if (cudaStreamSynchronize(c.stream) != cudaSuccess) { rc = fail("imp_weights: CUDA error after all-reduce"); break; }
    }
  } while (0);
  cudaFree(d_perm); cudaFree(d_seg); cudaFree(d_hist); cudaFree(d_w); cudaFree(d_tmp);
  if (rc) return rc;
  std::copy(W.begin(), W.end(), w);
  return 0;
}
