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

}  // namespace cb

using namespace cb;

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
