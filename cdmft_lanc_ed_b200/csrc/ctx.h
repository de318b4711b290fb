// ctx.h -- internal state of libcdmft_b200 (not part of the ABI).
// The reference keeps the active sector in module globals (ED_HAMILTONIAN_COMMON.f90:11-20,
// ED_VARS_GLOBAL.f90:142-146,286-306); this is the same thing as one C++ singleton.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <map>
#include <string>
#include <vector>

#include "../../include/cdmft_b200.h"

namespace cb {

struct Term {  // h(a,b) c^+_a c_b, a != b, 0-based bit positions inside one spin's string
  int32_t a, b;
  double re, im;
};

// Conflict-free schedule of a per-spin operator for the column-resident kernels (k_colres, hxv.cu):
// rows are taken in groups of G consecutive rows (G lanes = one shared-memory phase: 8 lanes x 16 B for
// complex vectors, 16 lanes x 8 B for real ones); the entries of a group are edge-coloured (rows x banks,
// bank = source row mod G) so that at every step the G lanes read G different banks.  Groups are sorted
// by their number of steps and dealt 32/G to a warp task.
struct Sched {
  int32_t G = 0, ntask = 0;
  int32_t fmt = 0;              // word format: 0 general (table ids), 1 fast 32-bit, 2 fast 16-bit, 3 fast4 (sector.cu)
  int32_t nwarps = 0;           // warps per CTA the streams were dealt for (= launch configuration)
  int64_t nquads = 0;           // quads (4 steps) over all tasks
  int32_t *tbase = nullptr;     // [nwarps+1] first task of each warp (tasks are numbered warp-major) (device)
  int32_t *qbase = nullptr;     // [nwarps+1] first quad of each warp's contiguous word stream
  uint32_t *meta = nullptr;     // [ntask*32] uint4 per task and lane: f_row, row index (-1 none), mu | nquad << 16
  uint32_t *words = nullptr;    // [(nquads+2)*32] uint4 / uint2 per quad and lane; formats in sector.cu
};

// Column-resident kernel for columns that do not fit in shared memory (Ns = 18: 778 KB): the column is cut into
// row blocks = runs of states sharing their top bits (K5: four blocks of <= 12 870 rows).  A CTA owns
// (column, block): the block's rows are staged by TMA, entries whose source lies inside the block follow the
// block's own edge-coloured schedule (relative indices), the others (hops that change the top bits) are
// gathered from global memory / L2.  One Sched-like stream set per block, concatenated.
#ifndef COLBLK_NW
#define COLBLK_NW 32
#endif
constexpr int kColblkWarps = COLBLK_NW;  // (measured: 32 warps with 2 prefetched off-block steps beat 24 / 4 and 20 / 6)
struct ColBlk {
  int32_t nblk = 0, nwarps = 0, fmt = 0, G = 0;
  int64_t max_rows = 0;
  bool even_blocks = false;   // every block starts and ends on an even row (needed for 8-byte elements: 16-byte aligned copies)
  int4 *blk = nullptr;        // [nblk] {first row, rows, first task, first unit}
  int32_t *tbase = nullptr;   // [nblk*(nwarps+1)] first task of each warp, relative to the block's first task
  int32_t *qbase = nullptr;   // [nblk*(nwarps+1)] first unit of each warp's stream, relative to the block's first unit
  uint32_t *meta = nullptr;   // [ntask*32] uint4: f_row, row relative to the block (-1 none), mu | units << 16
  uint32_t *words = nullptr;  // in-block words (formats of Sched, indices relative to the block)
  uint2 *toff = nullptr;      // [ntask] {first off-block step, off-block steps} of each task
  uint32_t *woff = nullptr;   // [steps*32] off-block words: sign<<31 | row<<3 | code3 (fast) or row<<7 | id; see sector.cu
};

// Operator streams of the tile-resident row pass (k_rowtile, rowtile.cu).  Columns (states of this spin) are cut
// into blocks = runs of states sharing their leading bits; a work item is one block x 8 rows of the other spin's
// index, kept in shared memory; per warp task of 4 columns (one per 8-lane group) the in-block entries run against
// the tile and the off-block entries (hops that change the leading bits) against global memory / L2.  Every
// (block, warp) owns contiguous runs of the three streams.
struct RowRes {
  int32_t nblocks = 0, max_block = 0, ntask = 0, nwarps = 0;
  int32_t fmt = 0;              // 0 = coefficient-table ids, 1 = sign / class bits (real H), 2 = sign / class / phase bits
  double in_frac = 0.0;         // share of the entries whose source lies inside the block
  int2 *blocks = nullptr;       // [nblocks] (first column, columns)
  int4 *wbase = nullptr;        // [nblocks*nwarps] {first task, tasks, first in-block unit, first off-block step}
  uint32_t *thdr = nullptr;     // [ntask*4 groups] column (0xFFFF = none) | quads << 16 | off-block steps << 24
  uint4 *win = nullptr;         // [units*4 groups] 8 (16-bit words) or 4 (32-bit words) steps of one lane group; formats in rowtile.cu
  uint4 *woff = nullptr;        // [steps] one off-block word per lane group
};

// Per-spin operator of the active sector: Hs(s)%map + spH0ups(1)/spH0dws(1).
struct SpinOp {
  int32_t npart = 0;        // Nup or Ndw
  int64_t n = 0;            // DimUp or DimDw
  int32_t *map = nullptr;   // [n] ascending Ns-bit integers with popcount npart (device)
  int32_t *lin_lo = nullptr, *lin_hi = nullptr;  // Lin tables: rank(s) = lin_hi[s>>L] + lin_lo[s & (2^L-1)]
  double *f = nullptr;      // [n] spin-local part of the diagonal (device)
  // one-body hop terms (device copy) -- used by the CSR builder and the matrix-free kernels
  Term *terms = nullptr;
  int32_t nterms = 0;
  bool real_h = true;
  // canonical CSR (mode SPARSE): rows ascending, cols ascending 0-based
  int64_t nnz = 0;
  int32_t *rowptr = nullptr;  // [n+1]
  int32_t *col = nullptr;     // [nnz]
  double2 *val = nullptr;     // [nnz]
  // ELL copy, column-major (k*n + i), row lengths in rowlen: coalesced for row-per-thread kernels
  int32_t ell_w = 0;
  int32_t *ell_col = nullptr;
  double2 *ell_val = nullptr;
  int32_t *rowlen = nullptr;
  double2 *coef = nullptr;       // [128] distinct signed coefficients (table decode), coef[0] = 0
  int32_t ncoef = 0;
  // column-resident kernels: schedules for 16-byte (sc8) and 8-byte (sc16) vector elements
  Sched sc8, sc16;
  Sched sc16x2;                  // 8-byte elements, dealt for 32 warps: the two-column kernel k_colres2 (hxv_real.cu)
  ColBlk cb8, cb16;
  RowRes rr;
  bool sc_fast = false;          // real H with <= 2 distinct |coefficients|: sign and class bits instead of table ids
  double sc_mag[4] = {0.0, 0.0, 0.0, 0.0};  // |coefficient| classes of the fast decodes
};

struct Split {  // first (n mod P) ranks get one more (ED_HAMILTONIAN.f90:92-105)
  int64_t q, off;
};
inline Split split_of(int64_t n, int P, int r) {
  int64_t q = n / P, rem = n % P;
  if (r < rem) return {q + 1, (int64_t)r * (q + 1)};
  return {q, (int64_t)r * q + rem};
}

// one rank's share of the active sector (SPMD: exactly one; sim: all P)
struct RankState {
  int rank = 0;
  Split dw{};  // owned columns of v(DimUp, DimDw)      (mpiQdw, offset)
  Split up{};  // owned rows in the transposed layout   (mpiQup, offset)
  int64_t nloc = 0;
  double2 *vt = nullptr, *hvt = nullptr;  // transposed work vectors [DimDw * up.q]
  double2 *sendbuf = nullptr, *recvbuf = nullptr;  // NCCL staging (SPMD only)
};

struct Options {
  // column pass: 6 = column-resident shared-memory kernels (default; falls back to 1 when a column does not fit
  // in shared memory and has no block schedules, or in DIRECT mode), 1 = generic global-gather kernel
  int64_t colpass_variant = 6;
  int64_t sched = 1;            // 1 = conflict-free edge-coloured schedule, 0 = natural CSR order (for comparison)
  int64_t colres_rows = 0;      // > 0: force the block-split column-resident kernel with at most this many rows per block
  // row pass: 1 = generic L2-slab kernel (default: 2.77 ms at K3), 4 = tile-resident shared-memory kernel (3.2 ms at K3:
  // HBM traffic is algorithmic, but the off-block gathers and the register budget keep it latency-bound)
  int64_t rowpass_variant = 1;
  int64_t rowres_cols = 0;      // > 0: cap on the columns of a row-pass block (default: what two tile buffers hold)
  int64_t tma2d = 1;            // row pass tiles by 2-D TMA tensor copies (0: one 128-byte bulk copy per column)
  int64_t force_sharded = 0;    // single rank: run the transpose path anyway (P=1)
  int64_t col_batch = 4;
  int64_t row_slab = 128;       // rows per slab of the row pass (slab x all columns stays in L2)
  int64_t use_ipc = 1;          // SPMD with peer windows (ipc_import): 1 = copy-engine exchange (default), 2 = transposing kernels that
                                // store into peer memory, 0 = NCCL send/recv
  int64_t direct_tables = 1;    // DIRECT mode (ed_sparse_H = F): 1 = per-spin operator tables + the table kernels, 0 = matrix-free kernels
  int64_t colres_pair = 1;      // real Krylov vectors: two columns per pass of the column-resident kernel (k_colres2)
  int64_t xchg_split = 1;       // copy-engine exchange: DMA streams per peer (each copy cut into this many pieces)
  int64_t xchg_chunks = 0;      // copy-engine exchange: chunks of the Hdw pass pipelined against the way back (0 = auto: 4 with one peer, 2 with more)
  int64_t row_rb = 2;           // row chunks per thread in the generic SPARSE row pass (1 = one row per thread)
  int64_t fast4 = 1;            // sign/class/phase decode for purely real-or-imaginary coefficients
  int64_t fuse_dot = 1;         // Krylov drivers: Re<u,Hu> reduced inside the last pass of H x v
  int64_t real_lanczos = 1;     // Krylov drivers keep real vectors when H and the start vector are real
  int64_t overlap = 1;          // SPMD: overlap the transpose of v with the diag+Hup pass
  int64_t lanczos_batch = 4;    // Krylov drivers: steps enqueued between two read-backs of (alfa, beta)
  int64_t lanczos_store = 1;    // ground-state driver: keep the Krylov vectors in HBM as far as they fit (1; n > 1: at most n); 0 = two passes
};

struct Ctx {
  bool inited = false;
  int device = 0;
  int nranks = 1, rank = 0;
  bool spmd = false, sim = false;
  void *nccl_comm = nullptr;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  cudaStream_t comm_stream = nullptr;  // SPMD: pack / all-to-all / unpack of the first transpose
  cudaEvent_t ev_in = nullptr, ev_comm = nullptr;
  int64_t launches = 0;
  int sm_count = 148;
  Options opt;

  // model (host copies)
  bool have_model = false;
  cdmft_b200_model m{};
  std::vector<double> imphloc, hbath, vbath;
  int32_t ns = 0, nimp = 0, nlso = 0;
  bool jhflag = false;
  std::vector<Term> terms_up, terms_dw;
  bool real_h = true;
  // diagonal coefficients (host)
  std::vector<double> e_up, e_dw;   // [Ns]
  std::vector<double> spair;        // [Nimp*Nimp] same-spin pair interaction (a<b)
  std::vector<double> wcross;       // [Nimp*Nimp] up-dw density-density
  double const0 = 0;
  double *cross_tab = nullptr;      // device [Nimp * 2^Nimp]: T[b][mu] = sum_a W[a][b] n_a(mu)

  // active sector
  bool hstatus = false;
  int32_t hsector = 0, mode = CDMFT_B200_SPARSE;
  bool tables = true;  // per-spin operator tables (CSR, schedules) exist for the active sector
  int64_t dim = 0, dimup = 0, dimdw = 0;
  int p_eff = 1;  // min(P, DimDw)
  SpinOp up, dw;
  // impurity-only hopping operators of the active sector (cdmft_b200_imp_kinetic: <E0> of lanc_local_energy); built on
  // first use, swapped into up / dw for one product with kin_only set
  SpinOp kin_up, kin_dw;
  bool kin_built = false, kin_only = false;
  std::map<int, SpinOp *> map_ops;  // map-only SpinOps per particle number (apply_op), valid for the current model
  std::vector<RankState> rk;
  // CUDA-IPC peer windows (SPMD, optional): peers' vt and recvbuf mapped into this process so the
  // transposing kernel stores straight into the destination GPU over NVLink (pack + exchange + unpack
  // in one kernel).  Indexed by rank; entry for this rank = local pointer.
  bool ipc_ready = false;
  std::vector<double2 *> peer_vt, peer_recv;
  // copy-engine exchange (hxv.cu, hxv_sharded_ce): one stream + event per peer, created on first use
  std::vector<cudaStream_t> peer_stream;
  std::vector<cudaEvent_t> peer_done;
  cudaEvent_t ev_pack = nullptr;
  double2 *vfull = nullptr;  // all-gathered vector for the non-local (Jx/Jp) term, SPMD only
  // staging for host-pointer calls
  double2 *stage_v = nullptr, *stage_hv = nullptr;
  int64_t stage_n = 0;
  // Krylov work vectors (allocated lazily)
  double2 *kv[3] = {nullptr, nullptr, nullptr};
  int64_t kv_n = 0;
  // per-kernel-kind CUDA-event timing (option "profile"): 0 column pass, 1 row pass, 2 transpose /
  // pack / unpack, 3 NCCL all-to-all, 4 Krylov vector kernels
  bool profile = false;
  struct ProfRec { int kind; cudaEvent_t a, b; };
  std::vector<ProfRec> prof;
  // fused Lanczos dot: the row pass leaves one partial of Re<v,Hv> per CTA when dot_request is set
  bool dot_request = false, dot_done = false;
  bool dot_final_rowpass = false;  // the accumulating row pass is the last contribution to H x v
  double *dot_partial = nullptr;
  int64_t dot_cap = 0, dot_npartial = 0;
  unsigned int *rt_queue = nullptr;  // work queue of the tile-resident row pass (rowtile.cu)
  double *red = nullptr;       // device reduction scratch
  double *lz_hist = nullptr;   // device history of the Krylov drivers: (alfa_j, beta_{j+1}) per step
  int lz_hist_cap = 0;
  std::vector<void *> lz_slots;  // Krylov-vector slots of the ground-state driver (freed with the sector)
  size_t lz_slot_bytes = 0;
  double *red_host = nullptr;  // pinned
};

Ctx &ctx();
inline bool use_tables() { return ctx().tables; }
void prof_begin(int kind);
void prof_end();
int fail(const char *fmt, ...);
void set_error(const std::string &s);

#define CB_CUDA(expr)                                                                         \
  do {                                                                                        \
    cudaError_t e__ = (expr);                                                                 \
    if (e__ != cudaSuccess)                                                                   \
      return cb::fail("CUDA error %s at %s:%d: %s", #expr, __FILE__, __LINE__, cudaGetErrorString(e__)); \
  } while (0)
#define CB_CHECK(expr)        \
  do {                        \
    int rc__ = (expr);        \
    if (rc__ != 0) return rc__; \
  } while (0)
#define CB_REQUIRE_INIT() \
  if (!cb::ctx().inited) return cb::fail("cdmft_b200: not initialised (call cdmft_b200_init first)")

// ---- internal entry points shared between translation units ----
int build_spin_op(SpinOp &op, int32_t npart, const std::vector<Term> &terms, const std::vector<double> &e,
                  double const_add, bool want_csr);
void free_spin_op(SpinOp &op);
int cached_map_op(int npart, const SpinOp **out);  // lanczos.cu
void free_map_ops();
void lz_free_slots();  // lanczos.cu
int allgather_full(const double2 *v, const double2 **out);  // hxv.cu
// <v| sum_t h_t c^+_a c_b |v> for hop lists of the two spins, through the regular H x v path with the matrix-free kernels (lanczos.cu)
int expect_terms(const std::vector<Term> &tu, const std::vector<Term> &td, const double2 *dv_local, double out[2]);
int scatter_gather_dims(void *vfull, void *vloc, int root, bool scatter, int64_t dimup, int64_t dimdw);  // hxv.cu
int hxv_device(const double2 *v, double2 *hv);  // local shard(s) on device, stream-ordered
int build_rowtile(SpinOp &op, int ns, const std::vector<int32_t> &rowptr, const std::vector<int32_t> &col,
                  const std::vector<uint8_t> &code, int fmt, int64_t cap);  // rowtile.cu
int64_t rowtile_cap();
int nccl_allreduce_sum(double *dev_buf, int n);
int nccl_all_to_all(const double2 *send, double2 *recv, const int64_t *counts_send, const int64_t *offs_send,
                    const int64_t *counts_recv, const int64_t *offs_recv);
int nccl_load();
int nccl_barrier();
void ipc_close_peers();
int ensure_stage(int64_t n);
bool is_device_ptr(const void *p);

template <typename T>
inline int dev_alloc(T **p, int64_t n) {
  *p = nullptr;
  if (n <= 0) n = 1;
  cudaError_t e = cudaMalloc((void **)p, (size_t)n * sizeof(T));
  if (e != cudaSuccess) return fail("cudaMalloc of %lld bytes failed: %s", (long long)(n * sizeof(T)), cudaGetErrorString(e));
  return 0;
}
template <typename T>
inline void dev_free(T *&p) {
  if (p) cudaFree(p);
  p = nullptr;
}

}  // namespace cb
