// runtime.cu -- context, error handling, NCCL bootstrap, model ingestion.
// Restates (in new form) what ED_HAMILTONIAN reads from globals: term lists of
// ED_HAMILTONIAN/sparse/H_up.f90:8-87 / H_dw.f90 and the diagonal of sparse/H_local.f90:1-102.
#include <dlfcn.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <cmath>

#include "ctx.h"

namespace cb {

static Ctx g_ctx;
static thread_local std::string g_err;
Ctx &ctx() { return g_ctx; }
void set_error(const std::string &s) { g_err = s; }
int fail(const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return 1;
}

void prof_begin(int kind) {
  Ctx &c = ctx();
  if (!c.profile) return;
  Ctx::ProfRec r;
  r.kind = kind;
  cudaEventCreate(&r.a);
  cudaEventCreate(&r.b);
  cudaEventRecord(r.a, c.stream);
  c.prof.push_back(r);
}
void prof_end() {
  Ctx &c = ctx();
  if (!c.profile || c.prof.empty()) return;
  cudaEventRecord(c.prof.back().b, c.stream);
}

bool is_device_ptr(const void *p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// ------------------------------------------------------------------------------------
// NCCL through dlopen: binds to the libnccl already loaded by the host program (torch ships
// one) or to the system library; the distributed transpose (vector_transpose_MPI,
// ED_HAMILTONIAN_COMMON.f90:30-94) and the Lanczos dot products use it.
// ------------------------------------------------------------------------------------
typedef struct { char internal[128]; } ncclUniqueId_t;
typedef int (*fn_GetUniqueId)(ncclUniqueId_t *);
typedef int (*fn_CommInitRank)(void **, int, ncclUniqueId_t, int);
typedef int (*fn_CommDestroy)(void *);
typedef int (*fn_Group)(void);
typedef int (*fn_SendRecv)(void *, size_t, int, int, void *, cudaStream_t);
typedef int (*fn_AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef const char *(*fn_ErrStr)(int);
static struct {
  void *h = nullptr;
  fn_GetUniqueId GetUniqueId;
  fn_CommInitRank CommInitRank;
  fn_CommDestroy CommDestroy;
  fn_Group GroupStart, GroupEnd;
  fn_SendRecv Send, Recv;
  fn_AllReduce AllReduce;
  fn_ErrStr GetErrorString;
} nccl;
static const int kNcclDouble = 8, kNcclSum = 0;  // ncclFloat64, ncclSum (nccl.h)

int nccl_load() {
  if (nccl.h) return 0;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *n : names) {
    nccl.h = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (nccl.h) break;
  }
  if (!nccl.h) {
    const char *env = getenv("CDMFT_B200_NCCL_LIB");
    if (env) nccl.h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
  }
  for (const char *n : names) {
    if (nccl.h) break;
    nccl.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
  }
  if (!nccl.h) return fail("cannot load libnccl.so.2 (%s); set CDMFT_B200_NCCL_LIB", dlerror());
#define LOADSYM(field, name)                              \
  *(void **)(&nccl.field) = dlsym(nccl.h, name);          \
  if (!nccl.field) return fail("libnccl: missing symbol %s", name)
  LOADSYM(GetUniqueId, "ncclGetUniqueId");
  LOADSYM(CommInitRank, "ncclCommInitRank");
  LOADSYM(CommDestroy, "ncclCommDestroy");
  LOADSYM(GroupStart, "ncclGroupStart");
  LOADSYM(GroupEnd, "ncclGroupEnd");
  LOADSYM(Send, "ncclSend");
  LOADSYM(Recv, "ncclRecv");
  LOADSYM(AllReduce, "ncclAllReduce");
  LOADSYM(GetErrorString, "ncclGetErrorString");
#undef LOADSYM
  return 0;
}
#define CB_NCCL(expr)                                                                   \
  do {                                                                                  \
    int r__ = (expr);                                                                   \
    if (r__ != 0) return fail("NCCL error at %s:%d: %s", __FILE__, __LINE__, nccl.GetErrorString(r__)); \
  } while (0)

int nccl_allreduce_sum(double *dev_buf, int n) {
  Ctx &c = ctx();
  if (!c.spmd || c.nranks == 1) return 0;
  CB_NCCL(nccl.AllReduce(dev_buf, dev_buf, (size_t)n, kNcclDouble, kNcclSum, c.nccl_comm, c.stream));
  return 0;
}

// cross-rank barrier in stream order (a 1-double all-reduce): brackets the peer-memory transposes
int nccl_barrier() {
  Ctx &c = ctx();
  if (!c.spmd || c.nranks == 1) return 0;
  double *b = c.red + 3000;
  CB_NCCL(nccl.AllReduce(b, b, 1, kNcclDouble, kNcclSum, c.nccl_comm, c.stream));
  return 0;
}

void ipc_close_peers() {
  Ctx &c = ctx();
  if (!c.ipc_ready) return;
  for (int p = 0; p < (int)c.peer_vt.size(); p++) {
    if (p == c.rank) continue;
    if (c.peer_vt[p]) cudaIpcCloseMemHandle(c.peer_vt[p]);
    if (c.peer_recv[p]) cudaIpcCloseMemHandle(c.peer_recv[p]);
  }
  c.peer_vt.clear();
  c.peer_recv.clear();
  c.ipc_ready = false;
}

// all-to-all of complex(8) blocks = the MPI_AllToAllV of vector_transpose_MPI, one grouped call
int nccl_all_to_all(const double2 *send, double2 *recv, const int64_t *cs, const int64_t *os, const int64_t *cr,
                    const int64_t *orr) {
  Ctx &c = ctx();
  CB_NCCL(nccl.GroupStart());
  for (int p = 0; p < c.nranks; p++) {
    if (cs[p] > 0) CB_NCCL(nccl.Send((void *)(send + os[p]), (size_t)cs[p] * 2, kNcclDouble, p, c.nccl_comm, c.stream));
    if (cr[p] > 0) CB_NCCL(nccl.Recv((void *)(recv + orr[p]), (size_t)cr[p] * 2, kNcclDouble, p, c.nccl_comm, c.stream));
  }
  CB_NCCL(nccl.GroupEnd());
  return 0;
}

static int init_common(int device) {
  Ctx &c = ctx();
  if (c.inited) return fail("cdmft_b200: already initialised");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail("cdmft_b200: no CUDA device available (%s) -- this library has no CPU fallback",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return fail("cdmft_b200: device %d out of range (have %d)", device, ndev);
  CB_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  CB_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) return fail("cdmft_b200: device %s is sm_%d%d; this build is sm_100a only", prop.name, prop.major, prop.minor);
  c.device = device;
  c.sm_count = prop.multiProcessorCount;
  CB_CUDA(cudaStreamCreateWithFlags(&c.own_stream, cudaStreamNonBlocking));
  c.stream = c.own_stream;
  CB_CHECK(dev_alloc(&c.red, 4096));
  CB_CUDA(cudaMallocHost((void **)&c.red_host, 64 * sizeof(double)));
  c.launches = 0;
  c.inited = true;
  return 0;
}

}  // namespace cb

using namespace cb;

extern "C" {

const char *cdmft_b200_last_error(void) { return g_err.c_str(); }

int cdmft_b200_init(int32_t device) {
  CB_CHECK(init_common(device));
  Ctx &c = ctx();
  c.nranks = 1; c.rank = 0; c.spmd = false; c.sim = false;
  return 0;
}

int cdmft_b200_init_sim(int32_t device, int32_t nranks) {
  if (nranks < 1) return fail("init_sim: nranks < 1");
  CB_CHECK(init_common(device));
  Ctx &c = ctx();
  c.nranks = nranks; c.rank = 0; c.spmd = false; c.sim = true;
  return 0;
}

int cdmft_b200_nccl_unique_id(void *uid128) {
  CB_CHECK(nccl_load());
  ncclUniqueId_t id;
  CB_NCCL(nccl.GetUniqueId(&id));
  memcpy(uid128, &id, sizeof id);
  return 0;
}

int cdmft_b200_init_rank(int32_t device, int32_t nranks, int32_t rank, const void *uid128) {
  if (nranks < 1 || rank < 0 || rank >= nranks) return fail("init_rank: bad rank %d of %d", rank, nranks);
  CB_CHECK(nccl_load());
  CB_CHECK(init_common(device));
  Ctx &c = ctx();
  c.nranks = nranks; c.rank = rank; c.spmd = true; c.sim = false;
  ncclUniqueId_t id;
  memcpy(&id, uid128, sizeof id);
  CB_NCCL(nccl.CommInitRank(&c.nccl_comm, nranks, id, rank));
  {
    int lo = 0, hi = 0;
    CB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CB_CUDA(cudaStreamCreateWithPriority(&c.comm_stream, cudaStreamNonBlocking, hi));
  }
  CB_CUDA(cudaEventCreateWithFlags(&c.ev_in, cudaEventDisableTiming));
  CB_CUDA(cudaEventCreateWithFlags(&c.ev_comm, cudaEventDisableTiming));
  // First use of a collective sets up its channels / peer connections (hundreds of ms): do it here, once, for the
  // two kinds this library issues -- the 2-double all-reduce of the Krylov scalars and the send/recv all-to-all of
  // the transposes (both streams) -- so that no product or Krylov step pays for it.
  if (nranks > 1) {
    double *w = c.red + 3008;  // 2*nranks doubles of scratch
    if (2 * nranks + 3008 > 4096) return fail("init_rank: too many ranks for the warm-up scratch");
    CB_CUDA(cudaMemsetAsync(w, 0, (size_t)2 * nranks * sizeof(double), c.stream));
    CB_NCCL(nccl.AllReduce(w, w, 2, kNcclDouble, kNcclSum, c.nccl_comm, c.stream));
    for (cudaStream_t st : {c.stream, c.comm_stream}) {
      CB_NCCL(nccl.GroupStart());
      for (int p = 0; p < nranks; p++) {
        CB_NCCL(nccl.Send((void *)(w + p), 1, kNcclDouble, p, c.nccl_comm, st));
        CB_NCCL(nccl.Recv((void *)(w + nranks + p), 1, kNcclDouble, p, c.nccl_comm, st));
      }
      CB_NCCL(nccl.GroupEnd());
      CB_CUDA(cudaStreamSynchronize(st));
    }
  }
  return 0;
}

int cdmft_b200_finalize(void) {
  Ctx &c = ctx();
  if (!c.inited) return 0;
  if (c.hstatus) cdmft_b200_delete_hv_sector();
  cudaStreamSynchronize(c.stream);
  if (c.nccl_comm) { nccl.CommDestroy(c.nccl_comm); c.nccl_comm = nullptr; }
  dev_free(c.red); dev_free(c.lz_hist); c.lz_hist_cap = 0; dev_free(c.rt_queue); dev_free(c.cross_tab); dev_free(c.dot_partial); c.dot_cap = 0;
  dev_free(c.stage_v); dev_free(c.stage_hv); c.stage_n = 0;
  for (auto &k : c.kv) dev_free(k);
  c.kv_n = 0;
  if (c.red_host) { cudaFreeHost(c.red_host); c.red_host = nullptr; }
  if (c.comm_stream) { cudaStreamSynchronize(c.comm_stream); cudaStreamDestroy(c.comm_stream); c.comm_stream = nullptr; }
  for (auto st : c.peer_stream) if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
  c.peer_stream.clear();
  for (auto ev : c.peer_done) if (ev) cudaEventDestroy(ev);
  c.peer_done.clear();
  if (c.ev_pack) { cudaEventDestroy(c.ev_pack); c.ev_pack = nullptr; }
  if (c.ev_in) { cudaEventDestroy(c.ev_in); c.ev_in = nullptr; }
  if (c.ev_comm) { cudaEventDestroy(c.ev_comm); c.ev_comm = nullptr; }
  if (c.own_stream) { cudaStreamDestroy(c.own_stream); c.own_stream = nullptr; }
  c.stream = nullptr;
  free_map_ops();
  lz_free_slots();
  c.inited = false; c.have_model = false;
  return 0;
}

// ---- CUDA IPC peer windows for the active sector (SPMD) --------------------------------------
// 1. every rank: cdmft_b200_ipc_export(buf)   -> 2 x 64-byte handles of its vt and recvbuf
// 2. host program all-gathers the 128-byte records (MPI_Allgather / torch.distributed.all_gather)
// 3. every rank: cdmft_b200_ipc_import(all, nranks) -> peers' buffers are mapped; the distributed
//    transpose then runs as one kernel per direction that stores straight into peer memory.
int cdmft_b200_ipc_export(void *handles128) {
  CB_REQUIRE_INIT();
  Ctx &c = ctx();
  memset(handles128, 0, 128);
  if (!c.spmd || !c.hstatus) return fail("ipc_export: needs SPMD mode and an active sector");
  if (c.rk.empty() || !c.rk[0].vt || !c.rk[0].recvbuf) return 0;  // rank outside the shrunk communicator
  cudaIpcMemHandle_t h[2];
  CB_CUDA(cudaIpcGetMemHandle(&h[0], c.rk[0].vt));
  CB_CUDA(cudaIpcGetMemHandle(&h[1], c.rk[0].recvbuf));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  memcpy(handles128, h, 128);
  return 0;
}

int cdmft_b200_ipc_import(const void *all_handles, int32_t nranks) {
  CB_REQUIRE_INIT();
  Ctx &c = ctx();
  if (!c.spmd || !c.hstatus) return fail("ipc_import: needs SPMD mode and an active sector");
  if (nranks != c.nranks) return fail("ipc_import: nranks mismatch");
  ipc_close_peers();
  if (c.rk.empty() || c.p_eff < 2) return 0;
  c.peer_vt.assign(c.nranks, nullptr);
  c.peer_recv.assign(c.nranks, nullptr);
  const char *hp = (const char *)all_handles;
  for (int p = 0; p < c.p_eff; p++) {
    if (p == c.rank) { c.peer_vt[p] = c.rk[0].vt; c.peer_recv[p] = c.rk[0].recvbuf; continue; }
    cudaIpcMemHandle_t h[2];
    memcpy(h, hp + (size_t)p * 128, 128);
    void *a = nullptr, *b = nullptr;
    CB_CUDA(cudaIpcOpenMemHandle(&a, h[0], cudaIpcMemLazyEnablePeerAccess));
    CB_CUDA(cudaIpcOpenMemHandle(&b, h[1], cudaIpcMemLazyEnablePeerAccess));
    c.peer_vt[p] = (double2 *)a;
    c.peer_recv[p] = (double2 *)b;
  }
  c.ipc_ready = true;
  return 0;
}

int cdmft_b200_set_stream(void *s) {
  CB_REQUIRE_INIT();
  ctx().stream = (cudaStream_t)s;  // NULL = the legacy default stream, as in CUDA
  return 0;
}

int cdmft_b200_reset_stream(void) {
  CB_REQUIRE_INIT();
  ctx().stream = ctx().own_stream;
  return 0;
}

// sum of the CUDA-event durations recorded for one kernel kind since the last query (clears them)
int cdmft_b200_profile_query(int32_t kind, double *ms_total, int64_t *count) {
  CB_REQUIRE_INIT();
  Ctx &c = ctx();
  CB_CUDA(cudaStreamSynchronize(c.stream));
  double tot = 0;
  int64_t n = 0;
  std::vector<Ctx::ProfRec> keep;
  for (auto &r : c.prof) {
    if (r.kind != kind) { keep.push_back(r); continue; }
    float ms = 0;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) { tot += ms; n++; }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  c.prof.swap(keep);
  if (ms_total) *ms_total = tot;
  if (count) *count = n;
  return 0;
}

int cdmft_b200_launch_count(int64_t *n) {
  *n = ctx().launches;
  return 0;
}

int cdmft_b200_set_option(const char *key, int64_t value) {
  Ctx &c = ctx();
  std::string k(key);
  if (k == "colpass_variant") c.opt.colpass_variant = value;
  else if (k == "rowpass_variant") c.opt.rowpass_variant = value;
  else if (k == "lanczos_batch") c.opt.lanczos_batch = value;
  else if (k == "lanczos_store") c.opt.lanczos_store = value;
  else if (k == "xchg_chunks") c.opt.xchg_chunks = value;
  else if (k == "xchg_split") c.opt.xchg_split = value;
  else if (k == "colres_pair") c.opt.colres_pair = value;
  else if (k == "direct_tables") c.opt.direct_tables = value;
  else if (k == "force_sharded") c.opt.force_sharded = value;
  else if (k == "col_batch") c.opt.col_batch = value;
  else if (k == "row_slab") c.opt.row_slab = value;
  else if (k == "tma2d") c.opt.tma2d = value;
  else if (k == "overlap") c.opt.overlap = value;
  else if (k == "real_lanczos") c.opt.real_lanczos = value;
  else if (k == "row_rb") c.opt.row_rb = value;
  else if (k == "use_ipc") c.opt.use_ipc = value;
  else if (k == "sched") c.opt.sched = value;
  else if (k == "fuse_dot") c.opt.fuse_dot = value;
  else if (k == "fast4") c.opt.fast4 = value;
  else if (k == "colres_rows") c.opt.colres_rows = value;
  else if (k == "rowres_cols") c.opt.rowres_cols = value;
  else if (k == "profile") c.profile = value != 0;
  else return fail("set_option: unknown key %s", key);
  return 0;
}

// ------------------------------------------------------------------------------------
// Model ingestion.  One-body matrix per spin h_s(a,b), a,b over the Ns orbitals of one spin:
//   cluster   impHloc(ilat,jlat,s,s,iorb,jorb)              sparse/H_up.f90:8-30
//   replica   Hbath(ilat,jlat,s,s,iorb,jorb,ib)             sparse/H_up.f90:32-58
//   hybrid    V(ilat,s,iorb,ib) both directions             sparse/H_up.f90:61-87
// Off-diagonal non-zeros become the hop-term list; diagonals go to the diagonal term.
// ------------------------------------------------------------------------------------
int cdmft_b200_set_model(const cdmft_b200_model *m) {
  CB_REQUIRE_INIT();
  Ctx &c = ctx();
  if (c.hstatus) return fail("set_model: a sector is active; call delete_hv_sector first");
  if (m->nlat < 1 || m->norb < 1 || m->norb > 5 || m->nspin < 1 || m->nspin > 2 || m->nbath < 0)
    return fail("set_model: bad dimensions Nlat=%d Norb=%d Nspin=%d Nbath=%d", m->nlat, m->norb, m->nspin, m->nbath);
  const int L = m->nlat, O = m->norb, S = m->nspin, B = m->nbath;
  const int nimp = L * O, ns = nimp * (B + 1);
  if (ns > 30) return fail("set_model: Ns=%d > 30 not supported (int32 Fock states per spin)", ns);
  if (nimp > 12) return fail("set_model: Nimp=%d > 12 not supported (cross-term table)", nimp);
  c.m = *m;
  c.ns = ns; c.nimp = nimp; c.nlso = L * S * O;
  const int64_t nh = (int64_t)L * L * S * S * O * O;
  c.imphloc.assign(m->imphloc, m->imphloc + 2 * nh);
  c.hbath.assign(m->hbath, m->hbath + 2 * nh * B);
  c.vbath.assign(m->vbath, m->vbath + (int64_t)c.nlso * B);
  c.m.imphloc = c.imphloc.data(); c.m.hbath = c.hbath.data(); c.m.vbath = c.vbath.data();
  c.jhflag = (O > 1 && (m->jx != 0.0 || m->jp != 0.0));  // ED_SETUP.f90:200-201

  auto hidx = [&](int il, int jl, int s, int iorb, int jorb) {
    return (int64_t)il + (int64_t)L * (jl + (int64_t)L * (s + (int64_t)S * (s + (int64_t)S * (iorb + (int64_t)O * jorb))));
  };
  auto build = [&](int s, std::vector<Term> &terms, std::vector<double> &e) -> int {
    std::vector<double> hre((size_t)ns * ns, 0.0), him((size_t)ns * ns, 0.0);
    for (int il = 0; il < L; il++)
      for (int jl = 0; jl < L; jl++)
        for (int io = 0; io < O; io++)
          for (int jo = 0; jo < O; jo++) {
            int a = io + il * O, b = jo + jl * O;  // imp_state_index - 1
            int64_t k = hidx(il, jl, s, io, jo);
            hre[(size_t)a * ns + b] += c.imphloc[2 * k];
            him[(size_t)a * ns + b] += c.imphloc[2 * k + 1];
            for (int ib = 0; ib < B; ib++) {
              int aa = nimp + a + ib * nimp, bb = nimp + b + ib * nimp;  // getBathStride - 1
              hre[(size_t)aa * ns + bb] += c.hbath[2 * (k + nh * ib)];
              him[(size_t)aa * ns + bb] += c.hbath[2 * (k + nh * ib) + 1];
            }
          }
    for (int il = 0; il < L; il++)
      for (int io = 0; io < O; io++)
        for (int ib = 0; ib < B; ib++) {
          int a = io + il * O, aa = nimp + a + ib * nimp;
          double v = c.vbath[(io + il * O + s * O * L) + (int64_t)c.nlso * ib];  // index_stride_lso
          hre[(size_t)a * ns + aa] += v;
          hre[(size_t)aa * ns + a] += v;
        }
    terms.clear();
    for (int a = 0; a < ns; a++)
      for (int b = 0; b < ns; b++) {
        if (a == b) continue;
        double re = hre[(size_t)a * ns + b], im = him[(size_t)a * ns + b];
        if (re != 0.0 || im != 0.0) {
          terms.push_back({a, b, re, im});
          if (im != 0.0) c.real_h = false;
        }
      }
    // diagonal one-body energies
    e.assign(ns, 0.0);
    const double hf = m->hfmode ? 1.0 : 0.0;
    for (int il = 0; il < L; il++)
      for (int io = 0; io < O; io++) {
        int a = io + il * O;
        if (him[(size_t)a * ns + a] != 0.0) return fail("set_model: impHloc diagonal has an imaginary part (non-Hermitian H unsupported)");
        e[a] = hre[(size_t)a * ns + a] - m->xmu - hf * 0.5 * m->uloc[io];
        if (O > 1) e[a] -= hf * 0.5 * (O - 1) * (m->ust + (m->ust - m->jh));
        for (int ib = 0; ib < B; ib++) {
          // bath_diag = Re Hbath(ilat,ilat,s,s,iorb,iorb,ib); quirk: direct path loops ilat=1..Norb only
          bool keep = m->quirk_direct_bathdiag ? (il < O) : true;
          int aa = nimp + a + ib * nimp;
          e[aa] = keep ? hre[(size_t)aa * ns + aa] : 0.0;
        }
      }
    return 0;
  };
  c.real_h = true;
  CB_CHECK(build(0, c.terms_up, c.e_up));
  CB_CHECK(build(S - 1, c.terms_dw, c.e_dw));

  // interaction coefficients (sparse/H_local.f90:33-77)
  c.spair.assign((size_t)nimp * nimp, 0.0);
  c.wcross.assign((size_t)nimp * nimp, 0.0);
  c.const0 = 0.0;
  const double hf = m->hfmode ? 1.0 : 0.0;
  for (int il = 0; il < L; il++)
    for (int io = 0; io < O; io++) {
      int a = io + il * O;
      c.wcross[(size_t)a * nimp + a] = m->uloc[io];
      c.const0 += hf * 0.25 * m->uloc[io];
      if (O > 1)
        for (int jo = io + 1; jo < O; jo++) {
          int b = jo + il * O;
          c.wcross[(size_t)a * nimp + b] = m->ust;
          c.wcross[(size_t)b * nimp + a] = m->ust;
          c.spair[(size_t)a * nimp + b] = m->ust - m->jh;
          c.const0 += hf * 0.25 * (m->ust + (m->ust - m->jh));
        }
    }
  // cross table T[b][mu] = sum_a W[a][b] n_a(mu)
  const int64_t nst = (int64_t)1 << nimp;
  std::vector<double> tab((size_t)nimp * nst, 0.0);
  for (int b = 0; b < nimp; b++)
    for (int64_t mu = 0; mu < nst; mu++) {
      double s = 0;
      for (int a = 0; a < nimp; a++)
        if ((mu >> a) & 1) s += c.wcross[(size_t)a * nimp + b];
      tab[(size_t)b * nst + mu] = s;
    }
  dev_free(c.cross_tab);
  CB_CHECK(dev_alloc(&c.cross_tab, (int64_t)tab.size()));
  CB_CUDA(cudaMemcpy(c.cross_tab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
  free_map_ops();  // Fock maps cached for the previous model (Ns may have changed)
  c.have_model = true;
  return 0;
}

int cdmft_b200_get_ns(int32_t *ns) {
  if (!ctx().have_model) return fail("get_ns: no model set");
  *ns = ctx().ns;
  return 0;
}

}  // extern "C"
