/*
 * cdmft_b200.h -- C ABI of libcdmft_b200.so (hand-written sm_100a CUDA behind it).
 *
 * Drop-in boundary for the Hamiltonian-times-vector hot path of QcmPlab/CDMFT-LANC-ED:
 * the three module procedures + one procedure pointer of the reference
 *
 *   build_Hv_sector(isector[,Hmat])   ED_HAMILTONIAN.f90:39-143
 *   delete_Hv_sector()                ED_HAMILTONIAN.f90:149-190
 *   vecDim_Hv_sector(isector)         ED_HAMILTONIAN.f90:197-221
 *   spHtimesV_p(Nloc,v,Hv)            ED_VARS_GLOBAL.f90:72-78,146  (abstract interface cc_sparse_HxV)
 *
 * plus the SciFortran Krylov drivers the callers wrap around the pointer
 * (sp_lanc_eigh ED_DIAG.f90:176-184, sp_lanc_tridiag ED_GF_NORMAL.f90:215) as fused
 * device-resident entry points, the start-vector construction of ED_GF_NORMAL.f90:180-199
 * and the pole/weight accumulation of ED_GF_NORMAL.f90:915-975.
 *
 * All functions return 0 on success, non-zero on error (message via
 * cdmft_b200_last_error(); the Fortran shim turns it into `stop`, the reference's only
 * error mechanism).  Plain pointers and sizes only; complex(8) = interleaved (re,im) doubles.
 * One global context, one active sector at a time (the reference keeps the sector in module
 * globals, ED_HAMILTONIAN_COMMON.f90:11-20); calls must be serialised by the caller.
 * There is NO CPU fallback: every entry point fails loudly without a CUDA device.
 */
#ifndef CDMFT_B200_H
#define CDMFT_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* What ED_HAMILTONIAN reads from module globals, packed by the caller:
 *   impHloc                       ED_VARS_GLOBAL.f90:119      complex [Nlat,Nlat,Nspin,Nspin,Norb,Norb]
 *   Hbath_build(lambda(:,ib))     ED_BATH/hbath_setup.f90:240-250   same shape x Nbath
 *   dmft_bath%item(ib)%v(:)       ED_HAMILTONIAN_SPARSE_HxV.f90:63-75    real [Nlat*Nspin*Norb, Nbath]
 *   Uloc,Ust,Jh,Jx,Jp,xmu,hfmode  ED_INPUT_VARS.f90:19-32,129-135,164
 * Arrays are Fortran (column-major) order. */
typedef struct cdmft_b200_model {
  int32_t nlat, norb, nspin, nbath;
  double uloc[5];
  double ust, jh, jx, jp, xmu;
  int32_t hfmode;
  /* reproduce the direct-path bath-diagonal loop bound (ilat=1..Norb) of
   * ED_HAMILTONIAN/direct/HxV_local.f90:83-84 instead of the sparse path's 1..Nlat
   * (sparse/H_local.f90:85); default 0 = sparse (correct) semantics */
  int32_t quirk_direct_bathdiag;
  const double *imphloc;
  const double *hbath;
  const double *vbath;
} cdmft_b200_model;

/* ed_sparse_H switch (ED_INPUT_VARS.f90:145, ED_HAMILTONIAN.f90:129-141) */
enum {
  CDMFT_B200_SPARSE = 1, /* stored per-spin CSR (spMatVec_main / spMatVec_MPI_main)            */
  CDMFT_B200_DIRECT = 0  /* matrix-free bit hopping (directMatVec_main / directMatVec_MPI_main) */
};

/* ---- runtime ------------------------------------------------------------------------ */
const char *cdmft_b200_last_error(void);
/* single rank on CUDA device `device` (MpiStatus=.false. in the reference) */
int cdmft_b200_init(int32_t device);
/* SPMD: one process per GPU (ed_set_MpiComm, ED_VARS_GLOBAL.f90:365-380).  `uid` is the
 * 128-byte NCCL unique id produced on rank 0 by cdmft_b200_nccl_unique_id and broadcast by
 * the host program (MPI_Bcast in Fortran, torch.distributed here). */
int cdmft_b200_nccl_unique_id(void *uid128);
int cdmft_b200_init_rank(int32_t device, int32_t nranks, int32_t rank, const void *uid128);
/* P simulated ranks on ONE device: the sharded code path (Ndw split + distributed transpose)
 * with device-local exchange; hxv then takes the gathered vector (shards = contiguous blocks
 * in rank order, the layout of gather_vector_MPI, ED_SETUP.f90:633-668). Test/debug aid. */
int cdmft_b200_init_sim(int32_t device, int32_t nranks);
int cdmft_b200_finalize(void);
/* SPMD, optional, after build_hv_sector: peer-memory windows for the distributed transpose.
 * Every rank exports 128 bytes (two CUDA IPC handles), the host program all-gathers them
 * (MPI_Allgather / torch.distributed) and every rank imports the nranks*128-byte array.  The
 * transposes then store straight into the destination GPU over NVLink (one kernel per direction,
 * bracketed by stream-ordered NCCL barriers) instead of pack -> all-to-all -> unpack.
 * delete_hv_sector unmaps the windows (collectively). */
int cdmft_b200_ipc_export(void *handles128);
int cdmft_b200_ipc_import(const void *all_handles, int32_t nranks);
/* launch on this cudaStream_t (NULL = CUDA's legacy default stream) instead of the library's own
 * non-blocking stream; reset_stream goes back to the library's stream */
int cdmft_b200_set_stream(void *cuda_stream);
int cdmft_b200_reset_stream(void);
/* with set_option("profile",1) the library brackets its kernels with CUDA events on the launch
 * stream; this returns the summed duration and count for one kind since the last query
 * (0 column pass, 1 row pass, 2 transpose/pack/unpack, 3 NCCL all-to-all) and clears them */
int cdmft_b200_profile_query(int32_t kind, double *ms_total, int64_t *count);
/* number of kernels this library has launched so far */
int cdmft_b200_launch_count(int64_t *n);
/* kernel variant selection for experiments/benchmarks: key/value, see DESIGN.md */
int cdmft_b200_set_option(const char *key, int64_t value);

/* ---- model -------------------------------------------------------------------------- */
int cdmft_b200_set_model(const cdmft_b200_model *m);
int cdmft_b200_get_ns(int32_t *ns); /* ed_setup_dimensions, ED_SETUP.f90:111-120 */

/* ---- sector lifecycle ----------------------------------------------------------------- */
/* isector = 1 + Nup*(Ns+1) + Ndw  (get_Sector, ED_SETUP.f90:446-457) */
int cdmft_b200_get_sector_dims(int32_t isector, int64_t *dimup, int64_t *dimdw, int64_t *dim);
/* vecDim_Hv_sector: DimUp*(DimDw/P + [rank < mod(DimDw,P)]); Dim for sim/single rank */
int cdmft_b200_vecdim_hv_sector(int32_t isector, int64_t *vecdim);
/* build_Hv_sector: Fock maps, Ndw shard (mpiQdw, mpiIstart, ...), H terms, per-spin CSR
 * (mode SPARSE) and the device work buffers.  *nloc receives the local vector length.
 * Building on an active sector is an error (reference: "alreay allocate can not init",
 * ED_SPARSE_MATRIX.f90:134). */
int cdmft_b200_build_hv_sector(int32_t isector, int32_t mode, int64_t *nloc);
int cdmft_b200_delete_hv_sector(void);
/* ranks that hold data for the active sector = min(P, DimDw) (ED_HAMILTONIAN.f90:62-90) */
int cdmft_b200_active_ranks(int32_t *p);

/* ---- the mat-vec: spHtimesV_p(Nloc,v,Hv) ---------------------------------------------- */
/* v, hv: complex(8)[nloc], host OR device pointers (detected); hv is fully overwritten and
 * must not alias v.  Synchronous on return for host pointers.  For device pointers the kernels are
 * only enqueued, on the library's own non-blocking stream unless cdmft_b200_set_stream was called:
 * the caller orders its producers/consumers of v and hv against that stream (or passes its stream).
 * nloc must equal the value build_hv_sector returned. */
int cdmft_b200_hxv(int32_t nloc, const void *v, void *hv);
int cdmft_b200_hxv64(int64_t nloc, const void *v, void *hv); /* Ns=18: Dim > 2^31 */

/* ---- inspection (parity tests: integer objects are bit-exact with the reference) ------ */
/* which: 1 = up (Hs(1)%map / spH0ups(1)), 2 = dw (Hs(2)%map / spH0dws(1)) */
int cdmft_b200_get_sector_map(int32_t which, int32_t *map);
int cdmft_b200_get_csr_nnz(int32_t which, int64_t *nnz);
/* canonical CSR: rowptr[n+1] 0-based offsets, col 1-based ascending (= the reference's
 * insertion order), val complex interleaved */
int cdmft_b200_get_csr(int32_t which, int64_t *rowptr, int32_t *col, double *val);
/* spH0d values of the local rows (sparse/H_local.f90) */
int cdmft_b200_get_diag(int64_t nloc, double *d);
/* ED_SPARSE_MAP of build_sector(...,itrace=.true.) (ED_SETUP.f90:757-759,
 * ED_SPARSE_MAP.f90:101-121): for imp state k in [0,2^Nimp) entries rowptr[k]..rowptr[k+1]-1 */
int cdmft_b200_get_sparse_map(int32_t which, int64_t *rowptr, int32_t *bath_state, int32_t *sector_indx);

/* Host-only (no CUDA call): the conflict-free gather schedule the column-resident kernel runs
 * (k_colres) for a CSR pattern with 0-based ascending columns.  g = 8 (16-byte vector
 * elements) or 16 (8-byte); nwarps = warps per CTA the tasks are dealt to.  code[nnz] = 7-bit coefficient id.
 * Tasks are numbered warp-major: warp w owns tasks tbase[w]..tbase[w+1]-1 and the quads (4 steps x 32 lanes)
 * from qbase[w] on, task after task.  meta[ntask*32*4]: per task and lane {f_row (2 words, 0 here), row index
 * (0xFFFFFFFF = none), mu | nquad << 16}; words[nquads*32*4]: per quad and lane 4 words (source row << 7) | code;
 * idle lanes read one of the g zero elements behind the column (rows >= ceil(n/g)*g, code 0).
 * Call with words == NULL to get *ntask / *nquads first.  Used by the CPU tests to check the schedule. */
int cdmft_b200_schedule_host(int64_t n, const int32_t *rowptr, const int32_t *col, const uint8_t *code,
                             int32_t g, int32_t natural, int32_t nwarps, int32_t *ntask, int64_t *nquads,
                             int32_t *tbase, int32_t *qbase, uint32_t *meta, uint32_t *words);

/* Host-only: the block-split schedules of k_colblk (columns larger than shared memory, Ns = 18)
 * for the CSR pattern of a one-spin operator on the sector of `npart` particles in `ns` orbitals (n = C(ns,npart) rows,
 * 0-based ascending columns).  Blocks = runs of states sharing their top bits with at most cap_rows rows.
 * sizes[8] = {nblk, len(tbase) = len(qbase) = nblk*(nwarps+1), len(meta), len(words), ntask, len(woff), nwarps, max rows};
 * blk[nblk*4] = {first row, rows, first task, first unit}; tbase / qbase relative to the block; meta (uint4 per task and
 * lane: row relative to the block in word 2, units << 16 in word 3); words: uint4 per unit and lane,
 * (source row relative to the block << 7) | code, idle lanes on the zero elements behind the block; toff[ntask*2] =
 * {first off-block step, steps}; woff[step*32 + lane] = (source row << 7) | code, 0 = none.  Arrays may be NULL to
 * query the sizes.  Used by the CPU tests. */
int cdmft_b200_colblk_host(int32_t ns, int32_t npart, const int32_t *rowptr, const int32_t *col, const uint8_t *code,
                           int32_t g, int32_t natural, int64_t cap_rows, int64_t *sizes, int32_t *blk, int32_t *tbase,
                           int32_t *qbase, uint32_t *meta, uint32_t *words, uint32_t *toff, uint32_t *woff);

/* ---- fused device-resident Krylov drivers (SciFortran SF_SP_LINALG semantics) --------- */
/* sp_lanc_tridiag(MatVec,vin,alanc,blanc): v0 = local shard of the start vector (host or
 * device; not modified), alanc/blanc host arrays of size nitermax (blanc[0] = 0);
 * *ndone = iterations performed (stops early when |beta| < threshold, default 1e-12). */
int cdmft_b200_lanczos_tridiag(int64_t nloc, const void *v0, int32_t nitermax, double threshold,
                               double *alanc, double *blanc, int32_t *ndone);
/* sp_lanc_eigh(MatVec,egs,vect,Nitermax,threshold,ncheck): vect in = start vector (all zero
 * -> constant 1/sqrt(Dim)), out = normalised ground-state shard; alanc/blanc may be NULL. */
int cdmft_b200_lanczos_gs(int64_t nloc, void *vect, int32_t nitermax, double threshold,
                          int32_t ncheck, double *egs, int32_t *niter, double *alanc, double *blanc);

/* sp_eigh(MatVec, eig_values, eig_basis, Nblock, Nitermax, tol) -- the reference's DEFAULT LANC_METHOD
 * (ED_DIAG.f90:94-97,150-170; SciFortran wraps (P)ARPACK in reverse communication, which = 'SA', nev = Neigen,
 * ncv = Nblock, mxiter = Nitermax): the neigen lowest eigenpairs of the ACTIVE sector by a device-resident
 * thick-restart Lanczos with full reorthogonalisation (the same Krylov subspaces as ARPACK's implicit restart for a
 * Hermitian operator; csrc/trlan.h).  Collective in SPMD mode (the P-ARPACK branch, ED_DIAG.f90:153-158): every
 * rank passes its shard sizes, dot products are all-reduced, the pseudo-random start vector is a function of the
 * global index, so the result does not depend on the number of ranks beyond rounding.
 *   eig_values[neigen] ascending (host); eig_basis complex(8)[nloc, neigen] column-major (host or device; may be
 *   NULL); nblock <= 0 -> max(2 neigen, 20); nblock is capped at 64, at Dim - 1 and at what HBM holds.
 *   tol: a pair is converged when its Ritz bound <= max(tol, eps) * max(eps^(2/3), |theta|) (ARPACK's test; the
 *   reference's default 1e-18 acts as machine precision).  *nconv = converged pairs (< neigen: Nitermax restarts were
 *   not enough -- ARPACK's info = 1 -- the current Ritz pairs are returned); *nmatvec = H x v products used. */
int cdmft_b200_eigh(int64_t nloc, int32_t neigen, int32_t nblock, int32_t nitermax, double tol, double *eig_values,
                    void *eig_basis, int32_t *nconv, int32_t *nmatvec);
/* Host-only TEST HOOK (no CUDA call, like cdmft_b200_schedule_host): the restart logic of cdmft_b200_eigh run on host
 * vectors around the CALLER's mat-vec (complex(8)[n] -> complex(8)[n]); the CPU tests drive it with the oracle's
 * H x v.  It is not a product path and computes no Hamiltonian itself.  Sharded runs (the gloo test): n = local length,
 * goff = global index of the first local element, ntot = global length, allreduce = in-place sum of `count` doubles over
 * the ranks (NULL with one rank: goff = 0, ntot = n) -- the contract the device backend has with NCCL. */
int cdmft_b200_eigh_logic_host(int64_t n, int64_t goff, int64_t ntot,
                               void (*matvec)(int64_t n, const double *v, double *hv, void *user),
                               void (*allreduce)(double *buf, int64_t count, void *user), void *user, int32_t neigen,
                               int32_t nblock, int32_t nitermax, double tol, double *eig_values, double *eig_basis,
                               int32_t *nconv, int32_t *nmatvec);

/* ---- Green's function helpers (ED_GF_NORMAL.f90) --------------------------------------- */
/* out = sum_k coef[k] * op(pos[k]) |state>  with op = c^+ (iop=+1) or c (iop=-1) acting on spin
 * ispin (1 = up, 2 = dw) of a vector living in sector `isector`; the result lives in
 * getCDGsector/getCsector(isector) (ED_SETUP.f90:377-418; loops ED_GF_NORMAL.f90:180-194,
 * 244-258, 590-620).  state[dim(isector)], out[dim(jsector)] host or device.  SPMD: state / out are the local
 * shards of the two sectors.  Spin-up operators (all the reference applies when Nspin=1) act shard by shard -- the
 * Ndw split is the same in both sectors; spin-down operators change the split: the state is gathered on rank 0,
 * transformed there and scattered with the target sector's split (collective; out needs vecDim(jsector) elements).
 * pos is the 1-based orbital position imp_state_index(ilat,iorb); coef complex interleaved. */
int cdmft_b200_apply_op(int32_t isector, int32_t iop, int32_t ispin, int32_t nops, const int32_t *pos,
                        const double *coef, const void *state, void *out, int32_t *jsector);
/* add_to_lanczos_gf_normal (ED_GF_NORMAL.f90:915-975), T=0: poles/weights from the Lanczos
 * tridiagonal matrix, g[i] += peso/(i*wm[i] - isign*(E_j - ei)), peso = vnorm2*Z(1,j)^2/zeta.
 * g complex interleaved [lmats] on the host; poles/weights (size nlanc) may be NULL. */
int cdmft_b200_add_to_lanczos_gf(const double vnorm2[2], double ei, int32_t nlanc, const double *alanc,
                                 const double *blanc, int32_t isign, double zeta, int32_t lmats,
                                 const double *wm, double *g, double *poles, double *weights);

/* the same routine with all its branches: finite_t != 0 weighs the channel with exp(-beta (ei - egs)) / zeta (dropped
 * when beta (ei - egs) >= 200, :930-936), and the real-axis function is accumulated next to the Matsubara one,
 * greal[i] += peso / (wr[i] + i eps - isign (E_j - ei)) (:968-971).  Host arrays; gmats / greal may be NULL with
 * lmats / lreal = 0.  No device is needed (the tridiagonal problem is nlanc x nlanc). */
int cdmft_b200_add_to_lanczos_gf_full(const double vnorm2[2], double ei, double egs, int32_t finite_t, double beta,
                                      int32_t nlanc, const double *alanc, const double *blanc, int32_t isign, double zeta,
                                      int32_t lmats, const double *wm, double *gmats, int32_t lreal, const double *wr,
                                      double eps, double *greal, double *poles, double *weights);

/* ---- dense sector matrix and shard <-> full vector moves ---------------------------------------- */
/* The optional `Hmat` of build_Hv_sector(isector, Hmat) (ED_HAMILTONIAN.f90:123-127 ->
 * ED_HAMILTONIAN_SPARSE_HxV.f90:112-148; caller ED_DIAG.f90:199, the LAPACK branch): the dense Hamiltonian of the
 * ACTIVE sector, complex(8) [Dim, Dim] column-major, host or device pointer.  Every rank receives the whole matrix,
 * as in the reference.  Dim <= 32768. */
int cdmft_b200_build_hmat(void *hmat);
/* scatter_vector_MPI / gather_vector_MPI (ED_SETUP.f90:575-668) for the active sector: vfull = complex(8)[Dim]
 * on `root` (ignored elsewhere), vloc = this rank's shard [nloc]; host or device pointers.  Collective in SPMD
 * mode (NCCL send/recv); a plain copy with one rank or simulated ranks. */
int cdmft_b200_scatter_vector(const void *vfull, void *vloc, int32_t root);
int cdmft_b200_gather_vector(const void *vloc, void *vfull, int32_t root);

/* ---- local observables (ED_OBSERVABLES.f90:94-236, lanc_observables) --------------------------- */
/* w[mu + md*2^Nimp] = sum over all basis states whose up / dw Fock states have the impurity bits mu / md of
 * |vec|^2 (4^Nimp doubles, host).  Every observable of the reference's master loop (:120-192: dens, dens_up,
 * dens_dw, docc, magz, s2tot, sz2, n2) is a weighted sum over this table -- the host layer evaluates the
 * reference's formulas on it.  vec = local shard (host or device) of a vector of the ACTIVE sector; with
 * sharded vectors the tables are all-reduced so every rank receives the full one.  Nimp <= 8. */
int cdmft_b200_imp_weights(int64_t nloc, const void *vec, double *w);

/* ---- local energy (ED_OBSERVABLES.f90:246-460, lanc_local_energy) --------------------------------------- */
/* Every piece of the reference's master loop (:290-424) is a weighted sum over the impurity-configuration table of
 * cdmft_b200_imp_weights (Epot, Ehartree, Dust, Dund and the diagonal part of Eknot) except the hopping part of Eknot,
 * sum_i impHloc(is,js) sg1 sg2 vec(i) conjg(vec(j)) over |j> = c^+_is c_js |i> (:305-345) = <vec| K |vec> with K the
 * impurity block of the off-diagonal one-body matrix of both spins.  out = (Re, Im) of that expectation value,
 * all-reduced over the ranks; vec = local shard (host or device) of a vector of the ACTIVE sector. */
int cdmft_b200_imp_kinetic(int64_t nloc, const void *vec, double out[2]);

/* ---- density matrices of the impurity (ED_OBSERVABLES.f90:465-686, density_matrix_impurity) ------------------ */
/* For one eigenvector `vec` (local shard, host or device) of the ACTIVE sector with weight peso, ACCUMULATED (+=) like the
 * reference's sum over state_list; either output may be NULL:
 *   cdm  complex [4^Nimp, 4^Nimp] column-major: rho_IMP = Tr_BATH |vec><vec|, row / column label IimpUp + 2^Nimp*IimpDw
 *        (0-based; the reference's io - 1).  Computed as Gram matrices of the amplitude blocks that share a bath
 *        configuration (k_cluster_gram); needs the whole vector on every process (all-gathered in SPMD mode).
 *   spdm complex [Nlat,Nlat,Nspin,Nspin,Norb,Norb] column-major: <C^+_a C_b>; diagonal from the weight table of
 *        cdmft_b200_imp_weights, off-diagonal one matrix-free product per pair and spin block through the regular H x v path.
 * Host arrays.  Nimp <= 6. */
int cdmft_b200_density_matrices(int64_t nloc, const void *vec, double peso, double *cdm, double *spdm);

#ifdef __cplusplus
}
#endif
#endif
