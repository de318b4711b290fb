/*
 * ed_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the Hamiltonian-times-vector hot path of
 * QcmPlab/CDMFT-LANC-ED, function by function, used ONLY by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * as the checker and the CPU baseline.  The product (cdmft-lanc-ed_b200) never
 * links, imports or executes anything from this directory.
 *
 * PARITY STATUS: "parity unpinned" by reference fixtures -- the reference ships
 * no tests, golden vectors or known-answer files for this path (SURVEY.md §4,
 * §8c) and cannot be compiled here (no Fortran compiler, no MPI, no SciFortran).
 * Substitute anchors (tests/test_oracle_pin.py): an independent full-Fock
 * Jordan-Wigner ED (oracle/jw_ed.py), dense == sparse == direct == MPI(P),
 * scipy eigsh, and -- the one anchor taken from the reference itself -- the
 * U=0 impurity Green's function of the whole pipeline against the reference's
 * analytic g0and_bath (ED_BATH_FUNCTIONS.f90:102-155).
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the reference root).  The Krylov routines restate SciFortran's SF_SP_LINALG
 * (un-vendored, no version pin; driver comment says 4.10.8) from its published
 * algorithm, see SURVEY.md App. B.
 */
#ifndef ED_ORACLE_H
#define ED_ORACLE_H
#include <stdint.h>
#include <complex.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef double complex edo_c64;

/* Model parameters = what ED_HAMILTONIAN reads from module globals:
 * impHloc (ED_VARS_GLOBAL.f90:119), Hbath_build(lambda) per replica
 * (ED_BATH/hbath_setup.f90:240-250), dmft_bath%item(ib)%v, and the interaction
 * scalars (ED_INPUT_VARS.f90:19-32).  Complex arrays are interleaved (re,im)
 * in Fortran (column-major) order. */
typedef struct edo_model {
  int32_t nlat, norb, nspin, nbath;
  double uloc[5];
  double ust, jh, jx, jp, xmu;
  int32_t hfmode;
  const double *imphloc; /* [Nlat,Nlat,Nspin,Nspin,Norb,Norb] complex */
  const double *hbath;   /* [Nlat,Nlat,Nspin,Nspin,Norb,Norb,Nbath] complex */
  const double *vbath;   /* [Nlso,Nbath] real; Nlso index = index_stride_lso (ED_AUX_FUNX.f90:81-87) */
} edo_model;

typedef struct edo_ctx edo_ctx;

/* which mat-vec restatement */
enum {
  EDO_SPARSE_SERIAL = 0, /* spMatVec_main           ED_HAMILTONIAN_SPARSE_HxV.f90:167-227 */
  EDO_SPARSE_MPI = 1,    /* spMatVec_mpi_main       ED_HAMILTONIAN_SPARSE_HxV.f90:230-315 */
  EDO_DIRECT_SERIAL = 2, /* directMatVec_main       ED_HAMILTONIAN_DIRECT_HxV.f90:37-90   */
  EDO_DIRECT_MPI = 3     /* directMatVec_MPI_main   ED_HAMILTONIAN_DIRECT_HxV.f90:94-171  */
};

edo_ctx *edo_create(const edo_model *m);
void edo_destroy(edo_ctx *c);
const char *edo_last_error(void);

/* ---- dimensions / sectors (ED_SETUP.f90:111-120, 446-520, 1019-1037) ---- */
int32_t edo_ns(const edo_ctx *c);
int32_t edo_nsectors(const edo_ctx *c);
int64_t edo_binomial(int32_t n1, int32_t n2);
int32_t edo_get_sector(int32_t ns, int32_t nup, int32_t ndw);
void edo_get_nup_ndw(int32_t ns, int32_t isector, int32_t *nup, int32_t *ndw);
int64_t edo_get_dim(int32_t ns, int32_t isector, int64_t *dimup, int64_t *dimdw);
/* getCsector / getCDGsector (ED_SETUP.f90:377-418): spin 1=up 2=dw; 0 if none */
int32_t edo_get_c_sector(int32_t ns, int32_t ispin, int32_t isector);
int32_t edo_get_cdg_sector(int32_t ns, int32_t ispin, int32_t isector);

/* ---- Fock maps and operators (ED_SETUP.f90:720-775, 807-833, 935-945, 1044-1061) ---- */
int64_t edo_build_sector_map(int32_t ns, int32_t n, int32_t *map /* may be NULL to count */);
int32_t edo_c(int32_t pos, int32_t in, int32_t *out, double *fsgn);   /* rc!=0: "C error" stop */
int32_t edo_cdg(int32_t pos, int32_t in, int32_t *out, double *fsgn);
int32_t edo_binary_search(const int32_t *a, int32_t n, int32_t value); /* 1-based, 0 = not found */
int32_t edo_imp_state_index(const edo_ctx *c, int32_t ilat, int32_t iorb);               /* 1-based */
int32_t edo_get_bath_stride(const edo_ctx *c, int32_t ilat, int32_t iorb, int32_t ibath); /* 1-based */

/* ---- ED_SPARSE_MAP (ED_SPARSE_MAP.f90:54-157; filled by build_sector(itrace)) ----
 * Flattened: for imp state k in [0,2^Nimp): entries rowptr[k]..rowptr[k+1]-1 hold
 * bath_state[] and sector_indx[] in insertion order. Returns total entries. */
int64_t edo_build_sparse_map(const edo_ctx *c, int32_t n, int64_t *rowptr, int32_t *bath_state,
                             int32_t *sector_indx);
int32_t edo_sparse_map_intersection(const int64_t *rowptr, const int32_t *bath_state, int32_t iimp,
                                    int32_t jimp, int32_t *out);

/* ---- sharding (ED_HAMILTONIAN.f90:92-105, 197-221) ---- */
typedef struct edo_shard {
  int64_t qdw, rdw, q, r, istart, iend, ishift; /* mpiQdw,mpiRdw,mpiQ,mpiR,mpiIstart,mpiIend,mpiIshift */
  int64_t qup, up_off;                          /* mpiQup and first owned up row (0-based) */
  int64_t dw_off;                               /* first owned dw column (0-based) */
} edo_shard;
void edo_shard_of(int64_t dimup, int64_t dimdw, int32_t P, int32_t rank, edo_shard *s);
int64_t edo_vecdim(int64_t dimup, int64_t dimdw, int32_t P, int32_t rank);

/* ---- build_Hv_sector / delete_Hv_sector (ED_HAMILTONIAN.f90:39-190) ----
 * P>1 simulates the MPI run in-process (ranks = loop iterations / OpenMP threads).
 * P is clamped to DimDw as the reference shrinks the communicator (:62-90);
 * the effective P is returned by edo_active_ranks(). quirk_direct_bathdiag
 * reproduces direct/HxV_local.f90:83-84 (ilat loop bound = Norb). */
int32_t edo_build_hv_sector(edo_ctx *c, int32_t isector, int32_t kind, int32_t P,
                            int32_t quirk_direct_bathdiag);
int32_t edo_delete_hv_sector(edo_ctx *c);
int32_t edo_active_ranks(const edo_ctx *c);
int64_t edo_sector_dims(const edo_ctx *c, int64_t *dimup, int64_t *dimdw);

/* spHtimesV_p on the GLOBAL vector (shards of the P simulated ranks are the
 * contiguous blocks of v in rank order, which is how gather_vector_MPI lays
 * them out, ED_SETUP.f90:633-668). */
int32_t edo_hxv(edo_ctx *c, int64_t n, const edo_c64 *v, edo_c64 *hv);

/* vector_transpose_MPI (ED_HAMILTONIAN_COMMON.f90:30-101) on P simulated ranks:
 * a = concatenated column blocks a_r(nrow,qcol_r); b = concatenated b_r(ncol,qrow_r). */
int32_t edo_vector_transpose_sim(int32_t P, int64_t nrow, int64_t ncol, const edo_c64 *a, edo_c64 *b);

/* inspection: per-spin CSR (which: 1=up, 2=dw), 1-based cols in insertion order;
 * diagonal spH0d values for global rows; non-local spH0nd */
int64_t edo_get_csr(const edo_ctx *c, int32_t which, int64_t *rowptr, int32_t *col, edo_c64 *val);
int32_t edo_get_diag(const edo_ctx *c, double *d /* [Dim] */);

/* ---- sampled rows of H x v against a counter-based synthetic vector -------------------------------
 * For sectors whose vector does not fit in host memory (Ns = 18 half filling: 38 GB; the reference itself cannot
 * address it, ED_SETUP.f90:321 int32 getDim).  v(i) = scale * (re, im)(hash(i, seed)) is a pure function of the
 * 0-based GLOBAL index, so any shard of it can be generated anywhere (edo_counter_vec here, synth.py for numpy /
 * torch) and single rows of H x v evaluated without materialising v.  Row i is evaluated with the pull form of the
 * direct products (direct/HxV_local.f90, HxV_up.f90, HxV_dw.f90, HxV_non_local.f90): the reference pushes
 * H(k,i) v(i) from every column i; a single ROW collects H(i,k) v(k) = conjg(H(k,i)) v(k) over the hops of its
 * own state (H is Hermitian by construction, checked against the Jordan-Wigner ED in tests/test_oracle_pin.py).
 * Needs an active sector of any kind (only the Fock maps are used). */
void edo_counter_vec(int64_t i0, int64_t n, uint64_t seed, double scale, edo_c64 *out);
int32_t edo_hxv_rows_counter(edo_ctx *c, int64_t nrows, const int64_t *rows /* 0-based global */, uint64_t seed, double scale,
                             edo_c64 *out /* [nrows] */);
int64_t edo_get_nonlocal(const edo_ctx *c, int64_t *rowptr, int64_t *col, edo_c64 *val);
/* dense Hmat = diag + kron(Hdw,1) + kron(1,Hup) (+nd)  (ED_HAMILTONIAN_SPARSE_HxV.f90:112-148) */
int32_t edo_dense_hmat(edo_ctx *c, int32_t isector, edo_c64 *hmat /* [Dim,Dim] column-major */);

/* ---- Krylov (SciFortran SF_SP_LINALG restated) ---- */
/* sp_lanc_tridiag: alanc[0..n-1], blanc[0..n-1] (blanc[0] unused =0). returns #iterations done */
int32_t edo_lanc_tridiag(edo_ctx *c, int64_t n, edo_c64 *vin, int32_t nitermax, double *alanc,
                         double *blanc, double threshold);
/* sp_lanc_eigh: vect in = start vector (all zero -> constant 1/sqrt(N)), out = eigenvector */
int32_t edo_lanc_eigh(edo_ctx *c, int64_t n, double *egs, edo_c64 *vect, int32_t nitermax,
                      double threshold, int32_t ncheck, int32_t *niter_out, double *alanc,
                      double *blanc);
/* eigh(diag,subdiag,Ev): symmetric tridiagonal, ascending eigenvalues, Z column-major [n,n] */
int32_t edo_tridiag_eigh(int32_t n, double *d, double *e /* e[1..n-1] used, e[0] ignored */, double *z);

/* ---- Green's function pieces (ED_GF_NORMAL.f90:123-306, 531-903, 915-975) ---- */
/* vvinit = sum_k coef[k] * op_k |state>, op = c (iop=-1) or cdg (iop=+1) on orbital pos (1-based),
 * spin 1=up 2=dw. state lives in isector, result in jsector (caller sizes it, zero-filled here). */
/* lanc_observables master loop (ED_OBSERVABLES.f90:120-192); Fortran-ordered outputs, accumulated (+=) */
int32_t edo_lanc_observables(int32_t ns, int32_t nlat, int32_t norb, int32_t isector, const edo_c64 *vec, double peso,
                             double *dens_up, double *dens_dw, double *docc, double *magz, double *s2tot, double *sz2,
                             double *n2);
/* lanc_local_energy (ED_OBSERVABLES.f90:246-460): out[5] += {Eknot, Epot (before + Ehartree), Ehartree, Dust, Dund} of one
 * eigenstate `vec` (full sector vector) with weight peso; see ed_oracle.c for the term-by-term restatement. */
int32_t edo_lanc_local_energy(const edo_ctx *c, int32_t isector, const edo_c64 *vec, double peso, double *out);
/* density_matrix_impurity (ED_OBSERVABLES.f90:465-686): cdm [4^Nimp,4^Nimp] += Tr_BATH |vec><vec| peso (io = IimpUp + 2^Nimp IimpDw),
 * spdm [Nlat,Nlat,Nspin,Nspin,Norb,Norb] += <C^+_a C_b> peso; either may be NULL. */
int32_t edo_density_matrix_impurity(const edo_ctx *c, int32_t isector, const edo_c64 *vec, double peso, edo_c64 *cdm, edo_c64 *spdm);
int32_t edo_apply_op(int32_t ns, int32_t isector, int32_t iop, int32_t ispin, int32_t nops,
                     const int32_t *pos, const edo_c64 *coef, const edo_c64 *state, edo_c64 *out,
                     int32_t *jsector_out);
/* add_to_lanczos_gf_normal: accumulates peso/(i wm - isign*de) into g[0..lmats-1];
 * also returns poles/weights (size nlanc) if non-NULL */
int32_t edo_add_to_lanczos_gf(edo_c64 vnorm2, double ei, int32_t nlanc, const double *alanc,
                              const double *blanc, int32_t isign, double zeta, int32_t lmats,
                              const double *wm, edo_c64 *g, double *poles, edo_c64 *weights);
int32_t edo_add_to_lanczos_gf_full(edo_c64 vnorm2, double ei, double egs, int32_t finite_t, double beta, int32_t nlanc,
                                   const double *alanc, const double *blanc, int32_t isign, double zeta, int32_t lmats,
                                   const double *wm, edo_c64 *gm, int32_t lreal, const double *wr, double eps, edo_c64 *gr,
                                   double *poles, edo_c64 *weights);

/* number of OpenMP threads the MPI-simulating paths will use */
int32_t edo_num_threads(void);
void edo_set_num_threads(int32_t n);

#ifdef __cplusplus
}
#endif
#endif
