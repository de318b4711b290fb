/*
 * ed_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 * See ed_oracle.h for the status header ("parity unpinned" by reference
 * fixtures; substitute anchors in tests/).  Every function cites the
 * reference file:line it restates (paths relative to the reference root).
 *
 * Conventions: all "positions" (pos, is, js, ialfa) are 1-based like the
 * Fortran; sector-map *indices* returned to callers are 1-based where the
 * header says so; C arrays are 0-based internally.
 */
#include "ed_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define EDO_MAXNS 30

static char g_err[512] = "";
const char *edo_last_error(void) { return g_err; }
#define FAIL(...)                                                                                  \
  do {                                                                                             \
    snprintf(g_err, sizeof g_err, __VA_ARGS__);                                                    \
    return -1;                                                                                     \
  } while (0)

/* growable row: sparse_row_csr (ED_SPARSE_MATRIX.f90:13-17) */
typedef struct {
  int32_t size, cap;
  edo_c64 *vals;
  int64_t *cols; /* 1-based; int64 because spH0d/spH0nd hold global columns */
} sp_row;
typedef struct {
  int64_t nrow, ncol;
  sp_row *row;
  int status;
} sp_mat;

typedef struct {
  /* per simulated MPI rank */
  edo_shard sh;
  /* spH0d always holds exactly ONE entry per local row (value htmp at column i, even when 0:
   * sparse/H_local.f90:95-100), so it is stored flat: d0[i-ishift-1] = htmp; column = global row. */
  edo_c64 *d0;
  sp_mat h0nd; /* local rows (ED_HAMILTONIAN_SPARSE_HxV.f90:77-91) */
} rank_state;

struct edo_ctx {
  edo_model m;
  double *imphloc, *hbath, *vbath; /* owned copies */
  int32_t ns, nimp, nlso;
  int jhflag; /* ED_SETUP.f90:200-201 */
  /* sector state = ED_HAMILTONIAN_COMMON.f90:11-20 */
  int hstatus;
  int32_t hsector, kind, P, quirk;
  int64_t dim, dimup, dimdw;
  int32_t *map_up, *map_dw; /* Hs(1)%map, Hs(2)%map */
  sp_mat h0up, h0dw;        /* spH0ups(1), spH0dws(1) */
  rank_state *rk;
};

/* ------------------------------------------------------------------ */
/* complex array accessors (Fortran column-major, interleaved re/im)    */
/* ------------------------------------------------------------------ */
static inline edo_c64 cget(const double *a, int64_t i) { return a[2 * i] + I * a[2 * i + 1]; }

/* impHloc(ilat,jlat,ispin,jspin,iorb,jorb), 1-based (ED_VARS_GLOBAL.f90:119) */
static edo_c64 imphloc_at(const edo_ctx *c, int ilat, int jlat, int is, int js, int iorb, int jorb) {
  const int L = c->m.nlat, S = c->m.nspin, O = c->m.norb;
  int64_t idx = (ilat - 1) + (int64_t)L * ((jlat - 1) + (int64_t)L * ((is - 1) + (int64_t)S * ((js - 1) + (int64_t)S * ((iorb - 1) + (int64_t)O * (jorb - 1)))));
  return cget(c->imphloc, idx);
}
/* Hbath_reconstructed(ilat,jlat,ispin,jspin,iorb,jorb,ibath) = Hbath_build(lambda_ibath)
 * (ED_BATH/hbath_setup.f90:240-250; extraction ED_HAMILTONIAN_SPARSE_HxV.f90:63-75) */
static edo_c64 hbath_at(const edo_ctx *c, int ilat, int jlat, int is, int js, int iorb, int jorb, int ib) {
  const int L = c->m.nlat, S = c->m.nspin, O = c->m.norb;
  int64_t blk = (int64_t)L * L * S * S * O * O;
  int64_t idx = (ilat - 1) + (int64_t)L * ((jlat - 1) + (int64_t)L * ((is - 1) + (int64_t)S * ((js - 1) + (int64_t)S * ((iorb - 1) + (int64_t)O * (jorb - 1)))));
  return cget(c->hbath, idx + blk * (ib - 1));
}
/* index_stride_lso (ED_AUX_FUNX.f90:81-87) */
static int index_stride_lso(const edo_ctx *c, int ilat, int ispin, int iorb) {
  return iorb + (ilat - 1) * c->m.norb + (ispin - 1) * c->m.norb * c->m.nlat;
}
/* diag_hybr(ilat,ispin,iorb,ibath)=dmft_bath%item(ibath)%v(index_stride_lso) */
static double diag_hybr_at(const edo_ctx *c, int ilat, int ispin, int iorb, int ib) {
  return c->vbath[(index_stride_lso(c, ilat, ispin, iorb) - 1) + (int64_t)c->nlso * (ib - 1)];
}
static double bath_diag_at(const edo_ctx *c, int ilat, int ispin, int iorb, int ib) {
  return creal(hbath_at(c, ilat, ilat, ispin, ispin, iorb, iorb, ib));
}

/* ------------------------------------------------------------------ */
/* dimensions / sectors                                                 */
/* ------------------------------------------------------------------ */
/* binomial (ED_SETUP.f90:1019-1037): floating product, rounded */
int64_t edo_binomial(int32_t n1, int32_t n2) {
  if (n2 < 0) return 0;
  if (n2 == 0) return 1;
  double xh = 1.0;
  for (int i = 1; i <= n2; i++) xh = xh * (double)(n1 + 1 - i) / (double)i;
  return (int64_t)(xh + 0.5);
}
/* get_Sector (ED_SETUP.f90:446-457) with indices=[nup,ndw], Ns_Ud=1 */
int32_t edo_get_sector(int32_t ns, int32_t nup, int32_t ndw) { return 1 + nup * (ns + 1) + ndw; }
/* get_Nup / get_Ndw (ED_SETUP.f90:476-500) */
void edo_get_nup_ndw(int32_t ns, int32_t isector, int32_t *nup, int32_t *ndw) {
  int32_t count = isector - 1;
  int32_t i1 = count % (ns + 1);
  count /= (ns + 1);
  int32_t i2 = count % (ns + 1);
  if (ndw) *ndw = i1;
  if (nup) *nup = i2;
}
/* getDim (ED_SETUP.f90:316-322) -- int64 here, the reference overflows at Ns=18 */
int64_t edo_get_dim(int32_t ns, int32_t isector, int64_t *dimup, int64_t *dimdw) {
  int32_t nup, ndw;
  edo_get_nup_ndw(ns, isector, &nup, &ndw);
  int64_t du = edo_binomial(ns, nup), dd = edo_binomial(ns, ndw);
  if (dimup) *dimup = du;
  if (dimdw) *dimdw = dd;
  return du * dd;
}
/* getCsector / getCDGsector (ED_SETUP.f90:377-418) */
int32_t edo_get_c_sector(int32_t ns, int32_t ispin, int32_t isector) {
  int32_t nup, ndw;
  edo_get_nup_ndw(ns, isector, &nup, &ndw);
  if (ispin == 1) nup--; else ndw--;
  if (nup < 0 || ndw < 0) return 0;
  return edo_get_sector(ns, nup, ndw);
}
int32_t edo_get_cdg_sector(int32_t ns, int32_t ispin, int32_t isector) {
  int32_t nup, ndw;
  edo_get_nup_ndw(ns, isector, &nup, &ndw);
  if (ispin == 1) nup++; else ndw++;
  if (nup > ns || ndw > ns) return 0;
  return edo_get_sector(ns, nup, ndw);
}

/* ------------------------------------------------------------------ */
/* Fock maps and operators                                              */
/* ------------------------------------------------------------------ */
/* build_sector loop (ED_SETUP.f90:749-769): scan 0..2^Ns-1, keep popcnt==n */
int64_t edo_build_sector_map(int32_t ns, int32_t n, int32_t *map) {
  int64_t imap = 0;
  for (int64_t s = 0; s < ((int64_t)1 << ns); s++) {
    if (__builtin_popcountll((unsigned long long)s) != n) continue;
    if (map) map[imap] = (int32_t)s;
    imap++;
  }
  return imap;
}
/* c (ED_SETUP.f90:807-819) */
int32_t edo_c(int32_t pos, int32_t in, int32_t *out, double *fsgn) {
  if (!((in >> (pos - 1)) & 1)) FAIL("C error: C_i|...0_i...>");
  double s = 1.0;
  for (int l = 1; l <= pos - 1; l++)
    if ((in >> (l - 1)) & 1) s = -s;
  *fsgn = s;
  *out = in & ~(1 << (pos - 1));
  return 0;
}
/* cdg (ED_SETUP.f90:821-833) */
int32_t edo_cdg(int32_t pos, int32_t in, int32_t *out, double *fsgn) {
  if ((in >> (pos - 1)) & 1) FAIL("C^+ error: C^+_i|...1_i...>");
  double s = 1.0;
  for (int l = 1; l <= pos - 1; l++)
    if ((in >> (l - 1)) & 1) s = -s;
  *fsgn = s;
  *out = in | (1 << (pos - 1));
  return 0;
}
/* binary_search (ED_SETUP.f90:1044-1061): recursive bisection, 1-based, 0 = not found */
static int32_t bsearch_rec(const int32_t *a, int32_t n, int32_t value) {
  if (n == 0) return 0;
  int32_t mid = n / 2 + 1; /* 1-based */
  if (a[mid - 1] > value) return bsearch_rec(a, mid - 1, value);
  if (a[mid - 1] < value) {
    int32_t r = bsearch_rec(a + mid, n - mid, value);
    return r ? mid + r : 0;
  }
  return mid;
}
int32_t edo_binary_search(const int32_t *a, int32_t n, int32_t value) { return bsearch_rec(a, n, value); }
/* imp_state_index (ED_SETUP.f90:563-568) */
int32_t edo_imp_state_index(const edo_ctx *c, int32_t ilat, int32_t iorb) { return iorb + (ilat - 1) * c->m.norb; }
/* getBathStride (ED_SETUP.f90:367-375) */
int32_t edo_get_bath_stride(const edo_ctx *c, int32_t ilat, int32_t iorb, int32_t ibath) {
  return c->m.nlat * c->m.norb + edo_imp_state_index(c, ilat, iorb) + (ibath - 1) * c->m.nlat * c->m.norb;
}

/* ED_SPARSE_MAP via build_sector(...,itrace=.true.) (ED_SETUP.f90:757-759, ED_SPARSE_MAP.f90:101-121) */
int64_t edo_build_sparse_map(const edo_ctx *c, int32_t n, int64_t *rowptr, int32_t *bath_state, int32_t *sector_indx) {
  const int nimp = c->nimp, ns = c->ns;
  const int64_t nimpst = (int64_t)1 << nimp;
  int64_t *cnt = (int64_t *)calloc(nimpst + 1, sizeof(int64_t));
  int64_t tot = 0;
  for (int64_t s = 0; s < ((int64_t)1 << ns); s++) {
    if (__builtin_popcountll((unsigned long long)s) != n) continue;
    cnt[(s & (nimpst - 1)) + 1]++;
    tot++;
  }
  for (int64_t k = 0; k < nimpst; k++) cnt[k + 1] += cnt[k];
  if (rowptr) memcpy(rowptr, cnt, (nimpst + 1) * sizeof(int64_t));
  if (bath_state && sector_indx) {
    int64_t *fill = (int64_t *)calloc(nimpst, sizeof(int64_t));
    int64_t imap = 0;
    for (int64_t s = 0; s < ((int64_t)1 << ns); s++) {
      if (__builtin_popcountll((unsigned long long)s) != n) continue;
      imap++;
      int64_t iimp = s & (nimpst - 1);      /* ibits(s,0,Nimp) */
      int64_t ibath = s >> nimp;            /* ibits(s,Nimp,Nimp*Nbath) */
      int64_t p = cnt[iimp] + fill[iimp]++; /* insertion order = ascending state */
      bath_state[p] = (int32_t)ibath;
      sector_indx[p] = (int32_t)imap;
    }
    free(fill);
  }
  free(cnt);
  return tot;
}
/* sp_return_intersection (ED_SPARSE_MAP.f90:126-157) */
int32_t edo_sparse_map_intersection(const int64_t *rowptr, const int32_t *bath_state, int32_t iimp, int32_t jimp, int32_t *out) {
  int64_t i0 = rowptr[iimp], i1 = rowptr[iimp + 1], j0 = rowptr[jimp], j1 = rowptr[jimp + 1];
  int32_t n = 0;
  if ((i1 - i0) < (j1 - j0)) {
    for (int64_t i = i0; i < i1; i++)
      for (int64_t j = j0; j < j1; j++)
        if (bath_state[j] == bath_state[i]) { if (out) out[n] = bath_state[i]; n++; break; }
  } else {
    for (int64_t j = j0; j < j1; j++)
      for (int64_t i = i0; i < i1; i++)
        if (bath_state[i] == bath_state[j]) { if (out) out[n] = bath_state[j]; n++; break; }
  }
  return n;
}

/* ------------------------------------------------------------------ */
/* sharding (ED_HAMILTONIAN.f90:92-105; mpiQup ED_HAMILTONIAN_DIRECT_HxV.f90:145-146) */
/* ------------------------------------------------------------------ */
void edo_shard_of(int64_t dimup, int64_t dimdw, int32_t P, int32_t rank, edo_shard *s) {
  s->qdw = dimdw / P;
  s->rdw = dimdw % P;
  if (rank < dimdw % P) { s->rdw = 0; s->qdw += 1; }
  s->q = dimup * s->qdw;
  s->r = dimup * s->rdw;
  s->istart = 1 + rank * s->q + s->r;
  s->iend = (rank + 1) * s->q + s->r;
  s->ishift = rank * s->q + s->r;
  s->dw_off = s->ishift / dimup;
  s->qup = dimup / P;
  int64_t rup = dimup % P;
  if (rank < rup) { s->qup += 1; s->up_off = rank * s->qup; }
  else s->up_off = rank * s->qup + rup;
}
int64_t edo_vecdim(int64_t dimup, int64_t dimdw, int32_t P, int32_t rank) {
  int64_t q = dimdw / P;
  if (rank < dimdw % P) q++;
  return dimup * q;
}

/* ------------------------------------------------------------------ */
/* ED_SPARSE_MATRIX container                                           */
/* ------------------------------------------------------------------ */
static void sp_init(sp_mat *s, int64_t nrow, int64_t ncol) { /* sp_init_matrix_csr :127-149 */
  s->nrow = nrow; s->ncol = ncol;
  s->row = (sp_row *)calloc((size_t)(nrow > 0 ? nrow : 1), sizeof(sp_row));
  s->status = 1;
}
static void sp_delete(sp_mat *s) { /* sp_delete_matrix_csr :184-205 */
  if (!s->status) return;
  for (int64_t i = 0; i < s->nrow; i++) { free(s->row[i].vals); free(s->row[i].cols); }
  free(s->row);
  memset(s, 0, sizeof *s);
}
/* sp_insert_element_csr :254-284: existing column -> add, else append (insertion order) */
static int sp_insert(sp_mat *s, edo_c64 value, int64_t i /*1-based local row*/, int64_t j /*1-based col*/) {
  sp_row *r = &s->row[i - 1];
  for (int32_t k = 0; k < r->size; k++)
    if (r->cols[k] == j) { r->vals[k] += value; return 0; }
  if (r->size == r->cap) {
    r->cap = r->cap ? 2 * r->cap : 4;
    r->vals = (edo_c64 *)realloc(r->vals, r->cap * sizeof(edo_c64));
    r->cols = (int64_t *)realloc(r->cols, r->cap * sizeof(int64_t));
  }
  r->vals[r->size] = value;
  r->cols[r->size] = j;
  r->size++;
  if (r->size > s->ncol) FAIL("sp_insert_element_csr ERROR: row%%Size > sparse%%Ncol");
  return 0;
}

/* ------------------------------------------------------------------ */
/* context                                                              */
/* ------------------------------------------------------------------ */
edo_ctx *edo_create(const edo_model *m) {
  edo_ctx *c = (edo_ctx *)calloc(1, sizeof *c);
  c->m = *m;
  c->nimp = m->nlat * m->norb;          /* ed_setup_dimensions ED_SETUP.f90:111-120 */
  c->ns = c->nimp * (m->nbath + 1);
  c->nlso = m->nlat * m->nspin * m->norb;
  if (c->ns > EDO_MAXNS) { snprintf(g_err, sizeof g_err, "Ns too large"); free(c); return NULL; }
  int64_t nh = (int64_t)m->nlat * m->nlat * m->nspin * m->nspin * m->norb * m->norb;
  c->imphloc = (double *)malloc(2 * nh * sizeof(double));
  memcpy(c->imphloc, m->imphloc, 2 * nh * sizeof(double));
  c->hbath = (double *)malloc(2 * nh * m->nbath * sizeof(double));
  memcpy(c->hbath, m->hbath, 2 * nh * m->nbath * sizeof(double));
  c->vbath = (double *)malloc((size_t)c->nlso * m->nbath * sizeof(double));
  memcpy(c->vbath, m->vbath, (size_t)c->nlso * m->nbath * sizeof(double));
  c->m.imphloc = c->imphloc; c->m.hbath = c->hbath; c->m.vbath = c->vbath;
  c->jhflag = (m->norb > 1 && (m->jx != 0.0 || m->jp != 0.0));
  return c;
}
void edo_destroy(edo_ctx *c) {
  if (!c) return;
  if (c->hstatus) edo_delete_hv_sector(c);
  free(c->imphloc); free(c->hbath); free(c->vbath);
  free(c);
}
int32_t edo_ns(const edo_ctx *c) { return c->ns; }
int32_t edo_nsectors(const edo_ctx *c) { return (c->ns + 1) * (c->ns + 1); }
int32_t edo_active_ranks(const edo_ctx *c) { return c->P; }
int64_t edo_sector_dims(const edo_ctx *c, int64_t *dimup, int64_t *dimdw) {
  if (dimup) *dimup = c->dimup;
  if (dimdw) *dimdw = c->dimdw;
  return c->dim;
}

/* ------------------------------------------------------------------ */
/* diagonal element  (sparse/H_local.f90:1-102, direct/HxV_local.f90:1-95) */
/* ------------------------------------------------------------------ */
static edo_c64 local_element(const edo_ctx *c, int32_t mup, int32_t mdw, int quirk) {
  const int Nlat = c->m.nlat, Norb = c->m.norb, Nspin = c->m.nspin, Nbath = c->m.nbath;
  const double *Uloc = c->m.uloc, Ust = c->m.ust, Jh = c->m.jh, xmu = c->m.xmu;
  double nup[16][8], ndw[16][8];
  for (int ilat = 1; ilat <= Nlat; ilat++)
    for (int iorb = 1; iorb <= Norb; iorb++) {
      int p = edo_imp_state_index(c, ilat, iorb);
      nup[ilat][iorb] = (double)((mup >> (p - 1)) & 1);
      ndw[ilat][iorb] = (double)((mdw >> (p - 1)) & 1);
    }
  edo_c64 htmp = 0;
  for (int ilat = 1; ilat <= Nlat; ilat++)
    for (int iorb = 1; iorb <= Norb; iorb++) {
      htmp += imphloc_at(c, ilat, ilat, 1, 1, iorb, iorb) * nup[ilat][iorb];
      htmp += imphloc_at(c, ilat, ilat, Nspin, Nspin, iorb, iorb) * ndw[ilat][iorb];
      htmp -= xmu * (nup[ilat][iorb] + ndw[ilat][iorb]);
    }
  for (int ilat = 1; ilat <= Nlat; ilat++)
    for (int iorb = 1; iorb <= Norb; iorb++) htmp += Uloc[iorb - 1] * nup[ilat][iorb] * ndw[ilat][iorb];
  if (Norb > 1) {
    for (int ilat = 1; ilat <= Nlat; ilat++)
      for (int iorb = 1; iorb <= Norb; iorb++)
        for (int jorb = iorb + 1; jorb <= Norb; jorb++)
          htmp += Ust * (nup[ilat][iorb] * ndw[ilat][jorb] + nup[ilat][jorb] * ndw[ilat][iorb]);
    for (int ilat = 1; ilat <= Nlat; ilat++)
      for (int iorb = 1; iorb <= Norb; iorb++)
        for (int jorb = iorb + 1; jorb <= Norb; jorb++)
          htmp += (Ust - Jh) * (nup[ilat][iorb] * nup[ilat][jorb] + ndw[ilat][iorb] * ndw[ilat][jorb]);
  }
  if (c->m.hfmode) {
    for (int ilat = 1; ilat <= Nlat; ilat++)
      for (int iorb = 1; iorb <= Norb; iorb++)
        htmp += -0.5 * Uloc[iorb - 1] * (nup[ilat][iorb] + ndw[ilat][iorb]) + 0.25 * Uloc[iorb - 1];
    if (Norb > 1)
      for (int ilat = 1; ilat <= Nlat; ilat++)
        for (int iorb = 1; iorb <= Norb; iorb++)
          for (int jorb = iorb + 1; jorb <= Norb; jorb++) {
            htmp += -0.5 * Ust * (nup[ilat][iorb] + ndw[ilat][iorb] + nup[ilat][jorb] + ndw[ilat][jorb]) + 0.25 * Ust;
            htmp += -0.5 * (Ust - Jh) * (nup[ilat][iorb] + ndw[ilat][iorb] + nup[ilat][jorb] + ndw[ilat][jorb]) + 0.25 * (Ust - Jh);
          }
  }
  /* bath: sparse loops ilat=1..size(bath_diag,1)=Nlat (sparse/H_local.f90:85);
   * the direct variants loop ilat=1..size(bath_diag,3)=Norb (direct/HxV_local.f90:83) */
  int ilat_max = quirk ? Norb : Nlat;
  for (int ilat = 1; ilat <= ilat_max; ilat++)
    for (int iorb = 1; iorb <= Norb; iorb++)
      for (int ib = 1; ib <= Nbath; ib++) {
        if (ilat > Nlat) continue; /* Fortran would read out of bounds; only reachable if Norb>Nlat */
        int ialfa = edo_get_bath_stride(c, ilat, iorb, ib);
        htmp += bath_diag_at(c, ilat, 1, iorb, ib) * (double)((mup >> (ialfa - 1)) & 1);
        htmp += bath_diag_at(c, ilat, Nspin, iorb, ib) * (double)((mdw >> (ialfa - 1)) & 1);
      }
  return htmp;
}

/* ------------------------------------------------------------------ */
/* one-spin hop enumeration shared by sparse/H_up|H_dw.f90 and          */
/* direct/HxV_up|HxV_dw.f90: for source state m, call emit(row k2, h*sg1*sg2)
 * in the reference's loop order: cluster, replica, hybridisation.      */
/* ------------------------------------------------------------------ */
typedef void (*hop_cb)(void *ud, int32_t k2, edo_c64 h);
static void spin_hops(const edo_ctx *c, int sp /*1 or Nspin*/, int32_t m, hop_cb emit, void *ud) {
  const int Nlat = c->m.nlat, Norb = c->m.norb, Nbath = c->m.nbath;
  int32_t k1, k2;
  double sg1, sg2;
  /* H_imp off-diagonal (sparse/H_up.f90:8-30) */
  for (int ilat = 1; ilat <= Nlat; ilat++)
    for (int jlat = 1; jlat <= Nlat; jlat++)
      for (int iorb = 1; iorb <= Norb; iorb++)
        for (int jorb = 1; jorb <= Norb; jorb++) {
          int is = edo_imp_state_index(c, ilat, iorb), js = edo_imp_state_index(c, jlat, jorb);
          edo_c64 h = imphloc_at(c, ilat, jlat, sp, sp, iorb, jorb);
          if (h != 0 && ((m >> (js - 1)) & 1) && !((m >> (is - 1)) & 1)) {
            edo_c(js, m, &k1, &sg1);
            edo_cdg(is, k1, &k2, &sg2);
            emit(ud, k2, h * sg1 * sg2);
          }
        }
  /* H_bath inter-orbital hopping (sparse/H_up.f90:32-58) */
  for (int ib = 1; ib <= Nbath; ib++)
    for (int ilat = 1; ilat <= Nlat; ilat++)
      for (int jlat = 1; jlat <= Nlat; jlat++)
        for (int iorb = 1; iorb <= Norb; iorb++)
          for (int jorb = 1; jorb <= Norb; jorb++) {
            int ialfa = edo_get_bath_stride(c, ilat, iorb, ib), ibeta = edo_get_bath_stride(c, jlat, jorb, ib);
            edo_c64 h = hbath_at(c, ilat, jlat, sp, sp, iorb, jorb, ib);
            if (h != 0 && ((m >> (ibeta - 1)) & 1) && !((m >> (ialfa - 1)) & 1)) {
              edo_c(ibeta, m, &k1, &sg1);
              edo_cdg(ialfa, k1, &k2, &sg2);
              emit(ud, k2, h * sg1 * sg2);
            }
          }
  /* H_hyb (sparse/H_up.f90:61-87) */
  for (int ilat = 1; ilat <= Nlat; ilat++)
    for (int iorb = 1; iorb <= Norb; iorb++)
      for (int ib = 1; ib <= Nbath; ib++) {
        int ialfa = edo_get_bath_stride(c, ilat, iorb, ib), is = edo_imp_state_index(c, ilat, iorb);
        double v = diag_hybr_at(c, ilat, sp, iorb, ib);
        if (v != 0.0 && ((m >> (is - 1)) & 1) && !((m >> (ialfa - 1)) & 1)) {
          edo_c(is, m, &k1, &sg1);
          edo_cdg(ialfa, k1, &k2, &sg2);
          emit(ud, k2, v * sg1 * sg2);
        }
        if (v != 0.0 && !((m >> (is - 1)) & 1) && ((m >> (ialfa - 1)) & 1)) {
          edo_c(ialfa, m, &k1, &sg1);
          edo_cdg(is, k1, &k2, &sg2);
          emit(ud, k2, v * sg1 * sg2);
        }
      }
}

/* non-local S-E / P-H row (sparse/H_non_local.f90:4-100, direct/HxV_non_local.f90:4-86) */
typedef void (*nl_cb)(void *ud, int64_t j /*1-based global col*/, edo_c64 h);
static void nonlocal_row(const edo_ctx *c, int64_t iup, int64_t idw /*1-based*/, nl_cb emit, void *ud) {
  const int Nlat = c->m.nlat, Norb = c->m.norb;
  const double Jx = c->m.jx, Jp = c->m.jp;
  int32_t mup = c->map_up[iup - 1], mdw = c->map_dw[idw - 1];
  int32_t k1, k2, k3, k4;
  double sg1, sg2, sg3, sg4;
#define NUP(p) ((mup >> ((p)-1)) & 1)
#define NDW(p) ((mdw >> ((p)-1)) & 1)
  if (c->jhflag && Jx != 0.0)
    for (int ilat = 1; ilat <= Nlat; ilat++)
      for (int iorb = 1; iorb <= Norb; iorb++)
        for (int jorb = 1; jorb <= Norb; jorb++) {
          int is = edo_imp_state_index(c, ilat, iorb), js = edo_imp_state_index(c, ilat, jorb);
          if (iorb != jorb && NUP(js) == 1 && NDW(is) == 1 && NDW(js) == 0 && NUP(is) == 0) {
            edo_c(is, mdw, &k1, &sg1);
            edo_cdg(js, k1, &k2, &sg2);
            int64_t jdw = edo_binary_search(c->map_dw, (int32_t)c->dimdw, k2);
            edo_c(js, mup, &k3, &sg3);
            edo_cdg(is, k3, &k4, &sg4);
            int64_t jup = edo_binary_search(c->map_up, (int32_t)c->dimup, k4);
            emit(ud, jup + (jdw - 1) * c->dimup, Jx * sg1 * sg2 * sg3 * sg4);
          }
        }
  if (c->jhflag && Jp != 0.0)
    for (int ilat = 1; ilat <= Nlat; ilat++)
      for (int iorb = 1; iorb <= Norb; iorb++)
        for (int jorb = 1; jorb <= Norb; jorb++) {
          int is = edo_imp_state_index(c, ilat, iorb), js = edo_imp_state_index(c, ilat, jorb);
          if (NUP(js) == 1 && NDW(js) == 1 && NDW(is) == 0 && NUP(is) == 0) {
            edo_c(js, mdw, &k1, &sg1);
            edo_cdg(is, k1, &k2, &sg2);
            int64_t jdw = edo_binary_search(c->map_dw, (int32_t)c->dimdw, k2);
            edo_c(js, mup, &k3, &sg3);
            edo_cdg(is, k3, &k4, &sg4);
            int64_t jup = edo_binary_search(c->map_up, (int32_t)c->dimup, k4);
            emit(ud, jup + (jdw - 1) * c->dimup, Jp * sg1 * sg2 * sg3 * sg4);
          }
        }
#undef NUP
#undef NDW
}

/* ------------------------------------------------------------------ */
/* ed_buildh_main (ED_HAMILTONIAN_SPARSE_HxV.f90:40-152)                */
/* ------------------------------------------------------------------ */
typedef struct { edo_ctx *c; sp_mat *mat; const int32_t *map; int32_t n; int64_t col; } ins_ud;
static void ins_cb(void *ud_, int32_t k2, edo_c64 h) {
  ins_ud *ud = (ins_ud *)ud_;
  int32_t row = edo_binary_search(ud->map, ud->n, k2);
  sp_insert(ud->mat, h, row, ud->col);
}
typedef struct { sp_mat *mat; int64_t irow_local; } nlins_ud;
static void nlins_cb(void *ud_, int64_t j, edo_c64 h) {
  nlins_ud *ud = (nlins_ud *)ud_;
  sp_insert(ud->mat, h, ud->irow_local, j);
}
static int buildh(edo_ctx *c) {
  const int Nspin = c->m.nspin;
  sp_init(&c->h0dw, c->dimdw, c->dimdw);
  sp_init(&c->h0up, c->dimup, c->dimup);
  for (int r = 0; r < c->P; r++) {
    rank_state *rk = &c->rk[r];
    int64_t nloc = rk->sh.iend - rk->sh.istart + 1;
    rk->d0 = (edo_c64 *)malloc((size_t)(nloc > 0 ? nloc : 1) * sizeof(edo_c64));
    if (c->jhflag) sp_init(&rk->h0nd, nloc, c->dim);
    /* H_local.f90: do i=MpiIstart,MpiIend ... sp_insert_element(spH0d,htmp,i,i) */
#pragma omp parallel for schedule(static)
    for (int64_t i = rk->sh.istart; i <= rk->sh.iend; i++) {
      int64_t iup = i % c->dimup; if (iup == 0) iup = c->dimup;
      int64_t idw = (i - 1) / c->dimup + 1;
      rk->d0[i - rk->sh.ishift - 1] = local_element(c, c->map_up[iup - 1], c->map_dw[idw - 1], 0);
    }
    if (c->jhflag)
      for (int64_t i = rk->sh.istart; i <= rk->sh.iend; i++) {
        int64_t iup = i % c->dimup; if (iup == 0) iup = c->dimup;
        int64_t idw = (i - 1) / c->dimup + 1;
        nlins_ud ud = {&rk->h0nd, i - rk->sh.ishift};
        nonlocal_row(c, iup, idw, nlins_cb, &ud);
      }
  }
  /* H_up.f90: do jup=1,DimUp ... sp_insert_element(spH0ups(1),htmp,iup,jup) */
  for (int64_t jup = 1; jup <= c->dimup; jup++) {
    ins_ud ud = {c, &c->h0up, c->map_up, (int32_t)c->dimup, jup};
    spin_hops(c, 1, c->map_up[jup - 1], ins_cb, &ud);
  }
  for (int64_t jdw = 1; jdw <= c->dimdw; jdw++) {
    ins_ud ud = {c, &c->h0dw, c->map_dw, (int32_t)c->dimdw, jdw};
    spin_hops(c, Nspin, c->map_dw[jdw - 1], ins_cb, &ud);
  }
  return 0;
}

/* ------------------------------------------------------------------ */
/* build_Hv_sector / delete_Hv_sector (ED_HAMILTONIAN.f90:39-190)       */
/* ------------------------------------------------------------------ */
int32_t edo_build_hv_sector(edo_ctx *c, int32_t isector, int32_t kind, int32_t P, int32_t quirk) {
  if (c->hstatus) FAIL("sp_init_matrix: alreay allocate can not init");
  if (isector < 1 || isector > edo_nsectors(c)) FAIL("bad sector");
  if (P < 1) P = 1;
  if ((kind == EDO_SPARSE_SERIAL || kind == EDO_DIRECT_SERIAL)) P = 1;
  c->hsector = isector; c->kind = kind; c->quirk = quirk;
  int32_t nup, ndw;
  edo_get_nup_ndw(c->ns, isector, &nup, &ndw);
  c->dim = edo_get_dim(c->ns, isector, &c->dimup, &c->dimdw);
  c->map_up = (int32_t *)malloc(c->dimup * sizeof(int32_t));
  c->map_dw = (int32_t *)malloc(c->dimdw * sizeof(int32_t));
  edo_build_sector_map(c->ns, nup, c->map_up);
  edo_build_sector_map(c->ns, ndw, c->map_dw);
  if (c->dimdw < P) P = (int32_t)c->dimdw; /* ED_HAMILTONIAN.f90:62-90 */
  c->P = P;
  c->rk = (rank_state *)calloc(P, sizeof(rank_state));
  for (int r = 0; r < P; r++) edo_shard_of(c->dimup, c->dimdw, P, r, &c->rk[r].sh);
  c->hstatus = 1;
  if (kind == EDO_SPARSE_SERIAL || kind == EDO_SPARSE_MPI) return buildh(c);
  return 0;
}
int32_t edo_delete_hv_sector(edo_ctx *c) {
  if (!c->hstatus) return 0;
  free(c->map_up); free(c->map_dw);
  c->map_up = c->map_dw = NULL;
  for (int r = 0; r < c->P; r++) { free(c->rk[r].d0); c->rk[r].d0 = NULL; sp_delete(&c->rk[r].h0nd); }
  sp_delete(&c->h0up); sp_delete(&c->h0dw);
  free(c->rk); c->rk = NULL;
  c->hsector = 0; c->hstatus = 0;
  return 0;
}

/* ------------------------------------------------------------------ */
/* vector_transpose_MPI (ED_HAMILTONIAN_COMMON.f90:30-101) simulated    */
/* ------------------------------------------------------------------ */
static void split_of(int64_t n, int P, int r, int64_t *q, int64_t *off) {
  int64_t qq = n / P, rr = n % P;
  if (r < rr) { *q = qq + 1; *off = r * (qq + 1); }
  else { *q = qq; *off = r * qq + rr; }
}
/* a_r(nrow,qcol_r) column blocks  ->  b_s(ncol,qrow_s).  Net effect of the per-column
 * MPI_AllToAllV (send row-chunk s of column j to rank s; receive into [qrow,ncol]) plus
 * local_transpose: b_s(jglobal, i-rowoff_s) = a_r(i, jglobal-coloff_r). */
int32_t edo_vector_transpose_sim(int32_t P, int64_t nrow, int64_t ncol, const edo_c64 *a, edo_c64 *b) {
#pragma omp parallel for schedule(static)
  for (int s = 0; s < P; s++) {
    int64_t qrow, rowoff;
    split_of(nrow, P, s, &qrow, &rowoff);
    edo_c64 *bs = b + rowoff * ncol;
    for (int r = 0; r < P; r++) {
      int64_t qcol, coloff;
      split_of(ncol, P, r, &qcol, &coloff);
      const edo_c64 *ar = a + coloff * nrow;
      for (int64_t i = 0; i < qrow; i++)
        for (int64_t j = 0; j < qcol; j++) bs[(coloff + j) + i * ncol] = ar[(rowoff + i) + j * nrow];
    }
  }
  return 0;
}

/* ------------------------------------------------------------------ */
/* mat-vecs                                                             */
/* ------------------------------------------------------------------ */
/* spMatVec_main (ED_HAMILTONIAN_SPARSE_HxV.f90:167-227): order diag -> DW -> UP -> ND */
static void spmatvec_serial(edo_ctx *c, const edo_c64 *v, edo_c64 *hv) {
  const int64_t DimUp = c->dimup, DimDw = c->dimdw, N = c->dim;
  rank_state *rk = &c->rk[0];
  for (int64_t i = 0; i < N; i++) hv[i] = 0;
  for (int64_t i = 0; i < N; i++) hv[i] += rk->d0[i] * v[i]; /* Hv(i) += spH0d%row(i)%vals(1)*v(cols(1)), cols(1)=i */
  for (int64_t iup = 1; iup <= DimUp; iup++)
    for (int64_t idw = 1; idw <= DimDw; idw++) {
      int64_t i = iup + (idw - 1) * DimUp;
      sp_row *r = &c->h0dw.row[idw - 1];
      for (int32_t jj = 0; jj < r->size; jj++) hv[i - 1] += r->vals[jj] * v[iup + (r->cols[jj] - 1) * DimUp - 1];
    }
  for (int64_t idw = 1; idw <= DimDw; idw++)
    for (int64_t iup = 1; iup <= DimUp; iup++) {
      int64_t i = iup + (idw - 1) * DimUp;
      sp_row *r = &c->h0up.row[iup - 1];
      for (int32_t jj = 0; jj < r->size; jj++) hv[i - 1] += r->vals[jj] * v[r->cols[jj] + (idw - 1) * DimUp - 1];
    }
  if (c->jhflag)
    for (int64_t i = 0; i < N; i++)
      for (int32_t j = 0; j < rk->h0nd.row[i].size; j++) hv[i] += rk->h0nd.row[i].vals[j] * v[rk->h0nd.row[i].cols[j] - 1];
}

/* spMatVec_mpi_main (:230-315): diag -> UP -> transpose -> DW on vt -> transpose back -> add -> ND(allgather) */
static void spmatvec_mpi(edo_ctx *c, const edo_c64 *v, edo_c64 *hv) {
  const int64_t DimUp = c->dimup, DimDw = c->dimdw, N = c->dim;
  const int P = c->P;
  edo_c64 *vt = (edo_c64 *)malloc(N * sizeof(edo_c64)), *hvt = (edo_c64 *)malloc(N * sizeof(edo_c64));
#pragma omp parallel for schedule(static)
  for (int rnk = 0; rnk < P; rnk++) {
    rank_state *rk = &c->rk[rnk];
    const edo_c64 *vl = v + rk->sh.ishift;
    edo_c64 *hl = hv + rk->sh.ishift;
    int64_t nloc = rk->sh.q;
    for (int64_t i = 0; i < nloc; i++) hl[i] = rk->d0[i] * vl[i]; /* uses v(i), local index (:251-255) */
    for (int64_t idw = 1; idw <= rk->sh.qdw; idw++)
      for (int64_t iup = 1; iup <= DimUp; iup++) {
        int64_t i = iup + (idw - 1) * DimUp;
        sp_row *r = &c->h0up.row[iup - 1];
        for (int32_t jj = 0; jj < r->size; jj++) hl[i - 1] += r->vals[jj] * vl[r->cols[jj] + (idw - 1) * DimUp - 1];
      }
  }
  edo_vector_transpose_sim(P, DimUp, DimDw, v, vt);
#pragma omp parallel for schedule(static)
  for (int rnk = 0; rnk < P; rnk++) {
    rank_state *rk = &c->rk[rnk];
    const edo_c64 *vtl = vt + rk->sh.up_off * DimDw;
    edo_c64 *hvtl = hvt + rk->sh.up_off * DimDw;
    for (int64_t i = 0; i < rk->sh.qup * DimDw; i++) hvtl[i] = 0;
    for (int64_t idw = 1; idw <= rk->sh.qup; idw++)   /* transposed order: column-wise DW <--> UP */
      for (int64_t iup = 1; iup <= DimDw; iup++) {
        int64_t i = iup + (idw - 1) * DimDw;
        sp_row *r = &c->h0dw.row[iup - 1];
        for (int32_t jj = 0; jj < r->size; jj++) hvtl[i - 1] += r->vals[jj] * vtl[r->cols[jj] + (idw - 1) * DimDw - 1];
      }
  }
  edo_vector_transpose_sim(P, DimDw, DimUp, hvt, vt);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < N; i++) hv[i] += vt[i];
  if (c->jhflag) { /* allgather_vector_MPI -> v is already the gathered vector */
#pragma omp parallel for schedule(static)
    for (int rnk = 0; rnk < P; rnk++) {
      rank_state *rk = &c->rk[rnk];
      edo_c64 *hl = hv + rk->sh.ishift;
      for (int64_t i = 0; i < rk->sh.q; i++)
        for (int32_t j = 0; j < rk->h0nd.row[i].size; j++) hl[i] += rk->h0nd.row[i].vals[j] * v[rk->h0nd.row[i].cols[j] - 1];
    }
  }
  free(vt); free(hvt);
}

/* push-style accumulation used by the direct variants */
typedef struct { const int32_t *map; int32_t n; edo_c64 *hcol; edo_c64 vin; int64_t stride; } push_ud;
static void push_cb(void *ud_, int32_t k2, edo_c64 h) {
  push_ud *ud = (push_ud *)ud_;
  int32_t row = edo_binary_search(ud->map, ud->n, k2);
  ud->hcol[(int64_t)(row - 1) * ud->stride] += h * ud->vin;
}
typedef struct { edo_c64 *acc; const edo_c64 *v; } nlmv_ud;
static void nlmv_cb(void *ud_, int64_t j, edo_c64 h) {
  nlmv_ud *ud = (nlmv_ud *)ud_;
  *ud->acc += h * ud->v[j - 1];
}

/* directMatVec_main (ED_HAMILTONIAN_DIRECT_HxV.f90:37-90 + direct/ files) */
static void directmatvec_serial(edo_ctx *c, const edo_c64 *v, edo_c64 *hv) {
  const int64_t DimUp = c->dimup, DimDw = c->dimdw, N = c->dim;
  const int Nspin = c->m.nspin;
  for (int64_t i = 0; i < N; i++) hv[i] = 0;
  for (int64_t i = 1; i <= N; i++) { /* HxV_local.f90 */
    int64_t iup = i % DimUp; if (iup == 0) iup = DimUp;
    int64_t idw = (i - 1) / DimUp + 1;
    hv[i - 1] += local_element(c, c->map_up[iup - 1], c->map_dw[idw - 1], c->quirk) * v[i - 1];
  }
  for (int64_t jdw = 1; jdw <= DimDw; jdw++) /* HxV_up.f90 */
    for (int64_t jup = 1; jup <= DimUp; jup++) {
      int64_t j = jup + (jdw - 1) * DimUp;
      push_ud ud = {c->map_up, (int32_t)DimUp, hv + (jdw - 1) * DimUp, v[j - 1], 1};
      spin_hops(c, 1, c->map_up[jup - 1], push_cb, &ud);
    }
  for (int64_t jup = 1; jup <= DimUp; jup++) /* HxV_dw.f90 */
    for (int64_t jdw = 1; jdw <= DimDw; jdw++) {
      int64_t j = jup + (jdw - 1) * DimUp;
      push_ud ud = {c->map_dw, (int32_t)DimDw, hv + (jup - 1), v[j - 1], DimUp};
      spin_hops(c, Nspin, c->map_dw[jdw - 1], push_cb, &ud);
    }
  if (c->jhflag) /* HxV_non_local.f90 */
    for (int64_t i = 1; i <= N; i++) {
      int64_t iup = i % DimUp; if (iup == 0) iup = DimUp;
      int64_t idw = (i - 1) / DimUp + 1;
      nlmv_ud ud = {&hv[i - 1], v};
      nonlocal_row(c, iup, idw, nlmv_cb, &ud);
    }
}

/* directMatVec_MPI_main (:94-171 + direct_mpi/ files) */
static void directmatvec_mpi(edo_ctx *c, const edo_c64 *v, edo_c64 *hv) {
  const int64_t DimUp = c->dimup, DimDw = c->dimdw, N = c->dim;
  const int P = c->P, Nspin = c->m.nspin;
  edo_c64 *vt = (edo_c64 *)malloc(N * sizeof(edo_c64)), *hvt = (edo_c64 *)malloc(N * sizeof(edo_c64));
#pragma omp parallel for schedule(static)
  for (int rnk = 0; rnk < P; rnk++) {
    rank_state *rk = &c->rk[rnk];
    const edo_c64 *vl = v + rk->sh.ishift;
    edo_c64 *hl = hv + rk->sh.ishift;
    for (int64_t i = 1; i <= rk->sh.q; i++) { /* direct_mpi/HxV_local.f90 */
      int64_t ig = i + rk->sh.ishift;
      int64_t iup = ig % DimUp; if (iup == 0) iup = DimUp;
      int64_t idw = (ig - 1) / DimUp + 1;
      hl[i - 1] = local_element(c, c->map_up[iup - 1], c->map_dw[idw - 1], c->quirk) * vl[i - 1];
    }
    for (int64_t jdw = 1; jdw <= rk->sh.qdw; jdw++) /* direct_mpi/HxV_up.f90 */
      for (int64_t jup = 1; jup <= DimUp; jup++) {
        int64_t j = jup + (jdw - 1) * DimUp;
        push_ud ud = {c->map_up, (int32_t)DimUp, hl + (jdw - 1) * DimUp, vl[j - 1], 1};
        spin_hops(c, 1, c->map_up[jup - 1], push_cb, &ud);
      }
  }
  edo_vector_transpose_sim(P, DimUp, DimDw, v, vt);
#pragma omp parallel for schedule(static)
  for (int rnk = 0; rnk < P; rnk++) {
    rank_state *rk = &c->rk[rnk];
    const edo_c64 *vtl = vt + rk->sh.up_off * DimDw;
    edo_c64 *hvtl = hvt + rk->sh.up_off * DimDw;
    for (int64_t i = 0; i < rk->sh.qup * DimDw; i++) hvtl[i] = 0;
    for (int64_t jdw = 1; jdw <= rk->sh.qup; jdw++) /* direct_mpi/HxV_dw.f90 (roles swapped) */
      for (int64_t jup = 1; jup <= DimDw; jup++) {
        int64_t j = jup + (jdw - 1) * DimDw;
        push_ud ud = {c->map_dw, (int32_t)DimDw, hvtl + (jdw - 1) * DimDw, vtl[j - 1], 1};
        spin_hops(c, Nspin, c->map_dw[jup - 1], push_cb, &ud);
      }
  }
  edo_vector_transpose_sim(P, DimDw, DimUp, hvt, vt);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < N; i++) hv[i] += vt[i];
  if (c->jhflag) {
#pragma omp parallel for schedule(static)
    for (int rnk = 0; rnk < P; rnk++) {
      rank_state *rk = &c->rk[rnk];
      edo_c64 *hl = hv + rk->sh.ishift;
      for (int64_t i = 1; i <= rk->sh.q; i++) {
        int64_t ig = i + rk->sh.ishift;
        int64_t iup = ig % DimUp; if (iup == 0) iup = DimUp;
        int64_t idw = (ig - 1) / DimUp + 1;
        nlmv_ud ud = {&hl[i - 1], v};
        nonlocal_row(c, iup, idw, nlmv_cb, &ud);
      }
    }
  }
  free(vt); free(hvt);
}

int32_t edo_hxv(edo_ctx *c, int64_t n, const edo_c64 *v, edo_c64 *hv) {
  if (!c->hstatus) FAIL("directMatVec_cc ERROR: Hsector NOT set");
  if (n != c->dim) FAIL("directMatVec_cc ERROR: Nloc != dim(isector)");
  switch (c->kind) {
  case EDO_SPARSE_SERIAL: spmatvec_serial(c, v, hv); break;
  case EDO_SPARSE_MPI: spmatvec_mpi(c, v, hv); break;
  case EDO_DIRECT_SERIAL: directmatvec_serial(c, v, hv); break;
  case EDO_DIRECT_MPI: directmatvec_mpi(c, v, hv); break;
  default: FAIL("bad kind");
  }
  return 0;
}

/* ------------------------------------------------------------------ */
/* sampled rows against a counter-based vector (see ed_oracle.h)         */
/* ------------------------------------------------------------------ */
static inline uint64_t mix64(uint64_t x) { x ^= x >> 31; x *= 0xD6E8FEB86659FD93ULL; x ^= x >> 32; return x; }
static inline edo_c64 counter_at(int64_t idx, uint64_t seed, double scale) {
  const uint64_t x = mix64(((uint64_t)idx + 1ULL) * 0x9E3779B97F4A7C15ULL + seed * 0xBF58476D1CE4E5B9ULL);
  const uint64_t y = mix64(x * 0x94D049BB133111EBULL + 0x2545F4914F6CDD1DULL);
  const double re = (double)(x >> 11) * 0x1p-53 * 2.0 - 1.0, im = (double)(y >> 11) * 0x1p-53 * 2.0 - 1.0;
  return scale * re + I * (scale * im);
}
void edo_counter_vec(int64_t i0, int64_t n, uint64_t seed, double scale, edo_c64 *out) {
#pragma omp parallel for schedule(static)
  for (int64_t k = 0; k < n; k++) out[k] = counter_at(i0 + k, seed, scale);
}
typedef struct { const edo_ctx *c; edo_c64 acc; int64_t fixed; int up; uint64_t seed; double scale; } pull_ud;
static void pull_cb(void *ud_, int32_t k2, edo_c64 h) {
  pull_ud *ud = (pull_ud *)ud_;
  const edo_ctx *c = ud->c;
  /* k2 = the state c^+_a c_b |m> the reference pushes H(k,i) v(i) to; row i collects conjg(H(k,i)) v(k) */
  int64_t k;
  if (ud->up) k = (int64_t)edo_binary_search(c->map_up, (int32_t)c->dimup, k2) - 1 + ud->fixed * c->dimup;
  else k = ud->fixed + ((int64_t)edo_binary_search(c->map_dw, (int32_t)c->dimdw, k2) - 1) * c->dimup;
  ud->acc += conj(h) * counter_at(k, ud->seed, ud->scale);
}
typedef struct { edo_c64 acc; uint64_t seed; double scale; } nlpull_ud;
static void nlpull_cb(void *ud_, int64_t j, edo_c64 h) {
  nlpull_ud *ud = (nlpull_ud *)ud_;
  ud->acc += h * counter_at(j - 1, ud->seed, ud->scale); /* nonlocal_row is already row-wise (H_non_local.f90) */
}
int32_t edo_hxv_rows_counter(edo_ctx *c, int64_t nrows, const int64_t *rows, uint64_t seed, double scale, edo_c64 *out) {
  if (!c->hstatus) FAIL("directMatVec_cc ERROR: Hsector NOT set");
  const int Nspin = c->m.nspin;
  for (int64_t r = 0; r < nrows; r++)
    if (rows[r] < 0 || rows[r] >= c->dim) FAIL("hxv_rows_counter: row out of range");
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t r = 0; r < nrows; r++) {
    const int64_t i = rows[r], iup = i % c->dimup, idw = i / c->dimup; /* 0-based */
    edo_c64 acc = local_element(c, c->map_up[iup], c->map_dw[idw], c->quirk) * counter_at(i, seed, scale);
    pull_ud uu = {c, 0, idw, 1, seed, scale};
    spin_hops(c, 1, c->map_up[iup], pull_cb, &uu);
    pull_ud ud = {c, 0, iup, 0, seed, scale};
    spin_hops(c, Nspin, c->map_dw[idw], pull_cb, &ud);
    acc += uu.acc + ud.acc;
    if (c->jhflag) {
      nlpull_ud nu = {0, seed, scale};
      nonlocal_row(c, iup + 1, idw + 1, nlpull_cb, &nu);
      acc += nu.acc;
    }
    out[r] = acc;
  }
  return 0;
}

/* ------------------------------------------------------------------ */
/* inspection                                                           */
/* ------------------------------------------------------------------ */
int64_t edo_get_csr(const edo_ctx *c, int32_t which, int64_t *rowptr, int32_t *col, edo_c64 *val) {
  const sp_mat *s = which == 1 ? &c->h0up : &c->h0dw;
  if (!s->status) return -1;
  int64_t nnz = 0;
  for (int64_t i = 0; i < s->nrow; i++) {
    if (rowptr) rowptr[i] = nnz;
    for (int32_t k = 0; k < s->row[i].size; k++) {
      if (col) col[nnz] = (int32_t)s->row[i].cols[k];
      if (val) val[nnz] = s->row[i].vals[k];
      nnz++;
    }
  }
  if (rowptr) rowptr[s->nrow] = nnz;
  return nnz;
}
int32_t edo_get_diag(const edo_ctx *c, double *d) {
  if (!c->hstatus) FAIL("no sector");
  for (int64_t i = 1; i <= c->dim; i++) {
    int64_t iup = i % c->dimup; if (iup == 0) iup = c->dimup;
    int64_t idw = (i - 1) / c->dimup + 1;
    int quirk = (c->kind == EDO_DIRECT_SERIAL || c->kind == EDO_DIRECT_MPI) ? c->quirk : 0;
    d[i - 1] = creal(local_element(c, c->map_up[iup - 1], c->map_dw[idw - 1], quirk));
  }
  return 0;
}
int64_t edo_get_nonlocal(const edo_ctx *c, int64_t *rowptr, int64_t *col, edo_c64 *val) {
  if (!c->hstatus || !c->jhflag) return 0;
  int64_t nnz = 0, row = 0;
  for (int r = 0; r < c->P; r++) {
    const sp_mat *s = &c->rk[r].h0nd;
    if (!s->status) return -1;
    for (int64_t i = 0; i < s->nrow; i++, row++) {
      if (rowptr) rowptr[row] = nnz;
      for (int32_t k = 0; k < s->row[i].size; k++) {
        if (col) col[nnz] = s->row[i].cols[k];
        if (val) val[nnz] = s->row[i].vals[k];
        nnz++;
      }
    }
  }
  if (rowptr) rowptr[row] = nnz;
  return nnz;
}
/* Hmat = spH0d (+spH0nd) + kron(Hdw,1_up) + kron(1_dw,Hup)  (ED_HAMILTONIAN_SPARSE_HxV.f90:112-148) */
int32_t edo_dense_hmat(edo_ctx *c, int32_t isector, edo_c64 *hmat) {
  if (edo_build_hv_sector(c, isector, EDO_SPARSE_SERIAL, 1, 0)) return -1;
  const int64_t N = c->dim, DimUp = c->dimup, DimDw = c->dimdw;
  memset(hmat, 0, (size_t)N * N * sizeof(edo_c64));
  rank_state *rk = &c->rk[0];
  for (int64_t i = 0; i < N; i++) {
    hmat[i + i * N] += rk->d0[i];
    if (c->jhflag)
      for (int32_t k = 0; k < rk->h0nd.row[i].size; k++) hmat[i + (rk->h0nd.row[i].cols[k] - 1) * N] += rk->h0nd.row[i].vals[k];
  }
  for (int64_t idw = 0; idw < DimDw; idw++) /* kron(Hdw, eye(DimUp)) */
    for (int32_t k = 0; k < c->h0dw.row[idw].size; k++) {
      int64_t jdw = c->h0dw.row[idw].cols[k] - 1;
      for (int64_t iup = 0; iup < DimUp; iup++) hmat[(iup + idw * DimUp) + (iup + jdw * DimUp) * N] += c->h0dw.row[idw].vals[k];
    }
  for (int64_t iup = 0; iup < DimUp; iup++) /* kron(eye(DimDw), Hup) */
    for (int32_t k = 0; k < c->h0up.row[iup].size; k++) {
      int64_t jup = c->h0up.row[iup].cols[k] - 1;
      for (int64_t idw = 0; idw < DimDw; idw++) hmat[(iup + idw * DimUp) + (jup + idw * DimUp) * N] += c->h0up.row[iup].vals[k];
    }
  return edo_delete_hv_sector(c);
}

/* ------------------------------------------------------------------ */
/* Krylov: SciFortran SF_SP_LINALG restated             */
/* ------------------------------------------------------------------ */
static edo_c64 zdot(int64_t n, const edo_c64 *a, const edo_c64 *b) { /* dot_product(a,b)=sum(conjg(a)*b) */
  double re = 0, im = 0;
#pragma omp parallel for reduction(+ : re, im) schedule(static)
  for (int64_t i = 0; i < n; i++) {
    edo_c64 t = conj(a[i]) * b[i];
    re += creal(t); im += cimag(t);
  }
  return re + I * im;
}
/* lanczos_iteration: one step of the 3-term recurrence */
static int lanczos_iteration(edo_ctx *c, int64_t n, int iter, edo_c64 *vin, edo_c64 *vout, edo_c64 *tmp, double *alfa, double *beta) {
  if (iter == 1) {
    double norm = sqrt(creal(zdot(n, vin, vin)));
    if (norm == 0.0) FAIL("LANCZOS_ITERATION: norm=0");
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) vin[i] /= norm;
  } else {
    double b = *beta;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
      edo_c64 t = vin[i];
      vin[i] = vout[i] / b;
      vout[i] = -b * t;
    }
  }
  if (edo_hxv(c, n, vin, tmp)) return -1;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; i++) vout[i] += tmp[i];
  double a = creal(zdot(n, vin, vout));
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; i++) vout[i] -= a * vin[i];
  *alfa = a;
  *beta = sqrt(creal(zdot(n, vout, vout)));
  return 0;
}
/* sp_lanc_tridiag */
int32_t edo_lanc_tridiag(edo_ctx *c, int64_t n, edo_c64 *vin, int32_t nitermax, double *alanc, double *blanc, double threshold) {
  edo_c64 *vout = (edo_c64 *)calloc(n, sizeof(edo_c64)), *tmp = (edo_c64 *)calloc(n, sizeof(edo_c64));
  double a = 0, b = 0;
  int32_t done = 0;
  for (int i = 0; i < nitermax; i++) { alanc[i] = 0; blanc[i] = 0; }
  for (int iter = 1; iter <= nitermax; iter++) {
    if (lanczos_iteration(c, n, iter, vin, vout, tmp, &a, &b)) { free(vout); free(tmp); return -1; }
    alanc[iter - 1] = a;
    done = iter;
    if (fabs(b) < threshold) break;
    if (iter < nitermax) blanc[iter] = b;
  }
  free(vout); free(tmp);
  return done;
}

/* symmetric tridiagonal eigensolver: implicit-shift QL with eigenvectors, ascending order.
 * d[0..n-1] diagonal, e[1..n-1] sub-diagonal (e[0] ignored), z column-major, on entry identity
 * (SF_LINALG eigh(diag,subdiag,Ev) wraps LAPACK dstev; any correct solver is equivalent). */
int32_t edo_tridiag_eigh(int32_t n, double *d, double *e_in, double *z) {
  double *e = (double *)calloc(n + 1, sizeof(double));
  for (int i = 1; i < n; i++) e[i - 1] = e_in[i];
  e[n - 1] = 0.0;
  if (z) for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) z[i + (size_t)j * n] = (i == j);
  for (int l = 0; l < n; l++) {
    int iter = 0, m;
    do {
      for (m = l; m < n - 1; m++) {
        double dd = fabs(d[m]) + fabs(d[m + 1]);
        if (fabs(e[m]) <= 2.220446049250313e-16 * dd) break;
      }
      if (m != l) {
        if (iter++ == 200) { free(e); FAIL("tridiag_eigh: too many iterations"); }
        double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
        double r = hypot(g, 1.0);
        g = d[m] - d[l] + e[l] / (g + (g >= 0 ? fabs(r) : -fabs(r)));
        double s = 1.0, cc = 1.0, p = 0.0;
        int i;
        for (i = m - 1; i >= l; i--) {
          double f = s * e[i], b = cc * e[i];
          e[i + 1] = (r = hypot(f, g));
          if (r == 0.0) { d[i + 1] -= p; e[m] = 0.0; break; }
          s = f / r; cc = g / r;
          g = d[i + 1] - p;
          r = (d[i] - g) * s + 2.0 * cc * b;
          d[i + 1] = g + (p = s * r);
          g = cc * r - b;
          if (z) for (int k = 0; k < n; k++) {
            double f2 = z[k + (size_t)(i + 1) * n];
            z[k + (size_t)(i + 1) * n] = s * z[k + (size_t)i * n] + cc * f2;
            z[k + (size_t)i * n] = cc * z[k + (size_t)i * n] - s * f2;
          }
        }
        if (r == 0.0 && i >= l) continue;
        d[l] -= p; e[l] = g; e[m] = 0.0;
      }
    } while (m != l);
  }
  /* sort ascending */
  for (int i = 0; i < n - 1; i++) {
    int k = i; double p = d[i];
    for (int j = i + 1; j < n; j++) if (d[j] < p) { k = j; p = d[j]; }
    if (k != i) {
      d[k] = d[i]; d[i] = p;
      if (z) for (int j = 0; j < n; j++) { double t = z[j + (size_t)i * n]; z[j + (size_t)i * n] = z[j + (size_t)k * n]; z[j + (size_t)k * n] = t; }
    }
  }
  free(e);
  return 0;
}

/* sp_lanc_eigh: Lanczos loop with per-step tridiagonal solve, stop when |dE0|<=threshold
 * once nlanc>=ncheck; second pass accumulates the eigenvector. */
int32_t edo_lanc_eigh(edo_ctx *c, int64_t n, double *egs, edo_c64 *vect, int32_t nitermax, double threshold, int32_t ncheck, int32_t *niter_out, double *alanc_out, double *blanc_out) {
  if (nitermax > n) nitermax = (int32_t)n;
  double nrm = creal(zdot(n, vect, vect));
  if (nrm == 0.0) { /* start vector unpinned in SciFortran: constant 1/sqrt(N) here */
    for (int64_t i = 0; i < n; i++) vect[i] = 1.0 / sqrt((double)n);
  }
  edo_c64 *vin = (edo_c64 *)malloc(n * sizeof(edo_c64)), *vout = (edo_c64 *)calloc(n, sizeof(edo_c64)), *tmp = (edo_c64 *)calloc(n, sizeof(edo_c64));
  memcpy(vin, vect, n * sizeof(edo_c64));
  double *alanc = (double *)calloc(nitermax, sizeof(double)), *blanc = (double *)calloc(nitermax, sizeof(double));
  double *diag = (double *)malloc(nitermax * sizeof(double)), *sub = (double *)malloc(nitermax * sizeof(double));
  double *Z = (double *)malloc((size_t)nitermax * nitermax * sizeof(double));
  double a = 0, b = 0, esave = 0, e0 = 0;
  int nlanc = 0;
  for (int iter = 1; iter <= nitermax; iter++) {
    if (lanczos_iteration(c, n, iter, vin, vout, tmp, &a, &b)) return -1;
    if (fabs(b) < threshold) { /* invariant subspace: keep alpha, stop */
      nlanc++; alanc[iter - 1] = a; break;
    }
    nlanc++;
    alanc[iter - 1] = a;
    if (iter < nitermax) blanc[iter] = b;
    memcpy(diag, alanc, nlanc * sizeof(double));
    memcpy(sub, blanc, nlanc * sizeof(double));
    edo_tridiag_eigh(nlanc, diag, sub, Z);
    e0 = diag[0];
    if (nlanc >= ncheck) {
      double diff = fabs(diag[0] - esave);
      if (diff <= threshold) break;
    }
    esave = diag[0];
  }
  memcpy(diag, alanc, nlanc * sizeof(double));
  memcpy(sub, blanc, nlanc * sizeof(double));
  edo_tridiag_eigh(nlanc, diag, sub, Z);
  e0 = diag[0];
  /* second pass: vect = sum_iter vin_iter * Z(iter,1) */
  memcpy(vin, vect, n * sizeof(edo_c64));
  memset(vout, 0, n * sizeof(edo_c64));
  memset(vect, 0, n * sizeof(edo_c64));
  for (int iter = 1; iter <= nlanc; iter++) {
    if (lanczos_iteration(c, n, iter, vin, vout, tmp, &a, &b)) return -1;
    double z = Z[iter - 1];
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) vect[i] += vin[i] * z;
  }
  double norm = sqrt(creal(zdot(n, vect, vect)));
  for (int64_t i = 0; i < n; i++) vect[i] /= norm;
  *egs = e0;
  if (niter_out) *niter_out = nlanc;
  if (alanc_out) memcpy(alanc_out, alanc, nlanc * sizeof(double));
  if (blanc_out) memcpy(blanc_out, blanc, nlanc * sizeof(double));
  free(vin); free(vout); free(tmp); free(alanc); free(blanc); free(diag); free(sub); free(Z);
  return 0;
}

/* ------------------------------------------------------------------ */
/* Green's function pieces                                              */
/* ------------------------------------------------------------------ */
/* vvinit loops of ED_GF_NORMAL.f90:180-194 (cdg), :244-258 (c), and the mixed channels :590-620 */
int32_t edo_apply_op(int32_t ns, int32_t isector, int32_t iop, int32_t ispin, int32_t nops, const int32_t *pos, const edo_c64 *coef, const edo_c64 *state, edo_c64 *out, int32_t *jsector_out) {
  int32_t jsector = iop > 0 ? edo_get_cdg_sector(ns, ispin, isector) : edo_get_c_sector(ns, ispin, isector);
  if (jsector_out) *jsector_out = jsector;
  if (jsector == 0) return 0;
  int32_t inup, indw, jnup, jndw;
  edo_get_nup_ndw(ns, isector, &inup, &indw);
  edo_get_nup_ndw(ns, jsector, &jnup, &jndw);
  int64_t idimup, idimdw, jdimup, jdimdw;
  int64_t idim = edo_get_dim(ns, isector, &idimup, &idimdw);
  int64_t jdim = edo_get_dim(ns, jsector, &jdimup, &jdimdw);
  int32_t *hiu = (int32_t *)malloc(idimup * 4), *hid = (int32_t *)malloc(idimdw * 4);
  int32_t *hju = (int32_t *)malloc(jdimup * 4), *hjd = (int32_t *)malloc(jdimdw * 4);
  edo_build_sector_map(ns, inup, hiu); edo_build_sector_map(ns, indw, hid);
  edo_build_sector_map(ns, jnup, hju); edo_build_sector_map(ns, jndw, hjd);
  for (int64_t j = 0; j < jdim; j++) out[j] = 0;
  for (int k = 0; k < nops; k++) {
    for (int64_t i = 1; i <= idim; i++) {
      int64_t iu = (i - 1) % idimup + 1, id = (i - 1) / idimup + 1; /* state2indices */
      int32_t s = ispin == 1 ? hiu[iu - 1] : hid[id - 1];
      int occ = (s >> (pos[k] - 1)) & 1;
      int32_t r; double sgn;
      if (iop > 0) { if (occ != 0) continue; edo_cdg(pos[k], s, &r, &sgn); }
      else { if (occ != 1) continue; edo_c(pos[k], s, &r, &sgn); }
      int64_t ju = iu, jd = id;
      if (ispin == 1) ju = edo_binary_search(hju, (int32_t)jdimup, r);
      else jd = edo_binary_search(hjd, (int32_t)jdimdw, r);
      int64_t j = ju + (jd - 1) * jdimup; /* indices2state */
      if (k == 0) out[j - 1] = coef[k] * sgn * state[i - 1];
      else out[j - 1] += coef[k] * sgn * state[i - 1];
    }
  }
  free(hiu); free(hid); free(hju); free(hjd);
  return 0;
}
/* add_to_lanczos_gf_normal (ED_GF_NORMAL.f90:915-975), T=0 branch: pesoBZ=vnorm2/zeta_function */
int32_t edo_add_to_lanczos_gf(edo_c64 vnorm2, double ei, int32_t nlanc, const double *alanc, const double *blanc, int32_t isign, double zeta, int32_t lmats, const double *wm, edo_c64 *g, double *poles, edo_c64 *weights) {
  double *diag = (double *)malloc(nlanc * sizeof(double)), *sub = (double *)malloc(nlanc * sizeof(double));
  double *Z = (double *)malloc((size_t)nlanc * nlanc * sizeof(double));
  memcpy(diag, alanc, nlanc * sizeof(double));
  memcpy(sub, blanc, nlanc * sizeof(double));
  if (edo_tridiag_eigh(nlanc, diag, sub, Z)) return -1;
  edo_c64 pesoBZ = vnorm2 / zeta;
  for (int j = 0; j < nlanc; j++) {
    double de = diag[j] - ei;
    edo_c64 peso = pesoBZ * Z[(size_t)j * nlanc] * Z[(size_t)j * nlanc];
    if (poles) poles[j] = isign * de;
    if (weights) weights[j] = peso;
    for (int i = 0; i < lmats; i++) g[i] += peso / (I * wm[i] - isign * de);
  }
  free(diag); free(sub); free(Z);
  return 0;
}

/* the same routine with its two other branches: the Boltzmann weight of finite-temperature runs (:930-936) and the
 * real-axis accumulation impGreal (:968-971, iw = dcmplx(wr(i), eps)); gr may be NULL (lreal = 0) */
int32_t edo_add_to_lanczos_gf_full(edo_c64 vnorm2, double ei, double egs, int32_t finite_t, double beta, int32_t nlanc,
                                   const double *alanc, const double *blanc, int32_t isign, double zeta, int32_t lmats,
                                   const double *wm, edo_c64 *gm, int32_t lreal, const double *wr, double eps, edo_c64 *gr,
                                   double *poles, edo_c64 *weights) {
  double *diag = (double *)malloc(nlanc * sizeof(double)), *sub = (double *)malloc(nlanc * sizeof(double));
  double *Z = (double *)malloc((size_t)nlanc * nlanc * sizeof(double));
  memcpy(diag, alanc, nlanc * sizeof(double));
  memcpy(sub, blanc, nlanc * sizeof(double));
  if (edo_tridiag_eigh(nlanc, diag, sub, Z)) return -1;
  edo_c64 pesoBZ;
  if (finite_t && beta * (ei - egs) < 200) pesoBZ = vnorm2 * exp(-beta * (ei - egs)) / zeta;
  else if (!finite_t) pesoBZ = vnorm2 / zeta;
  else pesoBZ = 0;
  for (int j = 0; j < nlanc; j++) {
    double de = diag[j] - ei;
    edo_c64 peso = pesoBZ * Z[(size_t)j * nlanc] * Z[(size_t)j * nlanc];
    if (poles) poles[j] = isign * de;
    if (weights) weights[j] = peso;
    for (int i = 0; i < lmats; i++) gm[i] += peso / (I * wm[i] - isign * de);
    for (int i = 0; i < lreal; i++) gr[i] += peso / ((wr[i] + I * eps) - isign * de);
  }
  free(diag); free(sub); free(Z);
  return 0;
}

/* ---- lanc_observables: local observables of one eigenstate --------------------------------------
 * ED_OBSERVABLES.f90:94-236 (the master-only loop :120-192): for every basis state i of the sector
 * gs_weight = peso*|vec(i)|^2, impurity occupations from Bdecomp of the up / dw Fock states
 * (imp_state_index(ilat,iorb) = iorb + (ilat-1)*Norb), then
 *   dens_up, dens_dw, docc = <n_up n_dw>, magz = <n_up - n_dw>            [Nlat,Norb]
 *   s2tot(ilat) = <(sum_orb sz(ilat,orb))^2>                                 [Nlat]
 *   sz2, n2 (ilat,jlat,iorb,jorb): the diagonal (ilat,ilat,iorb,iorb) and, for every jlat and jorb > iorb,
 *   the (iorb,jorb) and (jorb,iorb) pairs exactly as the reference loops :172-186 fill them (other
 *   entries stay zero).  Arrays are Fortran-ordered; all outputs are ACCUMULATED (+=) like the
 *   reference's sum over the eigenstates of state_list. */
int32_t edo_lanc_observables(int32_t ns, int32_t nlat, int32_t norb, int32_t isector, const edo_c64 *vec, double peso,
                             double *dens_up, double *dens_dw, double *docc, double *magz, double *s2tot, double *sz2,
                             double *n2) {
  int32_t nup_s, ndw_s;
  edo_get_nup_ndw(ns, isector, &nup_s, &ndw_s);
  int64_t dimup, dimdw;
  const int64_t dim = edo_get_dim(ns, isector, &dimup, &dimdw);
  int32_t *mapu = (int32_t *)malloc(sizeof(int32_t) * (size_t)dimup), *mapd = (int32_t *)malloc(sizeof(int32_t) * (size_t)dimdw);
  if (!mapu || !mapd) { free(mapu); free(mapd); FAIL("edo_lanc_observables: out of memory"); }
  edo_build_sector_map(ns, nup_s, mapu);
  edo_build_sector_map(ns, ndw_s, mapd);
  const int nimp = nlat * norb;
  double nu[64], nd[64], sz[64], nt[64];
  if (nimp > 64) { free(mapu); free(mapd); FAIL("edo_lanc_observables: Nimp > 64"); }
#define IX2(il, io) ((il) + nlat * (io))                                         /* (ilat,iorb), 0-based */
#define IX4(il, jl, io, jo) ((il) + nlat * ((jl) + nlat * ((io) + norb * (jo)))) /* (ilat,jlat,iorb,jorb) */
  for (int64_t i = 0; i < dim; i++) {
    const int64_t iup = i % dimup, idw = i / dimup; /* state2indices, ED_SETUP.f90:547-560 */
    const uint32_t mup = (uint32_t)mapu[iup], mdw = (uint32_t)mapd[idw];
    const double w = peso * (creal(vec[i]) * creal(vec[i]) + cimag(vec[i]) * cimag(vec[i]));
    for (int il = 0; il < nlat; il++)
      for (int io = 0; io < norb; io++) {
        const int pos = io + il * norb; /* imp_state_index - 1 */
        nu[IX2(il, io)] = (double)((mup >> pos) & 1u);
        nd[IX2(il, io)] = (double)((mdw >> pos) & 1u);
        sz[IX2(il, io)] = (nu[IX2(il, io)] - nd[IX2(il, io)]) / 2.0;
        nt[IX2(il, io)] = nu[IX2(il, io)] + nd[IX2(il, io)];
      }
    for (int il = 0; il < nlat; il++) {
      double ssum = 0.0;
      for (int io = 0; io < norb; io++) {
        dens_up[IX2(il, io)] += nu[IX2(il, io)] * w;
        dens_dw[IX2(il, io)] += nd[IX2(il, io)] * w;
        docc[IX2(il, io)] += nu[IX2(il, io)] * nd[IX2(il, io)] * w;
        magz[IX2(il, io)] += (nu[IX2(il, io)] - nd[IX2(il, io)]) * w;
        ssum += sz[IX2(il, io)];
      }
      s2tot[il] += ssum * ssum * w;
    }
    for (int il = 0; il < nlat; il++)
      for (int io = 0; io < norb; io++) {
        sz2[IX4(il, il, io, io)] += sz[IX2(il, io)] * sz[IX2(il, io)] * w;
        n2[IX4(il, il, io, io)] += nt[IX2(il, io)] * nt[IX2(il, io)] * w;
        for (int jl = 0; jl < nlat; jl++)
          for (int jo = io + 1; jo < norb; jo++) {
            sz2[IX4(il, jl, io, jo)] += sz[IX2(il, io)] * sz[IX2(jl, jo)] * w;
            sz2[IX4(il, jl, jo, io)] += sz[IX2(il, jo)] * sz[IX2(jl, io)] * w;
            n2[IX4(il, jl, io, jo)] += nt[IX2(il, io)] * nt[IX2(jl, jo)] * w;
            n2[IX4(il, jl, jo, io)] += nt[IX2(il, jo)] * nt[IX2(jl, io)] * w;
          }
      }
  }
#undef IX2
#undef IX4
  free(mapu);
  free(mapd);
  return 0;
}

/* ---- lanc_local_energy: energy pieces of one eigenstate (ED_OBSERVABLES.f90:246-460, master loop :290-424) ----
 * out[0] ed_Eknot   = sum_i gs_weight * impHloc(ilat,ilat) n  +  sum_i impHloc(is,js) sg1 sg2 vec(i) conjg(vec(j)) peso
 *                     over the impurity hops |j> = c^+_is c_js |i> of both spins (the complex sum lands in a real(8): real part)
 * out[1] ed_Epot    = Uloc n_up n_dw + Ust (n_up,i n_dw,j + n_up,j n_dw,i) + (Ust-Jh)(n_up,i n_up,j + n_dw,i n_dw,j)   [before "+ Ehartree"]
 * out[2] ed_Ehartree  (hfmode only) -- including the reference's constant term 0.25d0*uloc(is), indexed by the impurity
 *                     position `is` instead of the orbital (:392): faithful, uloc has 5 entries (index beyond 5 -> 0 here)
 * out[3] ed_Dust, out[4] ed_Dund.   All ACCUMULATED (+=).  vec = full sector vector. */
int32_t edo_lanc_local_energy(const edo_ctx *c, int32_t isector, const edo_c64 *vec, double peso, double *out) {
  const int ns = c->ns, Nlat = c->m.nlat, Norb = c->m.norb, Nspin = c->m.nspin;
  int32_t nup_s, ndw_s;
  edo_get_nup_ndw(ns, isector, &nup_s, &ndw_s);
  int64_t dimup, dimdw;
  const int64_t dim = edo_get_dim(ns, isector, &dimup, &dimdw);
  int32_t *mapu = (int32_t *)malloc(sizeof(int32_t) * (size_t)dimup), *mapd = (int32_t *)malloc(sizeof(int32_t) * (size_t)dimdw);
  if (!mapu || !mapd) { free(mapu); free(mapd); FAIL("edo_lanc_local_energy: out of memory"); }
  edo_build_sector_map(ns, nup_s, mapu);
  edo_build_sector_map(ns, ndw_s, mapd);
  const double Ust = c->m.ust, Jh = c->m.jh;
  double eknot = 0, epot = 0, ehart = 0, dust = 0, dund = 0;
  for (int64_t i = 0; i < dim; i++) {
    const int64_t iup = i % dimup, idw = i / dimup;
    const int32_t mup = mapu[iup], mdw = mapd[idw];
    const double gs_weight = peso * (creal(vec[i]) * creal(vec[i]) + cimag(vec[i]) * cimag(vec[i]));
#define NUP(is) ((double)((mup >> ((is) - 1)) & 1))
#define NDW(is) ((double)((mdw >> ((is) - 1)) & 1))
    for (int ilat = 1; ilat <= Nlat; ilat++)
      for (int iorb = 1; iorb <= Norb; iorb++) {
        const int is = edo_imp_state_index(c, ilat, iorb);
        eknot += creal(imphloc_at(c, ilat, ilat, 1, 1, iorb, iorb)) * NUP(is) * gs_weight;
        eknot += creal(imphloc_at(c, ilat, ilat, Nspin, Nspin, iorb, iorb)) * NDW(is) * gs_weight;
      }
    for (int ilat = 1; ilat <= Nlat; ilat++)
      for (int jlat = 1; jlat <= Nlat; jlat++)
        for (int iorb = 1; iorb <= Norb; iorb++)
          for (int jorb = 1; jorb <= Norb; jorb++) {
            const int is = edo_imp_state_index(c, ilat, iorb), js = edo_imp_state_index(c, jlat, jorb);
            int32_t k1, k2;
            double sg1, sg2;
            edo_c64 h = imphloc_at(c, ilat, jlat, 1, 1, iorb, jorb);
            if (h != 0 && NUP(js) == 1 && NUP(is) == 0) {
              edo_c(js, mup, &k1, &sg1);
              edo_cdg(is, k1, &k2, &sg2);
              const int64_t j = (edo_binary_search(mapu, (int32_t)dimup, k2) - 1) + idw * dimup;
              eknot += creal(h * sg1 * sg2 * vec[i] * conj(vec[j]) * peso);
            }
            h = imphloc_at(c, ilat, jlat, Nspin, Nspin, iorb, jorb);
            if (h != 0 && NDW(js) == 1 && NDW(is) == 0) {
              edo_c(js, mdw, &k1, &sg1);
              edo_cdg(is, k1, &k2, &sg2);
              const int64_t j = iup + (int64_t)(edo_binary_search(mapd, (int32_t)dimdw, k2) - 1) * dimup;
              eknot += creal(h * sg1 * sg2 * vec[i] * conj(vec[j]) * peso);
            }
          }
    for (int ilat = 1; ilat <= Nlat; ilat++)
      for (int iorb = 1; iorb <= Norb; iorb++) {
        const int is = edo_imp_state_index(c, ilat, iorb);
        epot += c->m.uloc[iorb - 1] * NUP(is) * NDW(is) * gs_weight;
      }
    if (Norb > 1)
      for (int ilat = 1; ilat <= Nlat; ilat++)
        for (int iorb = 1; iorb <= Norb; iorb++)
          for (int jorb = iorb + 1; jorb <= Norb; jorb++) {
            const int is = edo_imp_state_index(c, ilat, iorb), js = edo_imp_state_index(c, ilat, jorb);
            epot += Ust * (NUP(is) * NDW(js) + NUP(js) * NDW(is)) * gs_weight;
            dust += (NUP(is) * NDW(js) + NUP(js) * NDW(is)) * gs_weight;
            epot += (Ust - Jh) * (NUP(is) * NUP(js) + NDW(is) * NDW(js)) * gs_weight;
            dund += (NUP(is) * NUP(js) + NDW(is) * NDW(js)) * gs_weight;
          }
    if (c->m.hfmode) {
      for (int ilat = 1; ilat <= Nlat; ilat++)
        for (int iorb = 1; iorb <= Norb; iorb++) {
          const int is = edo_imp_state_index(c, ilat, iorb);
          const double uis = is <= 5 ? c->m.uloc[is - 1] : 0.0; /* the reference indexes uloc with `is` here (:392) */
          ehart += -0.5 * c->m.uloc[iorb - 1] * (NUP(is) + NDW(is)) * gs_weight + 0.25 * uis * gs_weight;
        }
      if (Norb > 1)
        for (int ilat = 1; ilat <= Nlat; ilat++)
          for (int iorb = 1; iorb <= Norb; iorb++)
            for (int jorb = iorb + 1; jorb <= Norb; jorb++) {
              const int is = edo_imp_state_index(c, ilat, iorb), js = edo_imp_state_index(c, ilat, jorb);
              const double nn = NUP(is) + NDW(is) + NUP(js) + NDW(js);
              ehart += -0.5 * Ust * nn * gs_weight + 0.25 * Ust * gs_weight;
              ehart += -0.5 * (Ust - Jh) * nn * gs_weight + 0.25 * (Ust - Jh) * gs_weight;
            }
    }
#undef NUP
#undef NDW
  }
  free(mapu); free(mapd);
  out[0] += eknot; out[1] += epot; out[2] += ehart; out[3] += dust; out[4] += dund;
  return 0;
}

/* ---- density_matrix_impurity (ED_OBSERVABLES.f90:465-686) for one eigenstate ------------------------------------
 * cdm  [4^Nimp, 4^Nimp] column-major, io = IimpUp + 2^Nimp*IimpDw (0-based here):   rho_IMP = Tr_BATH |vec><vec|,
 *      accumulated exactly as the reference does (:513-575): for every (IimpUp,JimpUp) the bath states shared by both
 *      (sp_return_intersection on the itrace map), the same for dw, then the sum over (IbathUp,IbathDw) of
 *      vec(i) conjg(vec(j)) peso with i, j found by binary_search on Iimp + 2^Nimp*Ibath;
 * spdm [Nlat,Nlat,Nspin,Nspin,Norb,Norb] column-major: <C^+_a C_b> (:600-676): diagonal peso*n*|vec|^2, off-diagonal
 *      peso*sgn1*vec(i)*sgn2*conjg(vec(j)) over the impurity hops |j> = c^+_is c_js |i> of spin ispin.
 * Both ACCUMULATED (+=); vec = full sector vector. */
int32_t edo_density_matrix_impurity(const edo_ctx *c, int32_t isector, const edo_c64 *vec, double peso, edo_c64 *cdm, edo_c64 *spdm) {
  const int ns = c->ns, nimp = c->nimp, Nlat = c->m.nlat, Norb = c->m.norb, Nspin = c->m.nspin;
  const int64_t NI = (int64_t)1 << nimp;
  int32_t nup_s, ndw_s;
  edo_get_nup_ndw(ns, isector, &nup_s, &ndw_s);
  int64_t dimup, dimdw;
  const int64_t dim = edo_get_dim(ns, isector, &dimup, &dimdw);
  int32_t *mapu = (int32_t *)malloc(sizeof(int32_t) * (size_t)dimup), *mapd = (int32_t *)malloc(sizeof(int32_t) * (size_t)dimdw);
  edo_build_sector_map(ns, nup_s, mapu);
  edo_build_sector_map(ns, ndw_s, mapd);
  /* build_sector(isector, HI, itrace=.true.): the sparse maps of both spins */
  int64_t *rpu = (int64_t *)malloc((NI + 1) * sizeof(int64_t)), *rpd = (int64_t *)malloc((NI + 1) * sizeof(int64_t));
  int32_t *bu = (int32_t *)malloc(dimup * sizeof(int32_t)), *iu = (int32_t *)malloc(dimup * sizeof(int32_t));
  int32_t *bd = (int32_t *)malloc(dimdw * sizeof(int32_t)), *id = (int32_t *)malloc(dimdw * sizeof(int32_t));
  edo_build_sparse_map(c, nup_s, rpu, bu, iu);
  edo_build_sparse_map(c, ndw_s, rpd, bd, id);
  int32_t *BATHup = (int32_t *)malloc((dimup + 1) * sizeof(int32_t)), *BATHdw = (int32_t *)malloc((dimdw + 1) * sizeof(int32_t));
  if (cdm)
    for (int64_t IimpUp = 0; IimpUp < NI; IimpUp++)
      for (int64_t JimpUp = 0; JimpUp < NI; JimpUp++) {
        const int32_t lenUp = edo_sparse_map_intersection(rpu, bu, (int32_t)IimpUp, (int32_t)JimpUp, BATHup);
        if (lenUp == 0) continue;
        for (int64_t IimpDw = 0; IimpDw < NI; IimpDw++)
          for (int64_t JimpDw = 0; JimpDw < NI; JimpDw++) {
            const int32_t lenDw = edo_sparse_map_intersection(rpd, bd, (int32_t)IimpDw, (int32_t)JimpDw, BATHdw);
            if (lenDw == 0) continue;
            const int64_t io = IimpUp + NI * IimpDw, jo = JimpUp + NI * JimpDw;
            for (int32_t bUP = 0; bUP < lenUp; bUP++)
              for (int32_t bDW = 0; bDW < lenDw; bDW++) {
                const int64_t IbathUp = BATHup[bUP], IbathDw = BATHdw[bDW];
                const int64_t iUP = edo_binary_search(mapu, (int32_t)dimup, (int32_t)(IimpUp + NI * IbathUp));
                const int64_t iDW = edo_binary_search(mapd, (int32_t)dimdw, (int32_t)(IimpDw + NI * IbathDw));
                const int64_t jUP = edo_binary_search(mapu, (int32_t)dimup, (int32_t)(JimpUp + NI * IbathUp));
                const int64_t jDW = edo_binary_search(mapd, (int32_t)dimdw, (int32_t)(JimpDw + NI * IbathDw));
                const int64_t i = iUP + (iDW - 1) * dimup, j = jUP + (jDW - 1) * dimup; /* 1-based */
                cdm[io + jo * NI * NI] += vec[i - 1] * conj(vec[j - 1]) * peso;
              }
          }
      }
  if (spdm) {
#define SPIX(il, jl, is_, js_, io_, jo_) ((il) + Nlat * ((jl) + Nlat * ((is_) + Nspin * ((js_) + Nspin * ((io_) + Norb * (jo_))))))
    for (int64_t i = 0; i < dim; i++) {
      const int64_t iup = i % dimup, idw = i / dimup;
      const int32_t m[2] = {mapu[iup], mapd[idw]};
      const double w = creal(vec[i]) * creal(vec[i]) + cimag(vec[i]) * cimag(vec[i]);
      for (int ilat = 1; ilat <= Nlat; ilat++)
        for (int ispin = 1; ispin <= Nspin; ispin++)
          for (int iorb = 1; iorb <= Norb; iorb++) {
            const int is = edo_imp_state_index(c, ilat, iorb);
            spdm[SPIX(ilat - 1, ilat - 1, ispin - 1, ispin - 1, iorb - 1, iorb - 1)] += peso * (double)((m[ispin - 1] >> (is - 1)) & 1) * w;
          }
      for (int ispin = 1; ispin <= Nspin; ispin++)
        for (int ilat = 1; ilat <= Nlat; ilat++)
          for (int jlat = 1; jlat <= Nlat; jlat++)
            for (int iorb = 1; iorb <= Norb; iorb++)
              for (int jorb = 1; jorb <= Norb; jorb++) {
                const int is = edo_imp_state_index(c, ilat, iorb), js = edo_imp_state_index(c, jlat, jorb);
                const int32_t st = m[ispin - 1];
                if (((st >> (js - 1)) & 1) == 1 && ((st >> (is - 1)) & 1) == 0) {
                  int32_t r, k;
                  double sgn1, sgn2;
                  edo_c(js, st, &r, &sgn1);
                  edo_cdg(is, r, &k, &sgn2);
                  int64_t j;
                  if (ispin == 1) j = (edo_binary_search(mapu, (int32_t)dimup, k) - 1) + idw * dimup;
                  else j = iup + (int64_t)(edo_binary_search(mapd, (int32_t)dimdw, k) - 1) * dimup;
                  spdm[SPIX(ilat - 1, jlat - 1, ispin - 1, ispin - 1, iorb - 1, jorb - 1)] += peso * sgn1 * vec[i] * sgn2 * conj(vec[j]);
                }
              }
    }
#undef SPIX
  }
  free(mapu); free(mapd); free(rpu); free(rpd); free(bu); free(iu); free(bd); free(id); free(BATHup); free(BATHdw);
  return 0;
}

int32_t edo_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void edo_set_num_threads(int32_t n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}
