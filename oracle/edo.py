"""ctypes binding of the CPU ORACLE (oracle/ed_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(cdmft_lanc_ed_b200) never imports this module.  Parity status: "parity
unpinned" by reference fixtures (the reference has none);
anchors are oracle/jw_ed.py and the cross-checks in tests/test_oracle_*.py.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

SPARSE_SERIAL, SPARSE_MPI, DIRECT_SERIAL, DIRECT_MPI = 0, 1, 2, 3


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libedoracle.so")
    src = [os.path.join(_HERE, f) for f in ("ed_oracle.c", "ed_oracle.h")]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "libedoracle.so"], check=True, capture_output=True)
    return so


class _Model(C.Structure):
    _fields_ = [("nlat", C.c_int32), ("norb", C.c_int32), ("nspin", C.c_int32), ("nbath", C.c_int32),
                ("uloc", C.c_double * 5), ("ust", C.c_double), ("jh", C.c_double), ("jx", C.c_double),
                ("jp", C.c_double), ("xmu", C.c_double), ("hfmode", C.c_int32),
                ("imphloc", C.c_void_p), ("hbath", C.c_void_p), ("vbath", C.c_void_p)]


class _Shard(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("qdw", "rdw", "q", "r", "istart", "iend", "ishift", "qup", "up_off", "dw_off")]


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.edo_create.restype = C.c_void_p
        L.edo_last_error.restype = C.c_char_p
        L.edo_binomial.restype = C.c_int64
        L.edo_get_dim.restype = C.c_int64
        L.edo_build_sector_map.restype = C.c_int64
        L.edo_build_sparse_map.restype = C.c_int64
        L.edo_vecdim.restype = C.c_int64
        L.edo_sector_dims.restype = C.c_int64
        L.edo_get_csr.restype = C.c_int64
        L.edo_get_nonlocal.restype = C.c_int64
        L.edo_counter_vec.restype = None
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _chk(rc):
    if rc != 0:
        raise RuntimeError("oracle: " + lib().edo_last_error().decode())


# ---- free functions -------------------------------------------------------
def binomial(n, k):
    return lib().edo_binomial(C.c_int32(n), C.c_int32(k))


def sector_map(ns, n):
    L = lib()
    cnt = L.edo_build_sector_map(C.c_int32(ns), C.c_int32(n), None)
    m = np.empty(cnt, dtype=np.int32)
    L.edo_build_sector_map(C.c_int32(ns), C.c_int32(n), _p(m))
    return m


def counter_vec(i0, n, seed, scale):
    """The oracle's copy of the counter-based synthetic vector (cdmft_lanc_ed_b200/synth.py holds the numpy / torch ones)."""
    out = np.empty(n, dtype=np.complex128)
    lib().edo_counter_vec(C.c_int64(i0), C.c_int64(n), C.c_uint64(seed), C.c_double(scale), _p(out))
    return out


def shard_of(dimup, dimdw, P, rank):
    s = _Shard()
    lib().edo_shard_of(C.c_int64(dimup), C.c_int64(dimdw), C.c_int32(P), C.c_int32(rank), C.byref(s))
    return {n: getattr(s, n) for n, _ in _Shard._fields_}


def vecdim(dimup, dimdw, P, rank):
    return lib().edo_vecdim(C.c_int64(dimup), C.c_int64(dimdw), C.c_int32(P), C.c_int32(rank))


def vector_transpose_sim(P, nrow, ncol, a):
    a = np.ascontiguousarray(a, dtype=np.complex128)
    b = np.empty_like(a)
    _chk(lib().edo_vector_transpose_sim(C.c_int32(P), C.c_int64(nrow), C.c_int64(ncol), _p(a), _p(b)))
    return b


def c_op(pos, state):
    out, sg = C.c_int32(), C.c_double()
    rc = lib().edo_c(C.c_int32(pos), C.c_int32(state), C.byref(out), C.byref(sg))
    return rc, out.value, sg.value


def cdg_op(pos, state):
    out, sg = C.c_int32(), C.c_double()
    rc = lib().edo_cdg(C.c_int32(pos), C.c_int32(state), C.byref(out), C.byref(sg))
    return rc, out.value, sg.value


def binary_search(a, value):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return lib().edo_binary_search(_p(a), C.c_int32(a.size), C.c_int32(value))


def tridiag_eigh(d, e):
    """d[0..n-1], e[0..n-1] with e[0] unused (blanc layout). Returns (evals, Z)."""
    n = len(d)
    d = np.array(d, dtype=np.float64)
    e = np.array(e, dtype=np.float64)
    z = np.zeros((n, n), order="F")
    _chk(lib().edo_tridiag_eigh(C.c_int32(n), _p(d), _p(e), _p(z)))
    return d, z


def add_to_lanczos_gf(vnorm2, ei, alanc, blanc, isign, zeta, wm, g):
    """Accumulates into g (complex128 [Lmats]) in place; returns (poles, weights)."""
    n = len(alanc)
    a = np.ascontiguousarray(alanc, dtype=np.float64)
    b = np.ascontiguousarray(blanc, dtype=np.float64)
    wm = np.ascontiguousarray(wm, dtype=np.float64)
    poles = np.zeros(n)
    weights = np.zeros(n, dtype=np.complex128)
    vn = (C.c_double * 2)(complex(vnorm2).real, complex(vnorm2).imag)
    f = lib().edo_add_to_lanczos_gf
    # double complex by value: pass as two doubles in SSE regs == struct of 2 doubles on SysV x86-64
    class _Z(C.Structure):
        _fields_ = [("re", C.c_double), ("im", C.c_double)]
    f.argtypes = [_Z, C.c_double, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_double, C.c_int32,
                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    _chk(f(_Z(vn[0], vn[1]), ei, n, _p(a), _p(b), isign, zeta, wm.size, _p(wm), _p(g), _p(poles), _p(weights)))
    return poles, weights


def add_to_lanczos_gf_full(vnorm2, ei, egs, finite_t, beta, alanc, blanc, isign, zeta, wm, gm, wr, eps, gr):
    """All branches of add_to_lanczos_gf_normal: T=0 / finite-T weight, Matsubara and real axis.  Accumulates into
    gm [Lmats] and gr [Lreal] in place; returns (poles, weights)."""
    n = len(alanc)
    a = np.ascontiguousarray(alanc, dtype=np.float64)
    b = np.ascontiguousarray(blanc, dtype=np.float64)
    wm = np.ascontiguousarray(wm, dtype=np.float64)
    wr = np.ascontiguousarray(wr, dtype=np.float64)
    poles = np.zeros(n)
    weights = np.zeros(n, dtype=np.complex128)

    class _Z(C.Structure):
        _fields_ = [("re", C.c_double), ("im", C.c_double)]
    f = lib().edo_add_to_lanczos_gf_full
    f.argtypes = [_Z, C.c_double, C.c_double, C.c_int32, C.c_double, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_double,
                  C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
    z = complex(vnorm2)
    _chk(f(_Z(z.real, z.imag), ei, egs, int(finite_t), beta, n, _p(a), _p(b), isign, zeta, wm.size, _p(wm), _p(gm),
           wr.size, _p(wr), eps, _p(gr), _p(poles), _p(weights)))
    return poles, weights


def apply_op(ns, isector, iop, ispin, pos, coef, state):
    """(sum_k coef[k] * op_{pos[k]}) |state>, op = cdg (iop=+1) / c (iop=-1). Returns (jsector, out)."""
    L = lib()
    jsec = (L.edo_get_cdg_sector if iop > 0 else L.edo_get_c_sector)(C.c_int32(ns), C.c_int32(ispin), C.c_int32(isector))
    if jsec == 0:
        return 0, None
    jdim = L.edo_get_dim(C.c_int32(ns), C.c_int32(jsec), None, None)
    pos = np.ascontiguousarray(pos, dtype=np.int32)
    coef = np.ascontiguousarray(coef, dtype=np.complex128)
    state = np.ascontiguousarray(state, dtype=np.complex128)
    out = np.zeros(jdim, dtype=np.complex128)
    js = C.c_int32()
    _chk(L.edo_apply_op(C.c_int32(ns), C.c_int32(isector), C.c_int32(iop), C.c_int32(ispin), C.c_int32(pos.size),
                        _p(pos), _p(coef), _p(state), _p(out), C.byref(js)))
    return js.value, out


def lanc_observables(ns, nlat, norb, isector, vec, peso=1.0):
    """Local observables of one eigenstate, ED_OBSERVABLES.f90:120-192.  Returns a dict of Fortran-ordered arrays."""
    L = lib()
    vec = np.ascontiguousarray(vec, dtype=np.complex128)
    out = {k: np.zeros((nlat, norb), order="F") for k in ("dens_up", "dens_dw", "docc", "magz")}
    out["s2tot"] = np.zeros(nlat)
    out["sz2"] = np.zeros((nlat, nlat, norb, norb), order="F")
    out["n2"] = np.zeros((nlat, nlat, norb, norb), order="F")
    _chk(L.edo_lanc_observables(C.c_int32(ns), C.c_int32(nlat), C.c_int32(norb), C.c_int32(isector), _p(vec), C.c_double(peso),
                                _p(out["dens_up"]), _p(out["dens_dw"]), _p(out["docc"]), _p(out["magz"]), _p(out["s2tot"]),
                                _p(out["sz2"]), _p(out["n2"])))
    out["dens"] = out["dens_up"] + out["dens_dw"]
    return out


# ---- context ---------------------------------------------------------------
class Oracle:
    """Owns one edo_ctx for a cdmft_lanc_ed_b200.models.Model."""

    def __init__(self, model):
        self.model = model
        self._keep = (np.asfortranarray(model.imphloc), np.asfortranarray(model.hbath), np.asfortranarray(model.vbath))
        m = _Model(model.nlat, model.norb, model.nspin, model.nbath, (C.c_double * 5)(*model.uloc),
                   model.ust, model.jh, model.jx, model.jp, model.xmu, int(model.hfmode),
                   self._keep[0].ctypes.data, self._keep[1].ctypes.data, self._keep[2].ctypes.data)
        self.L = lib()
        self.h = C.c_void_p(self.L.edo_create(C.byref(m)))
        if not self.h:
            raise RuntimeError("oracle: " + self.L.edo_last_error().decode())
        self.ns = self.L.edo_ns(self.h)

    def close(self):
        if self.h:
            self.L.edo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def get_dim(self, isector):
        du, dd = C.c_int64(), C.c_int64()
        d = self.L.edo_get_dim(C.c_int32(self.ns), C.c_int32(isector), C.byref(du), C.byref(dd))
        return d, du.value, dd.value

    def build_hv_sector(self, isector, kind=SPARSE_SERIAL, P=1, quirk=0):
        _chk(self.L.edo_build_hv_sector(self.h, C.c_int32(isector), C.c_int32(kind), C.c_int32(P), C.c_int32(quirk)))
        self.dim, self.dimup, self.dimdw = self.get_dim(isector)

    def delete_hv_sector(self):
        _chk(self.L.edo_delete_hv_sector(self.h))

    def active_ranks(self):
        return self.L.edo_active_ranks(self.h)

    def hxv(self, v):
        v = np.ascontiguousarray(v, dtype=np.complex128)
        hv = np.empty_like(v)
        _chk(self.L.edo_hxv(self.h, C.c_int64(v.size), _p(v), _p(hv)))
        return hv

    def hxv_rows_counter(self, rows, seed, scale):
        """Rows `rows` (0-based global indices) of H x v for the counter-based vector v(i) = scale * hash(i, seed)
        (ed_oracle.h: edo_hxv_rows_counter); works on any active sector, never materialises v."""
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        out = np.empty(rows.size, dtype=np.complex128)
        _chk(self.L.edo_hxv_rows_counter(self.h, C.c_int64(rows.size), _p(rows), C.c_uint64(seed), C.c_double(scale), _p(out)))
        return out

    def lanc_local_energy(self, isector, vec, peso=1.0):
        """{Eknot, Epot, Ehartree, Dust, Dund} of one eigenstate (ED_OBSERVABLES.f90:246-460); Epot before '+ Ehartree'."""
        vec = np.ascontiguousarray(vec, dtype=np.complex128)
        out = np.zeros(5)
        _chk(self.L.edo_lanc_local_energy(self.h, C.c_int32(isector), _p(vec), C.c_double(peso), _p(out)))
        return dict(zip(("Eknot", "Epot", "Ehartree", "Dust", "Dund"), out.tolist()))

    def density_matrix_impurity(self, isector, vec, peso=1.0):
        """(cluster_density_matrix [4^Nimp,4^Nimp], single_particle_density_matrix [Nlat,Nlat,Nspin,Nspin,Norb,Norb]) of one
        eigenstate, Fortran-ordered (ED_OBSERVABLES.f90:465-686)."""
        m = self.model
        ni = 1 << (2 * m.nlat * m.norb)
        vec = np.ascontiguousarray(vec, dtype=np.complex128)
        cdm = np.zeros((ni, ni), dtype=np.complex128, order="F")
        sp = np.zeros((m.nlat, m.nlat, m.nspin, m.nspin, m.norb, m.norb), dtype=np.complex128, order="F")
        _chk(self.L.edo_density_matrix_impurity(self.h, C.c_int32(isector), _p(vec), C.c_double(peso), _p(cdm), _p(sp)))
        return cdm, sp

    def get_csr(self, which):
        nnz = self.L.edo_get_csr(self.h, C.c_int32(which), None, None, None)
        if nnz < 0:
            raise RuntimeError("oracle: CSR not built (sparse kinds only)")
        n = self.dimup if which == 1 else self.dimdw
        rowptr = np.zeros(n + 1, dtype=np.int64)
        col = np.zeros(nnz, dtype=np.int32)
        val = np.zeros(nnz, dtype=np.complex128)
        self.L.edo_get_csr(self.h, C.c_int32(which), _p(rowptr), _p(col), _p(val))
        return rowptr, col, val

    def get_diag(self):
        d = np.zeros(self.dim)
        _chk(self.L.edo_get_diag(self.h, _p(d)))
        return d

    def get_nonlocal(self):
        nnz = self.L.edo_get_nonlocal(self.h, None, None, None)
        rowptr = np.zeros(self.dim + 1, dtype=np.int64)
        col = np.zeros(max(nnz, 0), dtype=np.int64)
        val = np.zeros(max(nnz, 0), dtype=np.complex128)
        if nnz > 0:
            self.L.edo_get_nonlocal(self.h, _p(rowptr), _p(col), _p(val))
        return rowptr, col, val

    def dense_hmat(self, isector):
        d, _, _ = self.get_dim(isector)
        h = np.zeros((d, d), dtype=np.complex128, order="F")
        _chk(self.L.edo_dense_hmat(self.h, C.c_int32(isector), _p(h)))
        return h

    def sparse_map(self, n):
        nimpst = 1 << self.model.nimp
        tot = self.L.edo_build_sparse_map(self.h, C.c_int32(n), None, None, None)
        rowptr = np.zeros(nimpst + 1, dtype=np.int64)
        bath = np.zeros(tot, dtype=np.int32)
        indx = np.zeros(tot, dtype=np.int32)
        self.L.edo_build_sparse_map(self.h, C.c_int32(n), _p(rowptr), _p(bath), _p(indx))
        return rowptr, bath, indx

    def lanc_tridiag(self, v0, nitermax, threshold=1e-12):
        v = np.array(v0, dtype=np.complex128)
        a = np.zeros(nitermax)
        b = np.zeros(nitermax)
        n = self.L.edo_lanc_tridiag(self.h, C.c_int64(v.size), _p(v), C.c_int32(nitermax), _p(a), _p(b), C.c_double(threshold))
        if n < 0:
            _chk(-1)
        return n, a, b

    def lanc_eigh(self, nitermax=512, threshold=1e-18, ncheck=10, v0=None):
        v = np.zeros(self.dim, dtype=np.complex128) if v0 is None else np.array(v0, dtype=np.complex128)
        e = C.c_double()
        nit = C.c_int32()
        a = np.zeros(nitermax)
        b = np.zeros(nitermax)
        _chk(self.L.edo_lanc_eigh(self.h, C.c_int64(v.size), C.byref(e), _p(v), C.c_int32(nitermax), C.c_double(threshold),
                                  C.c_int32(ncheck), C.byref(nit), _p(a), _p(b)))
        return e.value, v, nit.value, a[: nit.value], b[: nit.value]
