"""Host-side mirror of the reference's ED_HAMILTONIAN interface on top of the C ABI.

Same names, argument meaning and error behaviour as the Fortran module (the reference's
toolchain -- gfortran + MPI + SciFortran -- is absent from this image, so the host layer
above ``include/cdmft_b200.h`` is Python/ctypes here; the Fortran ISO_C_BINDING shim a
maintainer would compile instead is ``fortran/ED_HAMILTONIAN_B200.f90``):

    build_Hv_sector(isector)        ED_HAMILTONIAN.f90:39-143
    delete_Hv_sector()              ED_HAMILTONIAN.f90:149-190
    vecDim_Hv_sector(isector)       ED_HAMILTONIAN.f90:197-221
    spHtimesV_p(Nloc, v, Hv)        ED_VARS_GLOBAL.f90:72-78,146   (procedure pointer; None when unset)
    sp_lanc_eigh / sp_lanc_tridiag  SciFortran drivers called at ED_DIAG.f90:176, ED_GF_NORMAL.f90:215

The product path is CUDA only: importing works without a GPU (so CPU-side tests can check
symbols), every compute call raises ``EdB200Error`` when no device / library is available.
Nothing here imports ``oracle/``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
# CDMFT_B200_LIB: developer override (kernel experiments build several libraries side by side)
LIB_PATH = os.environ.get("CDMFT_B200_LIB") or os.path.join(_PKG, "libcdmft_b200.so")

SPARSE, DIRECT = 1, 0  # ed_sparse_H = T / F (ED_INPUT_VARS.f90:145)


class EdB200Error(RuntimeError):
    """Raised where the reference executes `stop "message"`."""


class _CModel(C.Structure):
    _fields_ = [("nlat", C.c_int32), ("norb", C.c_int32), ("nspin", C.c_int32), ("nbath", C.c_int32),
                ("uloc", C.c_double * 5), ("ust", C.c_double), ("jh", C.c_double), ("jx", C.c_double),
                ("jp", C.c_double), ("xmu", C.c_double), ("hfmode", C.c_int32),
                ("quirk_direct_bathdiag", C.c_int32),
                ("imphloc", C.c_void_p), ("hbath", C.c_void_p), ("vbath", C.c_void_p)]


_lib = None

# every symbol include/cdmft_b200.h declares (tests check the .so exports all of them)
ABI_SYMBOLS = [
    "cdmft_b200_last_error", "cdmft_b200_init", "cdmft_b200_nccl_unique_id", "cdmft_b200_init_rank",
    "cdmft_b200_init_sim", "cdmft_b200_finalize", "cdmft_b200_set_stream", "cdmft_b200_reset_stream",
    "cdmft_b200_profile_query", "cdmft_b200_launch_count", "cdmft_b200_ipc_export", "cdmft_b200_ipc_import",
    "cdmft_b200_set_option", "cdmft_b200_set_model", "cdmft_b200_get_ns", "cdmft_b200_get_sector_dims",
    "cdmft_b200_vecdim_hv_sector", "cdmft_b200_build_hv_sector", "cdmft_b200_delete_hv_sector",
    "cdmft_b200_active_ranks", "cdmft_b200_hxv", "cdmft_b200_hxv64", "cdmft_b200_get_sector_map",
    "cdmft_b200_get_csr_nnz", "cdmft_b200_get_csr", "cdmft_b200_get_diag", "cdmft_b200_get_sparse_map",
    "cdmft_b200_lanczos_tridiag", "cdmft_b200_lanczos_gs", "cdmft_b200_apply_op",
    "cdmft_b200_add_to_lanczos_gf", "cdmft_b200_schedule_host", "cdmft_b200_imp_weights",
    "cdmft_b200_colblk_host", "cdmft_b200_add_to_lanczos_gf_full", "cdmft_b200_build_hmat", "cdmft_b200_scatter_vector",
    "cdmft_b200_gather_vector", "cdmft_b200_imp_kinetic", "cdmft_b200_density_matrices",
    "cdmft_b200_eigh", "cdmft_b200_eigh_logic_host",
]


def load_library():
    """dlopen libcdmft_b200.so (built in-tree by __graft_entry__.build()). No fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EdB200Error(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.cdmft_b200_last_error.restype = C.c_char_p
        _lib = L
    return _lib


def _chk(rc):
    if rc != 0:
        raise EdB200Error(load_library().cdmft_b200_last_error().decode())


def _ptr(a):
    """numpy array, torch tensor (CPU or CUDA) or raw int address -> void*"""
    if a is None:
        return C.c_void_p(0)
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    raise TypeError(f"unsupported buffer type {type(a)}")


# --------------------------------------------------------------------------------------
# runtime (ed_set_MpiComm, ED_VARS_GLOBAL.f90:365-380)
# --------------------------------------------------------------------------------------
_state = {"inited": False, "nranks": 1, "rank": 0, "mode": None}
spHtimesV_p = None  # the procedure pointer: bound by build_Hv_sector, None after delete_Hv_sector


def ed_init(device: int = 0):
    """Serial solver (MpiStatus = .false.)."""
    _chk(load_library().cdmft_b200_init(C.c_int32(device)))
    _state.update(inited=True, nranks=1, rank=0, mode="single")


def ed_init_sim(nranks: int, device: int = 0):
    """P simulated MPI ranks on one GPU (exercises the Ndw-sharded code path)."""
    _chk(load_library().cdmft_b200_init_sim(C.c_int32(device), C.c_int32(nranks)))
    _state.update(inited=True, nranks=nranks, rank=0, mode="sim")


def ed_set_MpiComm(device: int | None = None):
    """SPMD: one process per GPU, ranks from torch.distributed (the communicator the reference
    receives from its driver, drivers/cdn_hm_2dsquare.f90:39).  The NCCL unique id is created on
    rank 0 and broadcast through the already-initialised process group (any backend)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        raise EdB200Error("ed_set_MpiComm: torch.distributed is not initialised")
    rank, world = dist.get_rank(), dist.get_world_size()
    L = load_library()
    uid = (C.c_ubyte * 128)()
    if rank == 0:
        _chk(L.cdmft_b200_nccl_unique_id(uid))
    t = torch.tensor(list(bytes(uid)), dtype=torch.uint8)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.broadcast(t, src=0)
    uid = (C.c_ubyte * 128)(*t.cpu().tolist())
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", rank))
    _chk(L.cdmft_b200_init_rank(C.c_int32(device), C.c_int32(world), C.c_int32(rank), uid))
    _state.update(inited=True, nranks=world, rank=rank, mode="spmd")


def ed_finalize():
    global spHtimesV_p
    if _lib is not None:
        _chk(_lib.cdmft_b200_finalize())
    _state.update(inited=False, nranks=1, rank=0, mode=None)
    spHtimesV_p = None


def set_stream(cuda_stream: int):
    """Launch on this cudaStream_t handle (0 = CUDA's legacy default stream, torch's default)."""
    _chk(load_library().cdmft_b200_set_stream(C.c_void_p(cuda_stream)))


def reset_stream():
    _chk(load_library().cdmft_b200_reset_stream())


def set_option(key: str, value: int):
    _chk(load_library().cdmft_b200_set_option(key.encode(), C.c_int64(value)))


def profile_query(kind: int):
    """(total ms, launches) of kernel kind 0 column pass / 1 row pass / 2 transpose / 3 NCCL since last query."""
    ms, n = C.c_double(), C.c_int64()
    _chk(load_library().cdmft_b200_profile_query(C.c_int32(kind), C.byref(ms), C.byref(n)))
    return ms.value, n.value


def launch_count() -> int:
    n = C.c_int64()
    _chk(load_library().cdmft_b200_launch_count(C.byref(n)))
    return n.value


# --------------------------------------------------------------------------------------
# model = what ED_HAMILTONIAN reads from globals (set_Himpurity + set_dmft_bath, ED_MAIN.f90:248-254)
# --------------------------------------------------------------------------------------
_model_keep = None


def ed_set_model(model, quirk_direct_bathdiag: bool = False):
    global _model_keep
    a = np.asfortranarray(model.imphloc, dtype=np.complex128)
    b = np.asfortranarray(model.hbath, dtype=np.complex128)
    v = np.asfortranarray(model.vbath, dtype=np.float64)
    _model_keep = (a, b, v)
    m = _CModel(model.nlat, model.norb, model.nspin, model.nbath, (C.c_double * 5)(*model.uloc), model.ust,
                model.jh, model.jx, model.jp, model.xmu, int(model.hfmode), int(quirk_direct_bathdiag),
                a.ctypes.data, b.ctypes.data, v.ctypes.data)
    _chk(load_library().cdmft_b200_set_model(C.byref(m)))


def get_Ns() -> int:
    n = C.c_int32()
    _chk(load_library().cdmft_b200_get_ns(C.byref(n)))
    return n.value


def get_Sector(nup: int, ndw: int, ns: int | None = None) -> int:
    """ED_SETUP.f90:446-457"""
    ns = get_Ns() if ns is None else ns
    return 1 + nup * (ns + 1) + ndw


def getDim(isector: int):
    """(Dim, DimUp, DimDw) -- ED_SETUP.f90:316-322, int64 (the reference overflows at Ns=18)."""
    du, dd, d = C.c_int64(), C.c_int64(), C.c_int64()
    _chk(load_library().cdmft_b200_get_sector_dims(C.c_int32(isector), C.byref(du), C.byref(dd), C.byref(d)))
    return d.value, du.value, dd.value


# --------------------------------------------------------------------------------------
# the three module procedures + the pointer
# --------------------------------------------------------------------------------------
_sector = {"isector": 0, "nloc": 0}


def _hxv(Nloc, v, Hv):
    """cc_sparse_HxV contract: Hv = H_sector * v on the local shard; host or device buffers."""
    _chk(load_library().cdmft_b200_hxv64(C.c_int64(Nloc), _ptr(v), _ptr(Hv)))


def build_Hv_sector(isector: int, ed_sparse_H: bool = True) -> int:
    """Returns the local vector length (what vecDim_Hv_sector reports)."""
    global spHtimesV_p
    nloc = C.c_int64()
    _chk(load_library().cdmft_b200_build_hv_sector(C.c_int32(isector), C.c_int32(SPARSE if ed_sparse_H else DIRECT),
                                                   C.byref(nloc)))
    _sector.update(isector=isector, nloc=nloc.value)
    spHtimesV_p = _hxv
    return nloc.value


def ipc_exchange():
    """SPMD, after build_Hv_sector: map every peer's transpose windows (CUDA IPC) so the distributed
    transposes store straight into peer memory over NVLink.  Collective over torch.distributed."""
    import torch
    import torch.distributed as dist
    L = load_library()
    mine = (C.c_ubyte * 128)()
    _chk(L.cdmft_b200_ipc_export(mine))
    world = dist.get_world_size()
    t = torch.tensor(list(bytes(mine)), dtype=torch.uint8)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    allh = bytes(torch.cat(out).cpu().tolist())
    buf = (C.c_ubyte * len(allh))(*allh)
    _chk(L.cdmft_b200_ipc_import(buf, C.c_int32(world)))


def delete_Hv_sector():
    global spHtimesV_p
    _chk(load_library().cdmft_b200_delete_hv_sector())
    _sector.update(isector=0, nloc=0)
    spHtimesV_p = None


def vecDim_Hv_sector(isector: int) -> int:
    n = C.c_int64()
    _chk(load_library().cdmft_b200_vecdim_hv_sector(C.c_int32(isector), C.byref(n)))
    return n.value


def active_ranks() -> int:
    p = C.c_int32()
    _chk(load_library().cdmft_b200_active_ranks(C.byref(p)))
    return p.value


def hxv(v: np.ndarray) -> np.ndarray:
    """Convenience: allocate Hv and call the pointer (host arrays)."""
    if spHtimesV_p is None:
        raise EdB200Error("spHtimesV_p is not associated (call build_Hv_sector)")
    v = np.ascontiguousarray(v, dtype=np.complex128)
    out = np.empty_like(v)
    spHtimesV_p(v.size, v, out)
    return out


# --------------------------------------------------------------------------------------
# inspection
# --------------------------------------------------------------------------------------
def schedule_host(rowptr, col, code, g: int = 8, natural: bool = False, nwarps: int = 32):
    """Gather schedule of the column-resident kernel for a CSR pattern (host only, no GPU needed).
    Returns a list over warps of lists of tasks; a task = (rows[32] (-1 = no row), words[nsteps, 32]) with
    words[k, lane] = (source row << 7) | code, in the order the warp executes them."""
    L = load_library()
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
    col = np.ascontiguousarray(col, dtype=np.int32)
    code = np.ascontiguousarray(code, dtype=np.uint8)
    n = len(rowptr) - 1
    ntask, nquads = C.c_int32(0), C.c_int64(0)
    args = (C.c_int64(n), _ptr(rowptr), _ptr(col), _ptr(code), C.c_int32(g), C.c_int32(1 if natural else 0), C.c_int32(nwarps))
    _chk(L.cdmft_b200_schedule_host(*args, C.byref(ntask), C.byref(nquads), None, None, None, None))
    tbase = np.zeros(nwarps + 1, np.int32)
    qbase = np.zeros(nwarps + 1, np.int32)
    meta = np.zeros(ntask.value * 32 * 4, np.uint32)
    words = np.zeros(nquads.value * 32 * 4, np.uint32)
    _chk(L.cdmft_b200_schedule_host(*args, C.byref(ntask), C.byref(nquads), _ptr(tbase), _ptr(qbase), _ptr(meta), _ptr(words)))
    meta = meta.reshape(-1, 32, 4)
    words = words.reshape(-1, 32, 4)  # [quad, lane, step in quad]
    out = []
    for w in range(nwarps):
        q = int(qbase[w])
        tasks = []
        for t in range(tbase[w], tbase[w + 1]):
            nq = int(meta[t, 0, 3] >> 16)
            assert (meta[t, :, 3] >> 16 == nq).all()
            rows = meta[t, :, 2].astype(np.int64)
            rows[rows == 0xFFFFFFFF] = -1
            tasks.append((rows, words[q:q + nq].transpose(0, 2, 1).reshape(nq * 4, 32)))
            q += nq
        assert q == qbase[w + 1]
        out.append(tasks)
    return out


def colblk_host(ns: int, npart: int, rowptr, col, code, g: int = 8, natural: bool = False, cap_rows: int = 64):
    """Block-split schedules of the column-resident kernel for big columns (host only).  Returns a list of blocks:
    dict(g0, ng, tasks=[(rows[32] relative to the block (-1 none), words[nsteps, 32], off[noff, 32])])."""
    L = load_library()
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
    col = np.ascontiguousarray(col, dtype=np.int32)
    code = np.ascontiguousarray(code, dtype=np.uint8)
    sizes = np.zeros(8, np.int64)
    args = (C.c_int32(ns), C.c_int32(npart), _ptr(rowptr), _ptr(col), _ptr(code), C.c_int32(g), C.c_int32(1 if natural else 0),
            C.c_int64(cap_rows), _ptr(sizes))
    _chk(L.cdmft_b200_colblk_host(*args, None, None, None, None, None, None, None))
    nblk, ntq, nmeta, nwords, ntask, nwoff, nwarps, _ = [int(x) for x in sizes]
    blk = np.zeros(nblk * 4, np.int32)
    tbase, qbase = np.zeros(ntq, np.int32), np.zeros(ntq, np.int32)
    meta, words = np.zeros(nmeta, np.uint32), np.zeros(nwords, np.uint32)
    toff, woff = np.zeros(ntask * 2, np.uint32), np.zeros(nwoff, np.uint32)
    _chk(L.cdmft_b200_colblk_host(*args, _ptr(blk), _ptr(tbase), _ptr(qbase), _ptr(meta), _ptr(words), _ptr(toff), _ptr(woff)))
    blk, meta, words = blk.reshape(-1, 4), meta.reshape(-1, 32, 4), words.reshape(-1, 32, 4)
    toff, woff = toff.reshape(-1, 2), woff.reshape(-1, 32)
    out = []
    for b in range(nblk):
        g0, ng, task0, unit0 = [int(x) for x in blk[b]]
        tasks = []
        for w in range(nwarps):
            q = unit0 + int(qbase[b * (nwarps + 1) + w])
            for t in range(task0 + tbase[b * (nwarps + 1) + w], task0 + tbase[b * (nwarps + 1) + w + 1]):
                nq = int(meta[t, 0, 3] >> 16)
                rows = meta[t, :, 2].astype(np.int64)
                rows[rows == 0xFFFFFFFF] = -1
                ob, no = int(toff[t, 0]), int(toff[t, 1])
                tasks.append((rows, words[q:q + nq].transpose(0, 2, 1).reshape(nq * 4, 32), woff[ob:ob + no]))
                q += nq
        out.append(dict(g0=g0, ng=ng, tasks=tasks))
    return out


def get_sector_map(which: int) -> np.ndarray:
    _, du, dd = getDim(_sector["isector"])
    m = np.empty(du if which == 1 else dd, dtype=np.int32)
    _chk(load_library().cdmft_b200_get_sector_map(C.c_int32(which), _ptr(m)))
    return m


def get_csr(which: int):
    nnz = C.c_int64()
    _chk(load_library().cdmft_b200_get_csr_nnz(C.c_int32(which), C.byref(nnz)))
    _, du, dd = getDim(_sector["isector"])
    n = du if which == 1 else dd
    rowptr = np.zeros(n + 1, dtype=np.int64)
    col = np.zeros(nnz.value, dtype=np.int32)
    val = np.zeros(nnz.value, dtype=np.complex128)
    _chk(load_library().cdmft_b200_get_csr(C.c_int32(which), _ptr(rowptr), _ptr(col), _ptr(val)))
    return rowptr, col, val


def get_diag() -> np.ndarray:
    d = np.zeros(_sector["nloc"], dtype=np.float64)
    _chk(load_library().cdmft_b200_get_diag(C.c_int64(d.size), _ptr(d)))
    return d


def get_sparse_map(which: int, nimp: int):
    _, du, dd = getDim(_sector["isector"])
    n = du if which == 1 else dd
    rowptr = np.zeros((1 << nimp) + 1, dtype=np.int64)
    bath = np.zeros(n, dtype=np.int32)
    indx = np.zeros(n, dtype=np.int32)
    _chk(load_library().cdmft_b200_get_sparse_map(C.c_int32(which), _ptr(rowptr), _ptr(bath), _ptr(indx)))
    return rowptr, bath, indx


# --------------------------------------------------------------------------------------
# Krylov drivers (SciFortran call sites)
# --------------------------------------------------------------------------------------
def sp_lanc_tridiag(vin, nitermax: int, threshold: float = 1e-12):
    """sp_lanc_tridiag(spHtimesV_p, vin, alanc, blanc): returns (ndone, alanc, blanc)."""
    n = vin.numel() if hasattr(vin, "numel") else vin.size
    a = np.zeros(nitermax)
    b = np.zeros(nitermax)
    nd = C.c_int32()
    _chk(load_library().cdmft_b200_lanczos_tridiag(C.c_int64(n), _ptr(vin), C.c_int32(nitermax), C.c_double(threshold),
                                                   _ptr(a), _ptr(b), C.byref(nd)))
    return nd.value, a, b


def sp_lanc_eigh(vect, nitermax: int = 512, threshold: float = 1e-18, ncheck: int = 10):
    """sp_lanc_eigh(spHtimesV_p, egs, vect, Nitermax, threshold): vect (numpy complex128 or CUDA
    tensor) is the start vector on entry (all zero -> constant) and the eigenvector on exit.
    Returns (egs, niter, alanc, blanc)."""
    n = vect.numel() if hasattr(vect, "numel") else vect.size
    e = C.c_double()
    nit = C.c_int32()
    a = np.zeros(nitermax)
    b = np.zeros(nitermax)
    _chk(load_library().cdmft_b200_lanczos_gs(C.c_int64(n), _ptr(vect), C.c_int32(nitermax), C.c_double(threshold),
                                              C.c_int32(ncheck), C.byref(e), C.byref(nit), _ptr(a), _ptr(b)))
    return e.value, nit.value, a[: nit.value], b[: nit.value]


# --------------------------------------------------------------------------------------
# Green's function helpers (ED_GF_NORMAL.f90)
# --------------------------------------------------------------------------------------
def sp_eigh(neigen: int, nblock: int | None = None, nitermax: int = 512, tol: float = 1e-18, matvec=None, n: int | None = None,
            v0=None):
    """SciFortran `sp_eigh(MatVec, eig_values, eig_basis, Nblock, Nitermax, tol)` -- the reference's default
    LANC_METHOD (ED_DIAG.f90:94-97,150-170): (P)ARPACK in reverse communication, which='SA', ncv = Nblock, the
    mat-vec being the procedure pointer.  Here ARPACK stays on the host exactly as in the reference (scipy's eigsh IS
    ARPACK's znaupd/zneupd) and every mat-vec goes through spHtimesV_p on the device: the drop-in a Fortran caller
    gets by only re-binding the pointer.  Returns (eig_values ascending [neigen], eig_basis [n, neigen]).
    `matvec`/`n` default to the active sector's hxv / Nloc (single rank; tests pass the oracle's mat-vec)."""
    from scipy.sparse.linalg import LinearOperator, eigsh
    if matvec is None:
        if spHtimesV_p is None:
            raise EdB200Error("sp_eigh: Hsector NOT set (call build_Hv_sector)")
        if _state["mode"] == "spmd":
            raise EdB200Error("sp_eigh: host ARPACK drives a single rank (P-ARPACK is not mirrored)")
        matvec, n = hxv, _sector["nloc"]
    if neigen >= n:
        raise EdB200Error("sp_eigh: Neigen must be < Dim (the reference diagonalises such sectors densely, ED_DIAG.f90:104-106)")
    # Nblock = min(dim, lanc_ncv_factor*max(Neigen, lanc_nstates_sector) + lanc_ncv_add), defaults 10, 2, 0 (ED_INPUT_VARS.f90:171-175)
    ncv = min(n, nblock if nblock else 10 * max(neigen, 2))
    ncv = max(ncv, min(n, neigen + 2))
    op = LinearOperator((n, n), matvec=lambda x: matvec(np.ascontiguousarray(x, dtype=np.complex128)), dtype=np.complex128)
    w, z = eigsh(op, k=neigen, which="SA", ncv=ncv, maxiter=nitermax * max(1, n // max(ncv, 1)) if nitermax else None,
                 tol=0.0 if tol < 1e-15 else tol, v0=v0)
    order = np.argsort(w)
    return w[order], z[:, order]


def sp_eigh_device(neigen: int, nblock: int | None = None, nitermax: int = 512, tol: float = 1e-18, basis=None):
    """The same call -- `sp_eigh(MpiComm, spHtimesV_p, eig_values, eig_basis, Nblock, Nitermax, tol)`, ED_DIAG.f90:150-170 --
    without ARPACK on the host: `cdmft_b200_eigh`, a device-resident thick-restart Lanczos (csrc/trlan.h).  Works on one
    rank, on simulated ranks and collectively in SPMD mode (the P-ARPACK branch).  `basis`: optional CUDA tensor
    [neigen, nloc] complex128 that receives the eigenvectors (row i = eig_basis(:, i)); otherwise a numpy array
    [nloc, neigen] is returned.  Returns (eig_values, eig_basis, info) with info = dict(nconv, nmatvec)."""
    if spHtimesV_p is None:
        raise EdB200Error("sp_eigh: Hsector NOT set (call build_Hv_sector)")
    n = _sector["nloc"]
    # Nblock = min(dim, lanc_ncv_factor*max(Neigen, lanc_nstates_sector) + lanc_ncv_add), defaults 10, 2, 0 (ED_INPUT_VARS.f90:171-175)
    ncv = nblock if nblock else 10 * max(neigen, 2)
    w = np.zeros(neigen)
    out = basis if basis is not None else np.zeros((neigen, max(n, 1)), dtype=np.complex128)
    nconv, nmv = C.c_int32(), C.c_int32()
    _chk(load_library().cdmft_b200_eigh(C.c_int64(n), C.c_int32(neigen), C.c_int32(ncv), C.c_int32(nitermax), C.c_double(tol),
                                        _ptr(w), _ptr(out), C.byref(nconv), C.byref(nmv)))
    info = dict(nconv=nconv.value, nmatvec=nmv.value)
    if basis is not None:
        return w, basis, info
    return w, np.ascontiguousarray(out[:, :n].T), info


_MATVEC_FN = C.CFUNCTYPE(None, C.c_int64, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p)
_ALLREDUCE_FN = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.c_int64, C.c_void_p)


def eigh_logic_host(matvec, n: int, neigen: int, nblock: int | None = None, nitermax: int = 512, tol: float = 1e-18,
                    goff: int = 0, ntot: int | None = None, allreduce=None):
    """CPU test hook: the restart logic of `cdmft_b200_eigh` on host vectors around `matvec` (complex128[n] -> complex128[n],
    e.g. the oracle's H x v).  Needs no GPU and computes no Hamiltonian itself.  Sharded runs: `n` local elements starting
    at global index `goff` of `ntot`, `allreduce(array)` sums a float64 array in place over the ranks.
    Returns (eig_values, eig_basis [n, neigen], info)."""
    def cb(nn, pv, phv, _user):
        if nn == 0:
            matvec(np.zeros(0, dtype=np.complex128))
            return
        v = np.ctypeslib.as_array(pv, shape=(2 * nn,)).view(np.complex128)
        hv = np.ctypeslib.as_array(phv, shape=(2 * nn,)).view(np.complex128)
        hv[:] = matvec(v.copy())

    def red(pbuf, count, _user):
        buf = np.ctypeslib.as_array(pbuf, shape=(count,))
        allreduce(buf)
    fn = _MATVEC_FN(cb)
    rfn = _ALLREDUCE_FN(red) if allreduce is not None else C.cast(None, _ALLREDUCE_FN)
    w = np.zeros(neigen)
    z = np.zeros((neigen, max(n, 1)), dtype=np.complex128)
    nconv, nmv = C.c_int32(), C.c_int32()
    ncv = nblock if nblock else 10 * max(neigen, 2)
    _chk(load_library().cdmft_b200_eigh_logic_host(C.c_int64(n), C.c_int64(goff), C.c_int64(n if ntot is None else ntot), fn, rfn, None,
                                                   C.c_int32(neigen), C.c_int32(ncv), C.c_int32(nitermax), C.c_double(tol), _ptr(w),
                                                   _ptr(z), C.byref(nconv), C.byref(nmv)))
    return w, np.ascontiguousarray(z[:, :n].T), dict(nconv=nconv.value, nmatvec=nmv.value)


def apply_op(isector: int, iop: int, ispin: int, pos, coef, state: np.ndarray):
    """(sum_k coef[k] op_{pos[k]}) |state>; returns (jsector, vector) or (0, None)."""
    ns = get_Ns()
    nup, ndw = (isector - 1) // (ns + 1), (isector - 1) % (ns + 1)
    jn = [nup, ndw]
    jn[ispin - 1] += 1 if iop > 0 else -1
    if min(jn) < 0 or max(jn) > ns:
        return 0, None
    jsec = get_Sector(jn[0], jn[1], ns)
    # SPMD: spin-up operators act shard by shard (same Ndw split in both sectors); the result is this rank's shard
    jdim = vecDim_Hv_sector(jsec) if _state["mode"] == "spmd" else getDim(jsec)[0]
    pos = np.ascontiguousarray(pos, dtype=np.int32)
    coef = np.ascontiguousarray(coef, dtype=np.complex128)
    state = np.ascontiguousarray(state, dtype=np.complex128)
    out = np.zeros(jdim, dtype=np.complex128)
    js = C.c_int32()
    _chk(load_library().cdmft_b200_apply_op(C.c_int32(isector), C.c_int32(iop), C.c_int32(ispin), C.c_int32(pos.size),
                                            _ptr(pos), _ptr(coef), _ptr(state), _ptr(out), C.byref(js)))
    return js.value, out


def add_to_lanczos_gf_normal(vnorm2, Ei, alanc, blanc, isign, zeta, wm, g):
    """Accumulates into g (complex128[Lmats]) in place; returns (poles, weights)."""
    n = len(alanc)
    a = np.ascontiguousarray(alanc, dtype=np.float64)
    b = np.ascontiguousarray(blanc, dtype=np.float64)
    wm = np.ascontiguousarray(wm, dtype=np.float64)
    vn = (C.c_double * 2)(complex(vnorm2).real, complex(vnorm2).imag)
    poles = np.zeros(n)
    weights = np.zeros(n, dtype=np.complex128)
    _chk(load_library().cdmft_b200_add_to_lanczos_gf(vn, C.c_double(Ei), C.c_int32(n), _ptr(a), _ptr(b), C.c_int32(isign),
                                                     C.c_double(zeta), C.c_int32(wm.size), _ptr(wm), _ptr(g), _ptr(poles),
                                                     _ptr(weights)))
    return poles, weights


def add_to_lanczos_gf_normal_full(vnorm2, Ei, Egs, finiteT, beta, alanc, blanc, isign, zeta, wm, gmats, wr, eps, greal):
    """add_to_lanczos_gf_normal with the finite-temperature weight and the real-axis accumulation (impGreal);
    gmats [Lmats] / greal [Lreal] complex128, accumulated in place; returns (poles, weights)."""
    n = len(alanc)
    a = np.ascontiguousarray(alanc, dtype=np.float64)
    b = np.ascontiguousarray(blanc, dtype=np.float64)
    wm = np.ascontiguousarray(wm, dtype=np.float64)
    wr = np.ascontiguousarray(wr, dtype=np.float64)
    vn = (C.c_double * 2)(complex(vnorm2).real, complex(vnorm2).imag)
    poles = np.zeros(n)
    weights = np.zeros(n, dtype=np.complex128)
    _chk(load_library().cdmft_b200_add_to_lanczos_gf_full(
        vn, C.c_double(Ei), C.c_double(Egs), C.c_int32(int(finiteT)), C.c_double(beta), C.c_int32(n), _ptr(a), _ptr(b),
        C.c_int32(isign), C.c_double(zeta), C.c_int32(wm.size), _ptr(wm), _ptr(gmats), C.c_int32(wr.size), _ptr(wr),
        C.c_double(eps), _ptr(greal), _ptr(poles), _ptr(weights)))
    return poles, weights


def build_Hmat() -> np.ndarray:
    """The dense `Hmat` of build_Hv_sector(isector, Hmat) for the active sector: complex128 [Dim, Dim], Fortran order."""
    dim = getDim(_sector["isector"])[0]
    h = np.zeros((dim, dim), dtype=np.complex128, order="F")
    _chk(load_library().cdmft_b200_build_hmat(_ptr(h)))
    return h


def scatter_vector_MPI(vfull, vloc, root: int = 0):
    """scatter_vector_MPI(MpiComm, v, vloc): the root's full vector -> every rank's shard (ED_SETUP.f90:575-611)."""
    _chk(load_library().cdmft_b200_scatter_vector(_ptr(vfull), _ptr(vloc), C.c_int32(root)))


def gather_vector_MPI(vloc, vfull, root: int = 0):
    """gather_vector_MPI(MpiComm, vloc, v): shards -> the root's full vector (ED_SETUP.f90:633-668)."""
    _chk(load_library().cdmft_b200_gather_vector(_ptr(vloc), _ptr(vfull), C.c_int32(root)))


# --------------------------------------------------------------------------------------
# local observables (ED_OBSERVABLES.f90:94-236)
# --------------------------------------------------------------------------------------
def imp_weights(vec, nimp: int) -> np.ndarray:
    """W[mu, md] = sum over bath configurations of |vec|^2 for impurity bits (mu, md); vec lives in the active sector."""
    n = int(vec.shape[0]) if hasattr(vec, "shape") else len(vec)
    w = np.zeros((1 << nimp) * (1 << nimp))
    _chk(load_library().cdmft_b200_imp_weights(C.c_int64(n), _ptr(vec), _ptr(w)))
    return w.reshape(1 << nimp, 1 << nimp).T.copy()  # [mu, md]


def observables_from_weights(W: np.ndarray, nlat: int, norb: int, peso: float = 1.0) -> dict:
    """Local observables of lanc_observables (ED_OBSERVABLES.f90:120-192) from the impurity-configuration weights
    W[mu, md], as matrix contractions instead of a loop over configurations: with the occupation table
    N[config, a] (a = imp_state_index - 1) and the marginals Wu = sum_md W, Wd = sum_mu W,
        <n_a,up> = N^T Wu,  <n_a,dw> = N^T Wd,  <n_a,up n_b,dw> = N^T W N,  <n_a,s n_b,s> = N^T diag(W_s) N,
    and Sz = (n_up - n_dw)/2, n = n_up + n_dw follow by linearity.  Returns Fortran-ordered arrays with the
    reference's fill pattern (sz2 / n2: same-orbital entries only on the same site)."""
    nimp = nlat * norb
    cfg = np.arange(1 << nimp)
    N = ((cfg[:, None] >> np.arange(nimp)[None, :]) & 1).astype(float)  # [config, a]
    W = peso * np.asarray(W, dtype=float)
    Wu, Wd = W.sum(axis=1), W.sum(axis=0)
    nu, nd = N.T @ Wu, N.T @ Wd
    UD = N.T @ W @ N                       # <n_a,up n_b,dw>
    UU, DD = (N.T * Wu) @ N, (N.T * Wd) @ N
    SZ = (UU - UD - UD.T + DD) / 4.0       # <Sz_a Sz_b>
    NN = UU + UD + UD.T + DD               # <n_a n_b>
    shape2 = lambda x: np.asfortranarray(x.reshape(nlat, norb))  # a = iorb + ilat*norb
    out = {"dens_up": shape2(nu), "dens_dw": shape2(nd), "docc": shape2(np.diag(UD).copy()), "magz": shape2(nu - nd)}
    a_of = (np.arange(nlat)[:, None] * norb + np.arange(norb)[None, :])  # [ilat, iorb] -> a
    out["s2tot"] = np.array([SZ[np.ix_(a_of[il], a_of[il])].sum() for il in range(nlat)])
    for name, M in (("sz2", SZ), ("n2", NN)):
        full = M[a_of[:, None, :, None], a_of[None, :, None, :]]  # [ilat, jlat, iorb, jorb]
        same_orb = np.eye(norb, dtype=bool)[None, None, :, :] & ~np.eye(nlat, dtype=bool)[:, :, None, None]
        out[name] = np.asfortranarray(np.where(same_orb, 0.0, full))
    out["dens"] = out["dens_up"] + out["dens_dw"]
    return out


def imp_kinetic(vec) -> complex:
    """<vec| K |vec>, K = impurity block of the hopping part of H (both spins); vec lives in the active sector."""
    n = int(vec.shape[0]) if hasattr(vec, "shape") else len(vec)
    out = (C.c_double * 2)()
    _chk(load_library().cdmft_b200_imp_kinetic(C.c_int64(n), _ptr(vec), out))
    return complex(out[0], out[1])


def local_energy_from_weights(W: np.ndarray, model, peso: float = 1.0) -> dict:
    """The occupation-dependent pieces of lanc_local_energy (ED_OBSERVABLES.f90:246-460) from the impurity-configuration
    weights W[mu, md], as contractions (N[config, a] = occupation table, a = imp_state_index - 1):
    Eknot_diag = sum_a impHloc(a,a) n_a, Epot = U n_up n_dw + Ust(...) + (Ust-Jh)(...), Dust, Dund, Ehartree (hfmode;
    the reference's constant term is 0.25*uloc(is) with the IMPURITY index, ED_OBSERVABLES.f90:392 -- kept)."""
    nlat, norb, nspin = model.nlat, model.norb, model.nspin
    nimp = nlat * norb
    cfg = np.arange(1 << nimp)
    N = ((cfg[:, None] >> np.arange(nimp)[None, :]) & 1).astype(float)
    W = peso * np.asarray(W, dtype=float)
    Wu, Wd = W.sum(axis=1), W.sum(axis=0)
    nu, nd = N.T @ Wu, N.T @ Wd                       # <n_a,up>, <n_a,dw>
    UD = N.T @ W @ N                                  # <n_a,up n_b,dw>
    UU, DD = (N.T * Wu) @ N, (N.T * Wd) @ N           # <n_a,s n_b,s>
    tot = W.sum()
    hu = np.array([model.imphloc[a // norb, a // norb, 0, 0, a % norb, a % norb].real for a in range(nimp)])
    hd = np.array([model.imphloc[a // norb, a // norb, nspin - 1, nspin - 1, a % norb, a % norb].real for a in range(nimp)])
    u = np.asarray(model.uloc, dtype=float)
    out = dict(Eknot_diag=float(hu @ nu + hd @ nd), Epot=0.0, Ehartree=0.0, Dust=0.0, Dund=0.0)
    for il in range(nlat):
        for io in range(norb):
            a = io + il * norb
            out["Epot"] += u[io] * UD[a, a]
            if model.hfmode:
                out["Ehartree"] += -0.5 * u[io] * (nu[a] + nd[a]) + 0.25 * (u[a] if a < 5 else 0.0) * tot
            for jo in range(io + 1, norb):
                b = jo + il * norb
                dust = UD[a, b] + UD[b, a]
                dund = UU[a, b] + DD[a, b]
                out["Dust"] += dust
                out["Dund"] += dund
                out["Epot"] += model.ust * dust + (model.ust - model.jh) * dund
                if model.hfmode:
                    nn = nu[a] + nd[a] + nu[b] + nd[b]
                    out["Ehartree"] += -0.5 * (2 * model.ust - model.jh) * nn + 0.25 * (2 * model.ust - model.jh) * tot
    return out


def lanc_local_energy(vec, model, peso: float = 1.0) -> dict:
    """lanc_local_energy for one eigenstate of the active sector: {Eknot, Epot (before '+ Ehartree'), Ehartree, Dust, Dund}.
    O(Dim) work on the device (weight table + one restricted H x v), formulas on the host."""
    out = local_energy_from_weights(imp_weights(vec, model.nlat * model.norb), model, peso)
    out["Eknot"] = out.pop("Eknot_diag") + peso * imp_kinetic(vec).real
    return out


def density_matrix_impurity(vec, nlat: int, norb: int, nspin: int, peso: float = 1.0):
    """density_matrix_impurity for one eigenstate of the active sector: (cluster_density_matrix [4^Nimp, 4^Nimp],
    single_particle_density_matrix [Nlat,Nlat,Nspin,Nspin,Norb,Norb]), Fortran-ordered complex128."""
    n = int(vec.shape[0]) if hasattr(vec, "shape") else len(vec)
    ni = 1 << (2 * nlat * norb)
    cdm = np.zeros((ni, ni), dtype=np.complex128, order="F")
    sp = np.zeros((nlat, nlat, nspin, nspin, norb, norb), dtype=np.complex128, order="F")
    _chk(load_library().cdmft_b200_density_matrices(C.c_int64(n), _ptr(vec), C.c_double(peso), _ptr(cdm), _ptr(sp)))
    return cdm, sp


def lanc_observables(vec, nlat: int, norb: int, peso: float = 1.0) -> dict:
    """lanc_observables for one eigenstate of the active sector (weights on the device, formulas on the host)."""
    return observables_from_weights(imp_weights(vec, nlat * norb), nlat, norb, peso)
