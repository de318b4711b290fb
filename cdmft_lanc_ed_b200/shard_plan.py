"""Host-side arithmetic of the Ndw sharding and of the distributed transpose, mirrored from the C++
(csrc/ctx.h split_of, csrc/hxv.cu hxv_device) so it can be tested on CPU with world_size > 1.

Reference: ED_HAMILTONIAN.f90:92-105 (mpiQdw, mpiIstart, ...), ED_HAMILTONIAN.f90:197-221
(vecDim_Hv_sector), ED_HAMILTONIAN_COMMON.f90:30-94 (vector_transpose_MPI counts/offsets: the first
mod(n,P) ranks get one more row/column).  The reference exchanges the counts with MPI_AllToAll
(:58-63); they are computable locally, so no exchange is needed here.
"""
from __future__ import annotations

import numpy as np


def split_of(n: int, P: int, r: int):
    """(count, offset) of rank r when n items are split over P ranks, remainder to the first ranks."""
    q, rem = divmod(n, P)
    if r < rem:
        return q + 1, r * (q + 1)
    return q, r * q + rem


def active_ranks(dimdw: int, P: int) -> int:
    """min(P, DimDw): ranks beyond DimDw are dropped from the communicator (ED_HAMILTONIAN.f90:62-90)."""
    return min(P, dimdw)


def vecdim(dimup: int, dimdw: int, P: int, rank: int) -> int:
    p = active_ranks(dimdw, P)
    return dimup * split_of(dimdw, p, rank)[0] if rank < p else 0


def transpose_plan(nrow: int, ncol: int, P: int, rank: int):
    """Send/recv element counts and offsets of rank `rank` for a(nrow, qcol_rank) -> b(ncol, qrow_rank).
    Block sent to s = my columns x s's rows, packed transposed ([row_local_of_s][my col], col fastest);
    block received from r = my rows x r's columns."""
    qcol, _ = split_of(ncol, P, rank)
    qrow, _ = split_of(nrow, P, rank)
    send_counts = [split_of(nrow, P, s)[0] * qcol for s in range(P)]
    recv_counts = [qrow * split_of(ncol, P, r)[0] for r in range(P)]
    send_offs = np.concatenate([[0], np.cumsum(send_counts)[:-1]]).tolist()
    recv_offs = np.concatenate([[0], np.cumsum(recv_counts)[:-1]]).tolist()
    return send_counts, send_offs, recv_counts, recv_offs


def pack_for_transpose(a_local: np.ndarray, nrow: int, ncol: int, P: int, rank: int) -> list[np.ndarray]:
    """a_local: column-major a(nrow, qcol) flattened.  Returns one contiguous block per destination."""
    qcol, _ = split_of(ncol, P, rank)
    A = a_local.reshape(qcol, nrow).T  # A[i, j]
    out = []
    for s in range(P):
        qr, ro = split_of(nrow, P, s)
        out.append(np.ascontiguousarray(A[ro:ro + qr, :]).ravel())  # [row_local][col], col fastest
    return out


def unpack_from_transpose(blocks: list[np.ndarray], nrow: int, ncol: int, P: int, rank: int) -> np.ndarray:
    """Blocks received from every rank -> b(ncol, qrow) column-major flattened."""
    qrow, _ = split_of(nrow, P, rank)
    B = np.empty((qrow, ncol), dtype=blocks[0].dtype)  # B[row_local, col_global]; flatten -> col fastest
    for r in range(P):
        qc, co = split_of(ncol, P, r)
        B[:, co:co + qc] = blocks[r].reshape(qrow, qc)
    return B.ravel()


# --------------------------------------------------------------------------------------------------
# Way back of the copy-engine exchange (csrc/hxv.cu: hxv_sharded_ce, k_xpose_multi, k_unpack_multi): the Hdw pass runs
# in `nch` chunks of the sender's up-rows; every (sender, chunk) block is packed transposed [iup_chunk + idw * q_chunk]
# and lands in the owner's receive window at q_dw(owner) * (up_off(sender) + chunk_off).  The same formulas, in numpy.
# --------------------------------------------------------------------------------------------------
def xchg_chunks(dimup_rows: int, dimdw: int, P: int, option: int = 0) -> int:
    """Chunk count every rank computes (it enters the window layout): about 96 MB per DMA copy, between 2 and 6."""
    if option > 0:
        return option
    block_bytes = (dimdw // P + 1) * (dimup_rows // P + 1) * 16
    return max(2, min(6, -(-block_bytes // (96 << 20))))


def back_pack(hvt_local: np.ndarray, DU: int, dimdw: int, P: int, rank: int, nch_opt: int):
    """hvt_local: my (Hdw vt) block, layout [idw + iup_local * dimdw] (dimdw x qup).  Returns {(dest, chunk): (window
    offset in elements, packed block)} for every destination != rank, plus the own block's contribution as a dense
    (qup, q_dw(rank)) array [iup_local, idw_local]."""
    qup, uoff = split_of(DU, P, rank)
    H = hvt_local.reshape(qup, dimdw)  # H[iup_local, idw]
    nch = max(1, min(nch_opt, qup))
    out = {}
    for p in range(P):
        qd, doff = split_of(dimdw, P, p)
        if p == rank:
            continue
        for ch in range(nch):
            cq, coff = split_of(qup, nch, ch)
            if cq == 0 or qd == 0:
                continue
            blk = H[coff:coff + cq, doff:doff + qd].T  # [idw_p, iup_chunk] -> flatten: iup_chunk fastest
            out[(p, ch)] = (qd * (uoff + coff), np.ascontiguousarray(blk).ravel())
    qd, doff = split_of(dimdw, P, rank)
    own = H[:, doff:doff + qd]
    return out, own


def back_unpack(window: np.ndarray, hv_local: np.ndarray, DU: int, dimdw: int, P: int, rank: int, nch_opt: int):
    """k_unpack_multi: hv(i, r) += window element of the sender that owns up-row i (all rows outside my own range)."""
    qdw, _ = split_of(dimdw, P, rank)
    HV = hv_local.reshape(qdw, DU)  # HV[r, i]
    for i in range(DU):
        p = next(s for s in range(P) if split_of(DU, P, s)[1] <= i < split_of(DU, P, s)[1] + split_of(DU, P, s)[0])
        if p == rank:
            continue
        uq, uoff = split_of(DU, P, p)
        nch = max(1, min(nch_opt, uq))
        q, rem = divmod(uq, nch)
        il = i - uoff
        if il < rem * (q + 1):
            ch = il // (q + 1); coff = ch * (q + 1); cq = q + 1
        else:
            ch = rem + (il - rem * (q + 1)) // max(q, 1); coff = ch * q + rem; cq = q
        for r in range(qdw):
            HV[r, i] += window[qdw * (uoff + coff) + (il - coff) + r * cq]
    return HV.ravel()
