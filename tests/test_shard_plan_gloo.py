"""N > 1 host logic on CPU: world_size-2 (and 3) gloo process groups run the distributed transpose of
vector_transpose_MPI (ED_HAMILTONIAN_COMMON.f90:30-101) with the product's shard plan and compare with
the oracle's simulated-MPI transpose; plus the vecDim / communicator-shrink arithmetic."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cdmft_lanc_ed_b200 import shard_plan as sp


def _worker(rank, world, port, nrow, ncol, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(7)
        A = rng.normal(size=(nrow, ncol)) + 1j * rng.normal(size=(nrow, ncol))
        qcol, coff = sp.split_of(ncol, world, rank)
        a_local = np.asfortranarray(A[:, coff:coff + qcol]).ravel(order="F")
        sc, so, rc, ro = sp.transpose_plan(nrow, ncol, world, rank)
        send = sp.pack_for_transpose(a_local, nrow, ncol, world, rank)
        assert [b.size for b in send] == sc
        # complex128 travels as float64 pairs (gloo has no complex all_to_all)
        send_t = [torch.from_numpy(np.ascontiguousarray(b).view(np.float64).copy()) for b in send]
        recv_t = [torch.empty(2 * n, dtype=torch.float64) for n in rc]
        # pairwise isend/irecv = all-to-all (gloo has no all_to_all on CPU tensors in every build)
        reqs = []
        for p in range(world):
            if p == rank:
                recv_t[p].copy_(send_t[p])
            else:
                reqs.append(dist.isend(send_t[p], p))
                reqs.append(dist.irecv(recv_t[p], p))
        for r_ in reqs:
            r_.wait()
        blocks = [t.numpy().view(np.complex128) for t in recv_t]
        b_local = sp.unpack_from_transpose(blocks, nrow, ncol, world, rank)
        qrow, roff = sp.split_of(nrow, world, rank)
        expect = np.asfortranarray(A.T[:, roff:roff + qrow]).ravel(order="F")
        q.put((rank, bool(np.array_equal(b_local, expect)), b_local))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nrow,ncol", [(2, 7, 5), (2, 70, 70), (3, 10, 4)])
def test_distributed_transpose_gloo(oracle_lib, world, nrow, ncol):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + world * 10 + nrow % 7
    procs = [ctx.Process(target=_worker, args=(r, world, port, nrow, ncol, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res)
    # same answer as the oracle's simulated vector_transpose_MPI on the gathered layout
    rng = np.random.default_rng(7)
    A = rng.normal(size=(nrow, ncol)) + 1j * rng.normal(size=(nrow, ncol))
    b = oracle_lib.vector_transpose_sim(world, nrow, ncol, np.asfortranarray(A).ravel(order="F"))
    assert np.array_equal(np.concatenate([x for _, _, x in res]), b)


def test_vecdim_and_shrink_match_oracle(oracle_lib):
    for dimup, dimdw, P in [(70, 70, 8), (12870, 12870, 8), (48620, 48620, 8), (8, 1, 4), (28, 3, 8), (1, 1, 2)]:
        p_eff = sp.active_ranks(dimdw, P)
        tot = 0
        for r in range(P):
            n = sp.vecdim(dimup, dimdw, P, r)
            if r < p_eff:
                assert n == oracle_lib.vecdim(dimup, dimdw, p_eff, r)
                s = oracle_lib.shard_of(dimup, dimdw, p_eff, r)
                assert sp.split_of(dimdw, p_eff, r) == (s["qdw"], s["dw_off"])
                assert sp.split_of(dimup, p_eff, r) == (s["qup"], s["up_off"])
            else:
                assert n == 0
            tot += n
        assert tot == dimup * dimdw


def _exchange(send, rc, rank, world):
    send_t = [torch.from_numpy(np.ascontiguousarray(b).view(np.float64).copy()) for b in send]
    recv_t = [torch.empty(2 * n, dtype=torch.float64) for n in rc]
    reqs = []
    for p in range(world):
        if p == rank:
            recv_t[p].copy_(send_t[p])
        else:
            reqs.append(dist.isend(send_t[p], p))
            reqs.append(dist.irecv(recv_t[p], p))
    for r_ in reqs:
        r_.wait()
    return [t.numpy().view(np.complex128) for t in recv_t]


def _worker_pairs(rank, world, port, nrow, ncol, q):
    """Real Krylov vectors on the sharded layout (csrc/hxv.cu, hxv_local_terms(..., pairs)): two adjacent up-rows of
    the REAL shard are one complex element, the distributed transposes and the Hdw pass (a real matrix acting on the
    dw index) run on that view with nrow/2 rows, and the result read back as reals equals the real computation."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(11)
        A = rng.normal(size=(nrow, ncol))            # real vector v(iup, idw)
        M = rng.normal(size=(ncol, ncol))            # real Hdw acting on the dw index
        qcol, coff = sp.split_of(ncol, world, rank)
        a_real = np.asfortranarray(A[:, coff:coff + qcol]).ravel(order="F")   # this rank's real shard
        a_pair = a_real.view(np.complex128)                                    # (nrow/2) x qcol "complex" shard
        du = nrow // 2
        # forward transpose of the pair view: vt(idw, pair) for my pairs
        sc, so, rc, ro = sp.transpose_plan(du, ncol, world, rank)
        blocks = _exchange(sp.pack_for_transpose(a_pair, du, ncol, world, rank), rc, rank, world)
        vt = sp.unpack_from_transpose(blocks, du, ncol, world, rank)           # column-major vt(ncol, qpair)
        qpair, poff = sp.split_of(du, world, rank)
        VT = vt.reshape(qpair, ncol).T                                         # [idw, pair]
        HVT = M @ VT                                                           # Hdw pass: real matrix on complex pairs
        hvt = np.asfortranarray(HVT).ravel(order="F")
        # backward transpose: hvt(ncol, qpair) -> hv(pair, my columns)
        sc2, so2, rc2, ro2 = sp.transpose_plan(ncol, du, world, rank)
        blocks2 = _exchange(sp.pack_for_transpose(hvt, ncol, du, world, rank), rc2, rank, world)
        hv_pair = sp.unpack_from_transpose(blocks2, ncol, du, world, rank)     # column-major (du, qcol)
        hv_real = np.ascontiguousarray(hv_pair).view(np.float64)               # back to reals: (nrow, qcol)
        expect = np.asfortranarray((A @ M.T)[:, coff:coff + qcol]).ravel(order="F")
        q.put((rank, bool(np.allclose(hv_real, expect, rtol=0, atol=1e-12)), float(np.abs(hv_real - expect).max())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nrow,ncol", [(2, 8, 5), (2, 70, 56), (3, 20, 7)])
def test_paired_row_view_of_real_vectors_gloo(world, nrow, ncol):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + world * 10 + nrow % 7
    procs = [ctx.Process(target=_worker_pairs, args=(r, world, port, nrow, ncol, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res


def _back_worker(rank, world, port, DU, dimdw, nch, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(11)
        # the full (Hdw vt) result as a matrix F[iup, idw]; every rank owns the rows of its up-range
        F = rng.normal(size=(DU, dimdw)) + 1j * rng.normal(size=(DU, dimdw))
        qup, uoff = sp.split_of(DU, world, rank)
        qdw, doff = sp.split_of(dimdw, world, rank)
        hvt_local = np.ascontiguousarray(F[uoff:uoff + qup, :]).ravel()  # [idw + iup_local*dimdw]
        blocks, own = sp.back_pack(hvt_local, DU, dimdw, world, rank, nch)
        # my receive window: q_dw(me) * DU elements; peers write their (sender, chunk) blocks at the layout offsets
        window = np.zeros(qdw * DU, dtype=np.complex128)
        reqs, bufs = [], {}
        for p in range(world):
            if p == rank:
                continue
            mine = [(ch, off, b) for (dst, ch), (off, b) in sorted(blocks.items()) if dst == p]
            uq, _ = sp.split_of(DU, world, p)
            nchp = max(1, min(nch, uq))
            for ch in range(nchp):  # what p sends me, chunk by chunk: sizes follow from the plan alone
                cq, coff = sp.split_of(uq, nchp, ch)
                if cq == 0 or qdw == 0:
                    continue
                t = torch.empty(2 * cq * qdw, dtype=torch.float64)
                bufs[(p, ch)] = (qdw * (sp.split_of(DU, world, p)[1] + coff), t)
                reqs.append(dist.irecv(t, p, tag=ch))
            for ch, off, b in mine:
                reqs.append(dist.isend(torch.from_numpy(b.view(np.float64).copy()), p, tag=ch))
        for r_ in reqs:
            r_.wait()
        for (p, ch), (off, t) in bufs.items():
            blk = t.numpy().view(np.complex128)
            window[off:off + blk.size] = blk
        hv = np.zeros(qdw * DU, dtype=np.complex128)  # hv(i, r) = [i + r*DU]
        HV = hv.reshape(qdw, DU)
        HV[:, uoff:uoff + qup] += own.T  # the own block is added in place by k_xpose_multi
        hv = sp.back_unpack(window, HV.ravel(), DU, dimdw, world, rank, nch)
        expect = np.ascontiguousarray(F[:, doff:doff + qdw].T).ravel()  # my columns of the transposed-back result
        q.put((rank, bool(np.allclose(hv, expect, rtol=0, atol=0))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,DU,dimdw,nch", [(2, 9, 7, 4), (2, 12, 12, 3), (3, 10, 8, 2), (2, 5, 6, 8)])
def test_copy_engine_exchange_window_layout_gloo(world, DU, dimdw, nch):
    """The receive-window layout of the copy-engine exchange's way back (sender blocks cut into chunks of the sender's
    up-rows, csrc/hxv.cu) with the formulas of k_xpose_multi / k_unpack_multi restated in shard_plan.py: the blocks
    travel over gloo between real processes and the unpacked sum equals the transposed-back matrix, including ragged
    splits and more chunks than rows."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + world * 10 + DU % 7 + nch
    procs = [ctx.Process(target=_back_worker, args=(r, world, port, DU, dimdw, nch, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


# ---- collective sp_eigh (P-ARPACK branch, ED_DIAG.f90:153-158): the contract cdmft_b200_eigh has with NCCL, on gloo --------
def _eigh_worker(rank, world, port, case, nup, ndw, neigen, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cdmft_lanc_ed_b200 import ed_hamiltonian as E
        from cdmft_lanc_ed_b200 import models
        from oracle import edo
        mdl = eval("models." + case)
        isec = models.get_sector(mdl.ns, nup, ndw)
        orc = edo.Oracle(mdl)
        orc.build_hv_sector(isec, edo.SPARSE_SERIAL)
        dim = orc.dim
        from math import comb
        dimup, dimdw = comb(mdl.ns, nup), comb(mdl.ns, ndw)
        nloc = sp.vecdim(dimup, dimdw, world, rank)  # this rank's Ndw shard (0 on ranks outside a shrunk communicator)
        off = sum(sp.vecdim(dimup, dimdw, world, r) for r in range(rank))
        sizes = [sp.vecdim(dimup, dimdw, world, r) for r in range(world)]

        def matvec(vloc):  # collective: all-gather the shards, the oracle's H x v on the full vector, keep my rows
            parts = [torch.zeros(2 * max(s, 1), dtype=torch.float64) for s in sizes]
            mine = torch.zeros(2 * max(nloc, 1), dtype=torch.float64)
            mine[: 2 * nloc] = torch.from_numpy(np.ascontiguousarray(vloc).view(np.float64).copy())
            dist.all_gather(parts, mine) if len(set(t.numel() for t in parts)) == 1 else _all_gather_ragged(parts, mine, rank, world)
            full = np.concatenate([p.numpy()[: 2 * s] for p, s in zip(parts, sizes)]).view(np.complex128)
            return orc.hxv(np.ascontiguousarray(full))[off:off + nloc]

        def allreduce(buf):
            t = torch.from_numpy(buf)  # shares memory: the sum lands in the caller's buffer
            dist.all_reduce(t)

        w, z, info = E.eigh_logic_host(matvec, nloc, neigen, nblock=12, tol=1e-13, goff=off, ntot=dim, allreduce=allreduce)
        # the one-process run of the same driver: same start vector (function of the global index), same control flow
        w1, z1, info1 = E.eigh_logic_host(orc.hxv, dim, neigen, nblock=12, tol=1e-13)
        ov = torch.zeros(2 * neigen, dtype=torch.float64)
        for k in range(neigen):
            d = np.vdot(z1[off:off + nloc, k], z[:, k]) if nloc else 0.0
            ov[2 * k], ov[2 * k + 1] = float(np.real(d)), float(np.imag(d))
        dist.all_reduce(ov)
        overlap = [abs(complex(ov[2 * k], ov[2 * k + 1])) for k in range(neigen)]
        q.put((rank, w, w1, info, info1, overlap, nloc))
    finally:
        dist.destroy_process_group()


def _all_gather_ragged(parts, mine, rank, world):
    for r in range(world):
        if r == rank:
            parts[r].copy_(mine)
        dist.broadcast(parts[r], r)


@pytest.mark.parametrize("world,case,nup,ndw,neigen", [(2, "hm2x2(1)", 4, 4, 2), (3, "bhz2(1)", 3, 4, 2), (4, "hm2x2(1)", 2, 1, 2), (4, "hm2x2(1)", 3, 8, 2),
                                                       (2, "random_model(2, 2, 1, seed=15, kanamori=True)", 7, 1, 1)])
def test_collective_sp_eigh_gloo(oracle_lib, world, case, nup, ndw, neigen):
    """Every rank owns its Ndw shard of the Krylov basis, dot products are all-reduced, the start vector depends on the
    global index only: eigenvalues, step counts and eigenvectors equal the one-process run (the property the SPMD GPU path
    relies on; includes ragged shards, a shrunk communicator and DimDw < P)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + world * 10 + nup
    procs = [ctx.Process(target=_eigh_worker, args=(r, world, port, case, nup, ndw, neigen, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
    for rank, w, w1, info, info1, overlap, nloc in res:
        assert info["nconv"] == neigen and info1["nconv"] == neigen
        assert np.abs(w - w1).max() < 1e-11 * max(1.0, np.abs(w1).max()), (rank, w, w1)
        assert abs(info["nmatvec"] - info1["nmatvec"]) <= 12  # rounding may move the last restart, nothing else
        assert all(abs(o - 1.0) < 1e-7 for o in overlap), overlap
    assert len({tuple(np.round(r[1], 12)) for r in res}) == 1  # every rank returns the same eigenvalues
