"""Multi-GPU parity check, run under torchrun (one rank per GPU):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node=N --master-addr 127.0.0.1 tools/spmd_check.py
Each rank owns its Ndw shard; H x v through every transpose back-end, the Lanczos tridiagonalisation (complex and real
start vectors) and a shard-local c^+ are compared on rank 0 with the CPU oracle's simulated-MPI path."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cdmft_lanc_ed_b200 import models, shard_plan as sp  # noqa: E402
from cdmft_lanc_ed_b200 import ed_hamiltonian as E  # noqa: E402


# One launch checks every transpose back-end: (use_ipc, xchg_chunks) = copy-engine exchange over CUDA-IPC windows
# (default; 4, 1 and 3 chunks), transposing kernels that store into peer memory, NCCL send/recv (no windows mapped).
# --full adds the Ns=12 case (slow on the single-threaded oracle under torchrun); --modes=a,b restricts the list.
MODES = {"ce4": (1, 4), "ce1": (1, 1), "ce3": (1, 3), "peer": (2, 4), "nccl": (0, 4)}
FULL = "--full" in sys.argv
SEL = [a.split("=")[1].split(",") for a in sys.argv[1:] if a.startswith("--modes=")]
SEL = SEL[0] if SEL else list(MODES)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // 2))  # torchrun pins it to 1; the oracle runs on rank 0 only
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    E.ed_set_MpiComm(local)
    ok = True
    cases = [(models.hm2x2(1), (4, 4)), (models.hm2x2(1), (1, 0)), (models.bhz2(1), (4, 3)),
             (models.random_model(2, 2, 1, nspin=2, seed=12), (4, 4)),
             (models.random_model(2, 2, 1, seed=15, kanamori=True), (4, 4)), (models.bhz2(1, kanamori=True), (5, 3))]
    if FULL:
        cases.append((models.hm2x2(2), (6, 6)))
    for mdl, (nup, ndw) in cases:
        E.ed_set_model(mdl)
        ns = mdl.ns
        isec = models.get_sector(ns, nup, ndw)
        dim, dimup, dimdw = E.getDim(isec)
        rng = np.random.default_rng(100 + isec)
        v = rng.normal(size=dim) + 1j * rng.normal(size=dim)
        v /= np.linalg.norm(v)
        vreal = (v.real / np.linalg.norm(v.real)).astype(np.complex128)
        for sparse in (True, False):
            ref = None
            if rank == 0:  # oracle once per (case, kind), shared by all back-ends
                from oracle import edo
                orc = edo.Oracle(mdl)
                orc.build_hv_sector(isec, edo.SPARSE_MPI if sparse else edo.DIRECT_MPI, world)
                ref = dict(hv=orc.hxv(v), tri=orc.lanc_tridiag(v, 30), trir=orc.lanc_tridiag(vreal, 30),
                           cop=edo.apply_op(ns, isec, +1, 1, [1], [1.0 + 0.0j], v),
                           copd=edo.apply_op(ns, isec, -1, 2, [min(2, ns)], [0.5 - 0.25j], v), p_eff=orc.active_ranks())
                orc.delete_hv_sector()
                ref["eig"] = np.linalg.eigvalsh(orc.dense_hmat(isec)) if sparse and 8 <= dim <= 5000 else None
            for mode in SEL:
                use_ipc, chunks = MODES[mode]
                E.set_option("use_ipc", use_ipc)
                E.set_option("xchg_chunks", chunks)
                nloc = E.build_Hv_sector(isec, sparse)
                if use_ipc:
                    E.ipc_exchange()
                assert nloc == E.vecDim_Hv_sector(isec) == sp.vecdim(dimup, dimdw, world, rank)
                p_eff = E.active_ranks()
                off = sum(sp.vecdim(dimup, dimdw, world, r) for r in range(rank))
                vloc = np.ascontiguousarray(v[off:off + nloc])
                hv = np.empty_like(vloc)
                E.spHtimesV_p(nloc, vloc, hv)
                hv2 = np.empty_like(vloc)
                E.spHtimesV_p(nloc, vloc, hv2)  # a second product re-uses vt and the windows: the barriers must order it
                assert np.array_equal(hv, hv2), "second product differs"
                nd, a, b = E.sp_lanc_tridiag(vloc, 30)
                # real start vector: real Krylov vectors on the sharded layout (paired-row view) when H is real
                vr = np.ascontiguousarray(vreal[off:off + nloc])
                ndr, ar, br = E.sp_lanc_tridiag(vr, 30)
                # sp_eigh, P-ARPACK branch (ED_DIAG.f90:153-158): the device-resident thick-restart Lanczos is collective;
                # two lowest pairs against the dense spectrum, residuals through the sharded mat-vec (first back-end only)
                eig_ok, eig_err = True, 0.0
                if mode == SEL[0] and sparse and 8 <= dim <= 5000:
                    w, z, info = E.sp_eigh_device(2, nblock=12, tol=1e-13)
                    res2 = torch.zeros(2, dtype=torch.float64, device="cuda")
                    for k in range(2):
                        zk = np.ascontiguousarray(z[:, k]) if nloc else np.zeros(0, dtype=np.complex128)
                        hz = np.empty_like(zk)
                        E.spHtimesV_p(nloc, zk, hz)
                        res2[k] = float(np.linalg.norm(hz - w[k] * zk) ** 2)
                    dist.all_reduce(res2)
                    if rank == 0:
                        dref = ref["eig"]
                        eig_err = float(np.abs(w - dref[:2]).max()) if abs(dref[2] - dref[1]) > 1e-8 and abs(dref[1] - dref[0]) > 1e-8 \
                            else float(min(np.abs(dref - w[0]).min(), 1.0))
                        eig_ok = info["nconv"] == 2 and eig_err < 1e-10 * max(1.0, float(np.abs(dref[:2]).max())) and float(res2.max()) < 1e-16
                # c^+_{1,up} on the sharded vector (start vector of a GF channel, ED_GF_NORMAL.f90:180-194): shard-local
                jsec, cv = E.apply_op(isec, +1, 1, [1], [1.0 + 0.0j], vloc)
                cparts = [None] * world
                dist.all_gather_object(cparts, cv if jsec else np.zeros(0, dtype=np.complex128))
                # c_{2,dw}: a spin-DOWN operator changes the column split (gather -> transform -> scatter inside the library)
                jsecd, cvd = E.apply_op(isec, -1, 2, [min(2, ns)], [0.5 - 0.25j], vloc)
                dparts = [None] * world
                dist.all_gather_object(dparts, cvd if jsecd else np.zeros(0, dtype=np.complex128))
                # gather on rank 0 (gather_vector_MPI, ED_SETUP.f90:633-668)
                parts = [None] * world
                dist.all_gather_object(parts, hv)
                # the library's own scatter_vector_MPI / gather_vector_MPI (root 0; host buffers)
                sc = np.zeros(max(nloc, 1), dtype=np.complex128)
                E.scatter_vector_MPI(v if rank == 0 else None, sc, 0)
                assert np.array_equal(sc[:nloc], vloc), "scatter_vector_MPI"
                gf = np.zeros(dim, dtype=np.complex128) if rank == 0 else None
                E.gather_vector_MPI(hv if nloc else np.zeros(1, dtype=np.complex128), gf, 0)
                if rank == 0:
                    assert np.array_equal(gf, np.concatenate([p for p in parts if p.size])), "gather_vector_MPI"
                E.delete_Hv_sector()
                if rank == 0:
                    ond, oa, ob = ref["tri"]
                    ondr, oar, obr = ref["trir"]
                    got = np.concatenate([p for p in parts if p.size])
                    err = np.abs(got - ref["hv"]).max() / max(np.abs(ref["hv"]).max(), 1e-300)
                    k = min(nd, ond, 20)
                    erra = np.abs(a[:k] - oa[:k]).max() / max(np.abs(oa[:k]).max(), 1e-300)
                    kr = min(ndr, ondr, 20)
                    errr = np.abs(ar[:kr] - oar[:kr]).max() / max(np.abs(oar[:kr]).max(), 1e-300)
                    ojsec, ocv = ref["cop"]
                    gotc = np.concatenate([p for p in cparts if p is not None and p.size] or [np.zeros(0, dtype=np.complex128)])
                    okc = (ojsec == jsec) and (ojsec == 0 or (gotc.size == ocv.size and np.abs(gotc - ocv).max() < 1e-14))
                    ojd, ocd = ref["copd"]
                    gotd = np.concatenate([p for p in dparts if p is not None and p.size] or [np.zeros(0, dtype=np.complex128)])
                    okc = okc and (ojd == jsecd) and (ojd == 0 or (gotd.size == ocd.size and np.abs(gotd - ocd).max() < 1e-14))
                    good = err < 1e-10 and erra < 1e-9 and errr < 1e-9 and p_eff == ref["p_eff"] and nd == ond and ndr == ondr and okc and eig_ok
                    ok &= bool(good)
                    print(f"{mdl.name} sector({nup},{ndw}) sparse={sparse} P={world} p_eff={p_eff} dim={dim} backend={mode} "
                          f"hxv_relerr={err:.2e} alpha_relerr={erra:.2e} alpha_realstart_relerr={errr:.2e} apply_op={'ok' if okc else 'BAD'} sp_eigh={'ok' if eig_ok else 'BAD'}({eig_err:.1e}) {'OK' if good else 'FAIL'}", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    E.ed_finalize()
    dist.destroy_process_group()
    if rank == 0:
        print("SPMD_CHECK", ",".join(SEL), "PASS" if ok else "FAIL", flush=True)
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
