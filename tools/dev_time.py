"""Developer timing sweep (not the bench contract): H x v variants on one GPU with CUDA events."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cdmft_lanc_ed_b200 import models  # noqa: E402
from cdmft_lanc_ed_b200 import ed_hamiltonian as E  # noqa: E402


def time_hxv(n, iters=10, warm=3):
    v = torch.randn(n, dtype=torch.complex128, device="cuda")
    v /= v.norm()
    hv = torch.empty_like(v)
    for _ in range(warm):
        E.spHtimesV_p(n, v, hv)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        E.spHtimesV_p(n, v, hv)
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


SWEEP = [dict(), dict(tma2d=0), dict(rowpass_variant=1), dict(row_slab=256), dict(row_slab=64), dict(rowres_cols=480)]
if os.environ.get("DEV_SWEEP"):
    SWEEP = [dict((kv.split("=")[0], int(kv.split("=")[1])) for kv in grp.split(",") if kv) for grp in os.environ["DEV_SWEEP"].split(";")]


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "K3"
    mdl, sec = {"K2": (models.hm2x2(2), (6, 6)), "K3": (models.hm2x2(3), (8, 8)), "K4": (models.bhz2(3), (8, 8))}[which]
    E.ed_init(0)
    E.set_stream(torch.cuda.current_stream().cuda_stream)
    E.ed_set_model(mdl)
    isec = models.get_sector(mdl.ns, *sec)
    for sparse in (True,):
        for opts in SWEEP:
            for k, v in dict(col_batch=4, force_sharded=0, colpass_variant=6, sched=1, rowpass_variant=4, rowres_cols=0, tma2d=1, row_rb=2, row_slab=128).items():
                E.set_option(k, v)
            for k, v in opts.items():
                E.set_option(k, v)
            t0 = time.time()
            n = E.build_Hv_sector(isec, sparse)
            tb = time.time() - t0
            ms = time_hxv(n)
            E.set_option("profile", 1)
            time_hxv(n, iters=5, warm=0)
            kt = {k: round(E.profile_query(k)[0] / 5, 3) for k in (0, 1, 2)}
            E.set_option("profile", 0)
            # one Krylov run (20 steps) with the per-kind timers: real vectors when H is real
            E.set_option("profile", 1)
            E.profile_query(0), E.profile_query(1), E.profile_query(4)
            v0 = torch.ones(n, dtype=torch.complex128, device="cuda")
            torch.cuda.synchronize()
            t0 = time.time()
            E.sp_lanc_tridiag(v0, 20)
            tl = (time.time() - t0) / 20 * 1e3
            kl = {k: round(E.profile_query(k)[0] / 20, 3) for k in (0, 1, 4)}
            E.set_option("profile", 0)
            del v0
            E.delete_Hv_sector()
            opts = dict(opts, lanc_ms_per_iter=round(tl, 3), lanc_col=kl[0], lanc_row=kl[1], lanc_vec=kl[4])
            print(json.dumps(dict(cfg=which, sparse=sparse, **opts, n=n, build_s=round(tb, 3), ms=round(ms, 4),
                                  gbs_alg=round(32 * n / ms / 1e6, 1), col_ms=kt[0], row_ms=kt[1], tr_ms=kt[2])), flush=True)
    E.ed_finalize()


if __name__ == "__main__":
    main()
