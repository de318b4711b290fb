"""Developer timing of the ground-state Lanczos driver (not the bench contract): python tools/lanc_time.py [K3] [key=value ...]"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cdmft_lanc_ed_b200 import models  # noqa: E402
from cdmft_lanc_ed_b200 import ed_hamiltonian as E  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "K3"
mdl, sec = {"K2": (models.hm2x2(2), (6, 6)), "K3": (models.hm2x2(3), (8, 8)), "K4": (models.bhz2(3), (8, 8))}[which]
E.ed_init(0)
E.set_stream(torch.cuda.current_stream().cuda_stream)
E.ed_set_model(mdl)
for kv in sys.argv[2:]:
    k, v = kv.split("=")
    E.set_option(k, int(v))
n = E.build_Hv_sector(models.get_sector(mdl.ns, *sec), True)
for rep in range(3):
    vec = torch.zeros(n, dtype=torch.complex128, device="cuda")
    E.set_option("profile", 1)
    for k in (0, 1, 4):
        E.profile_query(k)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, nit, _, _ = E.sp_lanc_eigh(vec, 512, 1e-12)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    kt = {k: E.profile_query(k)[0] for k in (0, 1, 4)}
    E.set_option("profile", 0)
    print(json.dumps(dict(cfg=which, rep=rep, seconds=round(dt, 4), iterations=nit, e0=e0, ms_per_step=round(dt / (2 * nit) * 1e3, 3),
                          col_ms=round(kt[0] / (2 * nit), 3), row_ms=round(kt[1] / (2 * nit), 3), vec_ms=round(kt[4] / (2 * nit), 3))), flush=True)
    del vec
E.delete_Hv_sector()
E.ed_finalize()
