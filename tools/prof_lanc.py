"""Minimal real-vector Krylov run for profiling (k_colres2 / k_colblk<double>): python tools/prof_lanc.py [K3|NS18] [steps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cdmft_lanc_ed_b200 import models  # noqa: E402
from cdmft_lanc_ed_b200 import ed_hamiltonian as E  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "K3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
mdl, sec = {"K3": (models.hm2x2(3), (8, 8)), "NS18": (models.hm_ns18(), (9, 4))}[which]
E.ed_init(0)
E.set_stream(torch.cuda.current_stream().cuda_stream)
E.ed_set_model(mdl)
n = E.build_Hv_sector(models.get_sector(mdl.ns, *sec), True)
v0 = torch.ones(n, dtype=torch.complex128, device="cuda")
print(E.sp_lanc_tridiag(v0, steps)[1][:steps])
E.delete_Hv_sector()
E.ed_finalize()
