"""Minimal H x v run for profiling: python tools/prof_hxv.py [K2|K3|K4] [sparse|direct] [n_iter] [key=value ...]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cdmft_lanc_ed_b200 import models  # noqa: E402
from cdmft_lanc_ed_b200 import ed_hamiltonian as E  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "K3"
sparse = (sys.argv[2] if len(sys.argv) > 2 else "sparse") == "sparse"
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
mdl, sec = {"K2": (models.hm2x2(2), (6, 6)), "K3": (models.hm2x2(3), (8, 8)), "K4": (models.bhz2(3), (8, 8))}[which]
E.ed_init(0)
E.set_stream(torch.cuda.current_stream().cuda_stream)
E.ed_set_model(mdl)
for kv in sys.argv[4:]:
    k, v = kv.split("=")
    E.set_option(k, int(v))
n = E.build_Hv_sector(models.get_sector(mdl.ns, *sec), sparse)
v = torch.randn(n, dtype=torch.complex128, device="cuda")
hv = torch.empty_like(v)
for _ in range(iters):
    E.spHtimesV_p(n, v, hv)
torch.cuda.synchronize()
print("done", n, float(hv.abs().sum()))
E.delete_Hv_sector()
E.ed_finalize()
