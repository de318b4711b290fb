"""Generates tests/golden/k3_gs_oracle.json: ground-state Lanczos of BASELINE config K3 (hm2x2 Nbath=3, Ns=16, sector (8,8),
Dim 165 636 900) by the CPU ORACLE (sp_lanc_eigh restatement around the spMatVec_mpi_main restatement, constant start vector,
threshold 1e-12, ncheck 10) -- a few minutes on 8 cores, ~20 GB of RAM.  bench.py and tests/test_gpu_fullsize.py compare the
GPU driver's E0, iteration count and leading alpha / beta with this file.  Re-run:  python tests/golden/make_k3_gs.py"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cdmft_lanc_ed_b200 import models  # noqa: E402
from oracle import edo  # noqa: E402

mdl = models.hm2x2(3)
isec = models.get_sector(16, 8, 8)
cores = os.cpu_count() or 1
edo.lib().edo_set_num_threads(cores)
orc = edo.Oracle(mdl)
orc.build_hv_sector(isec, edo.SPARSE_MPI, cores)
t0 = time.time()
e0, vec, nit, al, bl = orc.lanc_eigh(512, 1e-12)
dt = time.time() - t0
out = {"model": mdl.name, "sector": [8, 8], "dim": int(orc.dim), "threshold": 1e-12, "ncheck": 10, "start": "constant 1/sqrt(Dim)",
       "e0": e0, "iterations": int(nit), "alanc": [float(x) for x in al[:40]], "blanc": [float(x) for x in bl[:40]],
       "oracle_seconds": dt, "oracle_cores": cores,
       "vec_abs_max": float(abs(vec).max()), "vec_imag_abs_max": float(abs(vec.imag).max())}
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "k3_gs_oracle.json"), "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps({k: out[k] for k in ("e0", "iterations", "oracle_seconds")}))
