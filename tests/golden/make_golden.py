"""Generates tests/golden/golden_small.npz from the CPU oracle (the reference cannot run here:
Fortran + MPI + SciFortran absent), after checking each case against the independent
Jordan-Wigner ED.  Re-run:  python tests/golden/make_golden.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cdmft_lanc_ed_b200 import models  # noqa: E402
from oracle import edo, jw_ed  # noqa: E402

CASES = [
    dict(key="hm2x2_nb1_44", builder="hm2x2", args=[1], nup=4, ndw=4),
    dict(key="hm2x2_nb1_54", builder="hm2x2", args=[1], nup=5, ndw=4),
    dict(key="bhz2_nb1_44", builder="bhz2", args=[1], nup=4, ndw=4),
    dict(key="bhz2_nb1_34", builder="bhz2", args=[1], nup=3, ndw=4),
]

out = {}
for c in CASES:
    mdl = getattr(models, c["builder"])(*c["args"])
    ns = mdl.ns
    c["isector"] = models.get_sector(ns, c["nup"], c["ndw"])
    orc = edo.Oracle(mdl)
    orc.build_hv_sector(c["isector"], edo.SPARSE_SERIAL)
    k = c["key"]
    out[k + "_map_up"] = edo.sector_map(ns, c["nup"])
    out[k + "_map_dw"] = edo.sector_map(ns, c["ndw"])
    for which, nm in ((1, "up"), (2, "dw")):
        rp, col, val = orc.get_csr(which)
        out[f"{k}_{nm}_rowptr"], out[f"{k}_{nm}_col"], out[f"{k}_{nm}_val"] = rp, col, val
    out[k + "_diag"] = orc.get_diag()
    rng = np.random.default_rng(12345)
    v = rng.normal(size=orc.dim) + 1j * rng.normal(size=orc.dim)
    v /= np.linalg.norm(v)
    out[k + "_v"] = v
    out[k + "_hv"] = orc.hxv(v)
    n, a, b = orc.lanc_tridiag(v, 30)
    out[k + "_alanc"], out[k + "_blanc"] = a, b
    orc.delete_hv_sector()
    if ns <= 5:
        Hs = jw_ed.sector_hamiltonian(mdl, c["nup"], c["ndw"])
        assert np.abs(Hs @ v - out[k + "_hv"]).max() < 1e-13
out["meta"] = json.dumps({"cases": CASES})
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_small.npz"), **out)
print("written", sum(v.nbytes for k, v in out.items() if k != "meta"), "bytes")
