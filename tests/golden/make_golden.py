"""Generates tests/golden/golden_small.npz from the CPU oracle (the reference cannot run here:
Fortran + MPI + SciFortran absent).  The Ns = 4 cases (jw = true in the meta record) are checked against the
independent dense Jordan-Wigner ED (oracle/jw_ed.py) at generation time, and the script refuses to write the file
if one of them fails; the Ns = 8 cases are too large for the dense 4^Ns construction and are plain oracle dumps
(the oracle itself is pinned on small models in tests/test_oracle_pin.py).  Re-run:  python tests/golden/make_golden.py"""
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
    # Ns = 4: small enough for the dense Jordan-Wigner ED -> really cross-checked below
    dict(key="hub2x1_nb1_22", builder="hubbard_cluster", args=[2, 1, 1], nup=2, ndw=2),
    dict(key="rand_L1O2B1_kanamori_22", builder="random_model", args=[1, 2, 1], kwargs=dict(seed=2, kanamori=True), nup=2, ndw=2),
    dict(key="rand_L2O1B1_S2_21", builder="random_model", args=[2, 1, 1], kwargs=dict(nspin=2, seed=3), nup=2, ndw=1),
    dict(key="rand_L1O1B3_nohf_13", builder="random_model", args=[1, 1, 3], kwargs=dict(seed=5, hfmode=False), nup=1, ndw=3),
]

out = {}
for c in CASES:
    mdl = getattr(models, c["builder"])(*c["args"], **c.get("kwargs", {}))
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
    c["jw"] = bool(ns <= 5)
    if c["jw"]:
        Hs = jw_ed.sector_hamiltonian(mdl, c["nup"], c["ndw"])
        assert np.abs(Hs @ v - out[k + "_hv"]).max() < 1e-13, k
out["meta"] = json.dumps({"cases": CASES})
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_small.npz"), **out)
assert sum(c["jw"] for c in CASES) >= 4
print("written", sum(v.nbytes for k, v in out.items() if k != "meta"), "bytes")
