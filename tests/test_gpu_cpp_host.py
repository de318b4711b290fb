"""The compiled-language host mirror (include/ed_hamiltonian_b200.hpp) driven from C++ on the GPU and
checked against the oracle: same names and call sequence as the reference's ED_DIAG / ED_GF_NORMAL."""
import json
import os
import subprocess

import numpy as np
import pytest

from cdmft_lanc_ed_b200 import models

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build_driver(tmp_path):
    exe = str(tmp_path / "hxv_driver")
    pkg = os.path.join(ROOT, "cdmft_lanc_ed_b200")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "hxv_driver.cpp"),
                    "-o", exe, "-L", pkg, "-lcdmft_b200", f"-Wl,-rpath,{pkg}"], check=True)
    return exe


def test_cpp_driver_compiles_and_links(tmp_path):
    _build_driver(tmp_path)


@pytest.mark.gpu
def test_cpp_host_mirror_matches_oracle(tmp_path, oracle_lib):
    exe = _build_driver(tmp_path)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    r = json.loads(out)
    mdl = models.hm2x2(1)
    orc = oracle_lib.Oracle(mdl)
    isec = models.get_sector(8, 4, 4)
    orc.build_hv_sector(isec, oracle_lib.SPARSE_SERIAL)
    n = orc.dim
    i = np.arange(1, n + 1)
    v = np.sin(0.37 * i) + 1j * np.cos(0.11 * i)
    ref = orc.hxv(v)
    got = np.array([complex(a, b) for a, b in r["hv"]])
    assert np.abs(got - ref[::601][:8]).max() < 1e-10 * np.abs(ref).max()
    nd, a, b = orc.lanc_tridiag(v, 10)
    assert np.abs(np.array(r["alanc"]) - a).max() < 1e-10 * np.abs(a).max()
    assert np.abs(np.array(r["blanc"]) - b).max() < 1e-10 * np.abs(b).max()
    e0 = orc.lanc_eigh(512, 1e-14)[0]
    assert abs(r["e0"] - e0) < 1e-10 * abs(e0)
    assert r["threw_after_delete"] and r["n"] == 4900
    # build_Hv_sector(isector, Hmat), imp_kinetic, scatter / gather through the C++ mirror
    orc.delete_hv_sector()
    # sp_eigh (default LANC_METHOD) through the C++ mirror: two lowest pairs against the dense spectrum
    dref = np.linalg.eigvalsh(orc.dense_hmat(isec))
    assert r["sp_eigh_nconv"] == 2 and r["sp_eigh_residual"] < 1e-8
    assert abs(r["sp_eigh"][0] - dref[0]) < 1e-10 * abs(dref[0])
    assert np.abs(dref - r["sp_eigh"][1]).min() < 1e-10 * abs(dref[0]) and r["sp_eigh"][1] > r["sp_eigh"][0] - 1e-9
    ismall = models.get_sector(8, 2, 3)
    H = orc.dense_hmat(ismall)
    assert r["hmat_n"] == H.shape[0] and r["hmat_maxdiff"] < 1e-12 and r["scatter_gather_ok"]
    assert abs(r["hmat_trace"] - np.trace(H).real) < 1e-10 * max(1.0, abs(np.trace(H).real))
    j = np.arange(1, H.shape[0] + 1)
    w = np.cos(0.23 * j) + 1j * np.sin(0.31 * j)
    en = orc.lanc_local_energy(ismall, w, 1.0)
    diag_part = sum(mdl.imphloc[a, a, 0, 0, 0, 0].real for a in range(4))  # impHloc has no diagonal in this model
    assert diag_part == 0.0
    assert abs(r["imp_kinetic"] - en["Eknot"]) < 1e-10 * max(1.0, abs(en["Eknot"]))
