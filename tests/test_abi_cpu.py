"""CPU checks of the boundary: the C-ABI library loads, exports every symbol include/cdmft_b200.h
declares, and fails loudly (no CPU fallback) when there is no CUDA device."""
import os
import re

import pytest

from cdmft_lanc_ed_b200 import ed_hamiltonian as E

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "cdmft_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(cdmft_b200_[a-z0-9_]+)\s*\(", hdr)))
    assert declared, "no declarations found"
    L = E.load_library()
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert set(declared) == set(E.ABI_SYMBOLS)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(E.EdB200Error, match="no CUDA device|CUDA"):
        E.ed_init(0)
    with pytest.raises(E.EdB200Error):
        E.build_Hv_sector(1)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "cdmft_lanc_ed_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "ed_oracle" not in txt, f
