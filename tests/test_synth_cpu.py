"""CPU: the counter-based synthetic vector is the same function in C (oracle), numpy and torch, and the oracle's
sampled rows of H x v against it equal the rows of the full oracle product (complex hoppings, Kanamori terms,
quirky sectors included) -- the check the Ns=18 sector relies on, where no full host vector can exist."""
import numpy as np
import pytest

from cdmft_lanc_ed_b200 import models, synth
from oracle import edo


def test_counter_vector_c_numpy_torch_bitwise():
    import torch
    for i0, n, seed in [(0, 1000, 1), (123456789012, 777, 42), (2 ** 40 + 5, 64, 2 ** 33 + 7)]:
        sc = synth.default_scale(10 ** 6)
        a = edo.counter_vec(i0, n, seed, sc)
        b = synth.counter_vec_numpy(i0, n, seed, sc)
        c = synth.counter_vec_torch(i0, n, seed, sc, device="cpu", chunk=300).numpy()
        assert np.array_equal(a, b) and np.array_equal(a, c)
        assert np.abs(a.real).max() <= sc and np.abs(a.imag).max() <= sc
    v = synth.counter_vec_numpy(0, 200000, 3, 1.0)
    assert abs(v.real.mean()) < 0.01 and abs((np.abs(v) ** 2).mean() - 2 / 3) < 0.01


@pytest.mark.parametrize("mdl,sec", [(models.hm2x2(1), (4, 4)), (models.bhz2(1), (4, 3)), (models.hm2x2(2), (5, 6)),
                                     (models.random_model(2, 2, 1, nspin=2, seed=12), (3, 4)),
                                     (models.random_model(2, 2, 1, seed=15, kanamori=True), (4, 4)),
                                     (models.random_model(1, 3, 1, seed=14), (2, 3))])
def test_sampled_rows_equal_full_product(mdl, sec):
    orc = edo.Oracle(mdl)
    isec = models.get_sector(mdl.ns, *sec)
    for kind in (edo.SPARSE_SERIAL, edo.DIRECT_SERIAL):
        orc.build_hv_sector(isec, kind)
        dim = orc.dim
        sc = synth.default_scale(dim)
        v = edo.counter_vec(0, dim, 11, sc)
        ref = orc.hxv(v)
        rows = np.unique(np.random.default_rng(0).integers(0, dim, size=min(dim, 500)))
        got = orc.hxv_rows_counter(rows, 11, sc)
        assert np.abs(got - ref[rows]).max() <= 1e-13 * max(np.abs(ref).max(), 1e-300)
        orc.delete_hv_sector()
