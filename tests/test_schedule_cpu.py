"""CPU checks of the gather schedule the column-resident kernel (k_colres) executes: built on the host by
cdmft_b200_schedule_host from a CSR pattern, so it can be verified without a GPU.
  * every CSR entry (row, col, code) appears exactly once, in the lane that owns the row;
  * edge colouring: inside one shared-memory phase (g consecutive lanes) the g gathers of a step hit g
    different banks (bank = source row mod g), idle lanes included (they read distinct zero elements);
  * the number of steps of a group is max(longest row, busiest bank) -- Koenig's bound is reached.
Patterns: the oracle's Hup of the BASELINE models (pattern parity with the reference is pinned elsewhere)
and random ragged patterns (empty rows, n not a multiple of g)."""
import numpy as np
import pytest

from cdmft_lanc_ed_b200 import ed_hamiltonian as E
from cdmft_lanc_ed_b200 import models
from oracle import edo


def _check(rowptr, col, code, g, natural=False):
    n = len(rowptr) - 1
    toff, tgrp, w = E.schedule_host(rowptr, col, code, g, natural)
    per = 32 // g
    ngroups = (n + g - 1) // g
    npad = ngroups * g
    assert sorted(x for x in tgrp if x >= 0) == list(range(ngroups))
    seen = [[] for _ in range(n)]
    for t in range(len(toff) - 1):
        assert (toff[t + 1] - toff[t]) % 4 == 0
        for q in range(per):
            grp = tgrp[t * per + q]
            blk = w[toff[t]:toff[t + 1], q * g:(q + 1) * g]
            src = (blk >> 7).astype(np.int64)
            cd = blk & 127
            if not natural:  # conflict-free: g different banks in every step of every phase
                banks = src % g
                assert all(len(set(row)) == g for row in banks), (t, q)
            idle = src >= n
            assert (src[idle] >= npad).all() and (src[idle] < npad + g).all() and (cd[idle] == 0).all()
            if grp < 0:
                assert idle.all()
                continue
            for k, r in zip(*np.nonzero(~idle)):
                i = grp * g + r
                assert i < n
                seen[i].append((int(src[k, r]), int(cd[k, r])))
            if not natural:  # Koenig bound: steps actually used = max degree of the rows x banks multigraph
                rows = range(grp * g, min(n, grp * g + g))
                deg_r = max((rowptr[i + 1] - rowptr[i] for i in rows), default=0)
                bl = np.zeros(g, int)
                for i in rows:
                    bl += np.bincount(col[rowptr[i]:rowptr[i + 1]] % g, minlength=g)
                used = int((~idle).any(axis=1).sum())
                assert used == max(deg_r, bl.max())
    for i in range(n):
        want = sorted(zip(col[rowptr[i]:rowptr[i + 1]].tolist(), code[rowptr[i]:rowptr[i + 1]].tolist()))
        assert sorted(seen[i]) == want, i
    return toff[-1]


@pytest.mark.parametrize("g", [8, 16])
@pytest.mark.parametrize("name", ["K1", "K2", "bhz_small"])
def test_schedule_of_model_patterns(name, g):
    mdl, sec = {"K1": (models.hm2x2(1), (4, 4)), "K2": (models.hm2x2(2), (6, 6)),
                "bhz_small": (models.bhz2(1), (2, 2))}[name]
    o = edo.Oracle(mdl)
    o.build_hv_sector(models.get_sector(mdl.ns, *sec), edo.SPARSE_SERIAL)
    rowptr, col, val = o.get_csr(1)
    o.delete_hv_sector()
    rowptr = np.asarray(rowptr, dtype=np.int64)
    col = np.asarray(col, dtype=np.int64)
    if col.size and col.min() >= 1 and col.max() == len(rowptr) - 1:
        col = col - 1  # oracle reports 1-based columns
    vals = np.asarray(val).reshape(len(col), -1)
    _, code = np.unique(vals.round(12), axis=0, return_inverse=True)
    code = (code.reshape(-1) % 127 + 1).astype(np.uint8)
    steps = _check(rowptr, col, code, g)
    steps_nat = _check(rowptr, col, code, g, natural=True)
    assert steps > 0 and steps_nat > 0


@pytest.mark.parametrize("g", [8, 16])
@pytest.mark.parametrize("n", [1, 7, 8, 33, 250])
def test_schedule_of_ragged_random_patterns(n, g):
    rng = np.random.default_rng(100 * n + g)
    lens = rng.integers(0, 12, size=n)
    lens[rng.integers(0, n)] = 0
    rowptr = np.concatenate([[0], np.cumsum(lens)])
    col = np.concatenate([np.sort(rng.choice(n, size=min(l, n), replace=False)) for l in lens] + [np.zeros(0, int)])
    rowptr = np.concatenate([[0], np.cumsum([min(l, n) for l in lens])])
    code = rng.integers(1, 128, size=len(col)).astype(np.uint8)
    _check(rowptr, col.astype(np.int64), code, g)
