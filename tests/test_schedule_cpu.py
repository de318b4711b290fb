"""CPU checks of the gather schedule the column-resident kernel (k_colres) executes: built on the host by
cdmft_b200_schedule_host from a CSR pattern, so it can be verified without a GPU.
  * every row is owned by exactly one lane of one warp task, every CSR entry (row, col, code) appears exactly
    once, in that lane;
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


def _check(rowptr, col, code, g, natural=False, nwarps=32):
    n = len(rowptr) - 1
    sched = E.schedule_host(rowptr, col, code, g, natural, nwarps)
    per = 32 // g
    ngroups = (n + g - 1) // g
    npad = ngroups * g
    seen = [[] for _ in range(n)]
    owned = np.zeros(n, int)
    nsteps = 0
    assert len(sched) == nwarps
    for tasks in sched:
        for rows, w in tasks:
            nsteps += len(w)
            for q in range(per):
                r8 = rows[q * g:(q + 1) * g]
                blk = w[:, q * g:(q + 1) * g]
                src = (blk >> 7).astype(np.int64)
                cd = blk & 127
                if not natural:  # conflict-free: g different banks in every step of every phase
                    banks = src % g
                    assert all(len(set(row)) == g for row in banks)
                idle = src >= n
                assert (src[idle] >= npad).all() and (src[idle] < npad + g).all() and (cd[idle] == 0).all()
                live = r8[r8 >= 0]
                owned[live] += 1
                if len(live):  # a lane group owns g consecutive rows (one 128-byte line of the output)
                    assert live[0] % g == 0 and (np.diff(live) == 1).all()
                assert idle[:, r8 < 0].all()
                for k, r in zip(*np.nonzero(~idle)):
                    seen[r8[r]].append((int(src[k, r]), int(cd[k, r])))
                if not natural and len(live):  # Koenig bound: steps used = max degree of the rows x banks multigraph
                    deg_r = max(rowptr[i + 1] - rowptr[i] for i in live)
                    bl = np.zeros(g, int)
                    for i in live:
                        bl += np.bincount(col[rowptr[i]:rowptr[i + 1]] % g, minlength=g)
                    assert int((~idle).any(axis=1).sum()) == max(deg_r, bl.max())
    assert (owned == 1).all()
    for i in range(n):
        want = sorted(zip(col[rowptr[i]:rowptr[i + 1]].tolist(), code[rowptr[i]:rowptr[i + 1]].tolist()))
        assert sorted(seen[i]) == want, i
    return nsteps


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
    _check(rowptr, col.astype(np.int64), code, g, nwarps=16)
    _check(rowptr, col.astype(np.int64), code, g, nwarps=3)


@pytest.mark.parametrize("g", [8, 16])
@pytest.mark.parametrize("cap", [40, 130, 400])
def test_block_split_schedules_of_a_sector_operator(g, cap):
    """Block-split schedules of k_colblk (columns larger than shared memory, Ns=18) for Hup of hm2x2(2) (Ns=12, 6
    particles, 924 rows), forced into small row blocks: blocks = runs of states sharing their top bits that cover the
    sector; every CSR entry is executed exactly once by the lane that owns its row -- in-block entries on the block's
    conflict-free schedule (relative indices), off-block entries in the lane-parallel stream whose steps each gather
    from ONE source block (neighbouring rows taking the same top-bit hop read neighbouring sources)."""
    mdl = models.hm2x2(2)
    o = edo.Oracle(mdl)
    o.build_hv_sector(models.get_sector(mdl.ns, 6, 6), edo.SPARSE_SERIAL)
    rowptr, col, val = o.get_csr(1)
    o.delete_hv_sector()
    rowptr = np.asarray(rowptr, dtype=np.int64)
    col = np.asarray(col, dtype=np.int64)
    if col.size and col.min() >= 1 and col.max() == len(rowptr) - 1:
        col = col - 1
    n = len(rowptr) - 1
    vals = np.asarray(val).reshape(len(col), -1)
    _, code = np.unique(vals.round(12), axis=0, return_inverse=True)
    code = (code.reshape(-1) % 127 + 1).astype(np.uint8)
    blocks = E.colblk_host(mdl.ns, 6, rowptr, col, code, g, False, cap)
    # blocks: contiguous cover, states of a block share their top bits, size bound met when the split allows it
    starts = [b["g0"] for b in blocks]
    assert starts[0] == 0 and all(blocks[k]["g0"] + blocks[k]["ng"] == (blocks[k + 1]["g0"] if k + 1 < len(blocks) else n)
                                   for k in range(len(blocks)))
    assert max(b["ng"] for b in blocks) <= max(cap, g)
    smap = np.asarray(edo.sector_map(mdl.ns, 6))
    block_of = np.zeros(n, int)
    for k, b in enumerate(blocks):
        block_of[b["g0"]:b["g0"] + b["ng"]] = k
    tbits = 0
    while len({int(s) >> (mdl.ns - tbits) for s in smap}) < len(blocks):
        tbits += 1
    for b in blocks:
        assert len({int(s) >> (mdl.ns - tbits) for s in smap[b["g0"]:b["g0"] + b["ng"]]}) == 1
    seen = [[] for _ in range(n)]
    owned = np.zeros(n, int)
    for k, b in enumerate(blocks):
        g0, ng = b["g0"], b["ng"]
        npad = (ng + g - 1) // g * g
        for rows, w, off in b["tasks"]:
            live = rows >= 0
            owned[g0 + rows[live]] += 1
            src = (w >> 7).astype(np.int64)
            cd = w & 127
            idle = src >= ng
            assert (src[idle] >= npad).all() and (src[idle] < npad + g).all() and (cd[idle] == 0).all()
            assert idle[:, ~live].all()
            for q in range(32 // g):  # conflict-free in-block steps
                banks = src[:, q * g:(q + 1) * g] % g
                assert all(len(set(r)) == g for r in banks)
            for s, lane in zip(*np.nonzero(~idle)):
                seen[g0 + rows[lane]].append((g0 + int(src[s, lane]), int(cd[s, lane])))
            on = (off & 127) != 0
            assert not on[:, ~live].any()
            for s in range(off.shape[0]):
                tgt = {int(block_of[j]) for j in (off[s][on[s]] >> 7)}
                assert len(tgt) <= 1 and k not in tgt   # one source block per step, never the own block
            for s, lane in zip(*np.nonzero(on)):
                seen[g0 + rows[lane]].append((int(off[s, lane] >> 7), int(off[s, lane] & 127)))
    assert (owned == 1).all()
    for i in range(n):
        want = sorted(zip(col[rowptr[i]:rowptr[i + 1]].tolist(), code[rowptr[i]:rowptr[i + 1]].tolist()))
        assert sorted(seen[i]) == want, i
