"""CPU tests of the dynamic decomposition plan (sphx_cell_plan_build_host, the host half of the multi-rank
Domain::sync): assignment balance, halo completeness against brute force, and consistency of every rank's send lists
with its peers' receive ranges. The plan is a pure function of the global histogram, so all ranks are built here in
one process; tests/test_dist_host.py::test_gloo_* cover the same logic behind a real process group."""
import numpy as np
import pytest

import sphexa_b200 as sx
from sphexa_b200 import dist as sdist
from sphexa_b200 import host


def _setup(n, boundary, seed, clustered):
    rng = np.random.default_rng(seed)
    if clustered:
        p = np.concatenate([rng.normal(0.1, 0.08, (3, n // 2)), rng.uniform(-0.5, 0.5, (3, n - n // 2))], axis=1)
        p = np.clip(p, -0.5, 0.5 - 1e-9)
    else:
        p = rng.uniform(-0.5, 0.5, (3, n))
    box = [-0.5, 0.5] * 3
    keys = host.hilbert_keys(p[0], p[1], p[2], box, boundary)
    o = np.argsort(keys, kind="stable")
    return p[:, o], keys[o], box


@pytest.mark.parametrize("nranks", [2, 3, 8])
@pytest.mark.parametrize("boundary,clustered", [([1, 1, 1], False), ([0, 0, 0], True), ([1, 0, 1], True)])
@pytest.mark.parametrize("variable_h", [False, True])
def test_cell_plan_is_complete_and_consistent(nranks, boundary, clustered, variable_h):
    n = 60000
    (x, y, z), keys, box = _setup(n, boundary, 11 + nranks, clustered)
    level = 4
    edge = 1.0 / (1 << level)
    cells = (keys >> np.uint64(3 * (21 - level))).astype(np.int64)
    G = np.bincount(cells, minlength=8 ** level).astype(np.uint32)
    if variable_h:
        # smoothing lengths between 0.2 and 1.4 cell edges, smooth in space: cells reach 1 - 3 rings (per-cell reach =
        # ceil(2 max h of the cell / edge), as DistributedSimulation.sync computes it)
        h = (edge * (0.2 + 1.2 * (0.5 + 0.5 * np.sin(7.0 * x) * np.cos(5.0 * y + 3.0 * z)))).astype(np.float32)
        hcell = np.zeros(8 ** level, np.float32)
        np.maximum.at(hcell, cells, h)
        rings = np.clip(np.ceil(hcell.astype(np.float64) * (2.0 * 1.0001 / edge)), 1, 3).astype(np.uint8)
        assert rings.min() == 1 and rings.max() == 3
    else:
        h = np.full(n, 0.5 * edge * 0.999, np.float32)  # 2h just below the cell edge
        assert sdist.cell_level(box, float(h.max())) == level
        rings = None
    plans = [sdist.cell_plan(G, level, boundary, r, nranks, rings=rings) for r in range(nranks)]

    # assignment: identical on all ranks, partitions the cells, balanced to within one cell
    for pl in plans:
        assert np.array_equal(pl.cell_splits, plans[0].cell_splits)
    sp = plans[0].cell_splits.astype(np.int64)
    assert sp[0] == 0 and sp[-1] == 8 ** level and np.all(np.diff(sp) >= 0)
    assert sum(pl.n_assigned for pl in plans) == n
    assert max(abs(pl.n_assigned - n / nranks) for pl in plans) <= G.max() + 1
    pstart = np.searchsorted(cells, sp)  # first global (sorted) particle of each rank

    def local_to_global(r):
        """global sorted index of every local particle of rank r in the layout [halos | assigned | halos]"""
        pl = plans[r]
        cell_first = np.searchsorted(cells, pl.recv_cells.astype(np.int64))
        halo = np.concatenate([np.arange(f, f + G[c]) for f, c in zip(cell_first, pl.recv_cells)] or
                              [np.zeros(0, np.int64)]).astype(np.int64)
        left, right = halo[halo < pstart[r]], halo[halo >= pstart[r + 1]]
        assert left.size == pl.n_halo_left and right.size == pl.n_halo_right
        return np.concatenate([left, np.arange(pstart[r], pstart[r + 1]), right])

    l2g = [local_to_global(r) for r in range(nranks)]
    L = np.array([1.0, 1.0, 1.0])
    P = np.stack([x, y, z], 1)
    rng = np.random.default_rng(0)
    for r, pl in enumerate(plans):
        have = np.zeros(n, bool)
        have[l2g[r]] = True
        # halo completeness: every particle within 2h of a sampled assigned particle is local
        if pl.n_assigned:
            for i in rng.integers(pstart[r], pstart[r + 1], 60):
                d = P - P[i]
                for k in range(3):
                    if boundary[k] == 1:
                        d[:, k] -= L[k] * np.rint(d[:, k] / L[k])
                nb = np.nonzero((d ** 2).sum(1) < (2.0 * h[i]) ** 2)[0]
                assert have[nb].all(), (r, i)
        # exchange consistency: what r sends to q is exactly, and in the same order, what q expects from r
        for k, q in enumerate(pl.peers):
            sent_global = l2g[r][pl.send_idx[pl.send_offsets[k]:pl.send_offsets[k + 1]]]
            pq = plans[q]
            kk = int(np.nonzero(pq.peers == r)[0][0])
            b, c = int(pq.recv_begin[kk]), int(pq.recv_count[kk])
            assert np.array_equal(sent_global, l2g[q][b:b + c]), (r, q)
        # every halo particle is received from exactly one peer
        assert int(pl.recv_count.sum()) == pl.n_halo_left + pl.n_halo_right


def test_cell_plan_edge_cases():
    # one rank: no peers, no halos
    G = np.zeros(8, np.uint32)
    G[3] = 10
    pl = sdist.cell_plan(G, 1, [1, 1, 1], 0, 1)
    assert pl.peers.size == 0 and pl.n_assigned == 10 and pl.n_local == 10
    # more ranks than non-empty cells: trailing ranks own nothing and talk to nobody
    plans = [sdist.cell_plan(G, 1, [0, 0, 0], r, 4) for r in range(4)]
    assert sum(p.n_assigned for p in plans) == 10
    assert all(p.peers.size == 0 for p in plans)
    # level 0: a single cell
    pl = sdist.cell_plan(np.array([7], np.uint32), 0, [1, 1, 1], 0, 2)
    assert pl.n_global == 7
    assert sdist.cell_level([-0.5, 0.5] * 3, 0.3) == 0
    assert sdist.cell_level([-0.5, 0.5] * 3, 0.0036) == 7


@pytest.mark.parametrize("level,boundary,nranks,max_ring", [(2, [1, 1, 1], 3, 3), (3, [1, 0, 1], 4, 4), (3, [0, 0, 0], 5, 2),
                                                            (1, [1, 1, 1], 2, 2)])
def test_cell_plan_rings_against_brute_force(level, boundary, nranks, max_ring):
    """per-cell reach on small grids, where a reach of several rings wraps around the periodic box more than once: the
    halo cells of every rank against the definition (non-empty foreign cells within rings[c] rings, Chebyshev distance
    with the minimum image on periodic axes, of an own non-empty cell c), and what a rank sends against what its peers
    expect"""
    side = 1 << level
    ncell = side ** 3
    rng = np.random.default_rng(level * 10 + nranks)
    # cell index of every grid position (the cells are Hilbert-ordered)
    g = (np.arange(side) + 0.5) / side - 0.5
    gx, gy, gz = (a.ravel() for a in np.meshgrid(g, g, g, indexing="ij"))
    keys = host.hilbert_keys(gx, gy, gz, [-0.5, 0.5] * 3, boundary)
    cell = (keys >> np.uint64(3 * (21 - level))).astype(np.int64)
    pos = np.zeros((ncell, 3), np.int64)
    ijk = np.stack(np.meshgrid(np.arange(side), np.arange(side), np.arange(side), indexing="ij"), -1).reshape(-1, 3)
    pos[cell] = ijk
    G = rng.integers(0, 9, ncell).astype(np.uint32)
    G[rng.random(ncell) < 0.25] = 0
    rings = rng.integers(1, max_ring + 1, ncell).astype(np.uint8)
    plans = [sdist.cell_plan(G, level, boundary, r, nranks, rings=rings) for r in range(nranks)]
    sp = plans[0].cell_splits.astype(np.int64)
    owner = np.searchsorted(sp[1:-1], np.arange(ncell), side="right")
    d = np.abs(pos[:, None, :] - pos[None, :, :])
    for k in range(3):
        if boundary[k] == 1:
            d[:, :, k] = np.minimum(d[:, :, k], side - d[:, :, k])
    cheb = d.max(-1)  # cheb[c, c2]
    nonempty = G > 0
    for r, pl in enumerate(plans):
        own = (owner == r) & nonempty
        need = ((cheb <= rings[:, None]) & own[:, None] & nonempty[None, :] & (owner != r)[None, :]).any(0)
        np.testing.assert_array_equal(pl.recv_cells, np.nonzero(need)[0])
        assert pl.n_halo_left + pl.n_halo_right == int(G[need].sum())
        # send lists: cell c of rank r goes to rank q iff some non-empty cell of q has c within its reach
        for k, q in enumerate(pl.peers):
            wants = ((cheb <= rings[:, None]) & ((owner == q) & nonempty)[:, None]).any(0) & own
            n_send = int(pl.send_offsets[k + 1] - pl.send_offsets[k])
            assert n_send == int(G[wants].sum()), (r, q)
            kk = np.nonzero(plans[q].peers == r)[0]
            assert kk.size == 1 and int(plans[q].recv_count[kk[0]]) == n_send
