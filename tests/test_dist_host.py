"""Host-side logic of the multi-GPU path, CPU only: SFC assignment, halo discovery, local layout and the
halo exchange plan, single-process for all ranks and with a world_size-2 gloo group for the N>1 plumbing."""
import os

import numpy as np
import pytest

import sphexa_b200 as sx
from sphexa_b200 import dist as sdist
from sphexa_b200 import host


def jittered(side, seed=3):
    rng = np.random.default_rng(seed)
    c = -0.5 + (np.arange(side) + 0.5) / side
    z, y, x = np.meshgrid(c, c, c, indexing="ij")
    p = np.stack([x.ravel(), y.ravel(), z.ravel()], 1) + rng.uniform(-0.3, 0.3, (side ** 3, 3)) / side
    p = np.where(p >= 0.5, p - 1.0, np.where(p < -0.5, p + 1.0, p))
    h = (0.5 * np.cbrt(3 / (4 * np.pi) * 60 / side ** 3) * rng.uniform(0.8, 1.3, side ** 3)).astype(np.float32)
    return p[:, 0].copy(), p[:, 1].copy(), p[:, 2].copy(), h


BOX = [-0.5, 0.5] * 3


def sorted_global(side, boundary):
    x, y, z, h = jittered(side)
    t = host.build_tree(x, y, z, BOX, boundary, 64)
    o = t.order
    return x[o], y[o], z[o], h[o], t.keys


@pytest.mark.parametrize("nranks", [1, 2, 3, 8])
def test_sfc_assignment_balanced(nranks):
    xs, ys, zs, hs, keys = sorted_global(16, [1, 1, 1])
    n = keys.size
    bucket = max(64, n // (100 * nranks))
    splits = sdist.sfc_assignment(keys, nranks, bucket)
    assert splits[0] == 0 and splits[-1] == n and np.all(np.diff(splits) >= 0)
    counts = np.diff(splits)
    assert np.abs(counts - n / nranks).max() <= bucket  # boundaries move at most one global-tree leaf
    # rank boundaries never split particles with equal keys, and sit on octree-node boundaries of the bucket tree
    for s in splits[1:-1]:
        assert keys[s - 1] < keys[s]


@pytest.mark.parametrize("boundary", [[1, 1, 1], [0, 0, 0], [1, 0, 1]])
def test_halos_complete_and_plan_consistent(boundary):
    nranks = 3
    xs, ys, zs, hs, keys = sorted_global(14, boundary)
    n = keys.size
    splits = sdist.sfc_assignment(keys, nranks, 64)
    sets, halos = [], []
    for r in range(nranks):
        hl = sdist.find_halos(xs, ys, zs, hs, BOX, boundary, int(splits[r]), int(splits[r + 1]))
        halos.append(hl)
        sets.append(sdist.local_set(r, nranks, splits, hl))
    pos = np.stack([xs, ys, zs], 1)
    L = np.where(np.array(boundary) == 1, 1.0, 0.0)
    for r, ls in enumerate(sets):
        assert np.all(np.diff(ls.local_idx) > 0)                       # local arrays stay SFC-sorted
        assert np.array_equal(ls.local_idx[ls.first:ls.last], np.arange(splits[r], splits[r + 1]))
        present = np.zeros(n, bool)
        present[ls.local_idx] = True
        # completeness: every particle within 2 h_i of an assigned particle i is held locally
        for i in range(int(splits[r]), int(splits[r + 1]), 7):
            d = pos - pos[i]
            d -= L * np.rint(d / np.where(L > 0, L, 1.0)) * (L > 0)
            nb = np.nonzero((d * d).sum(1) < (2.0 * float(hs[i])) ** 2)[0]
            assert present[nb].all()
    plans = [sdist.halo_plan(ls, halos) for ls in sets]
    for r, (ls, pl) in enumerate(zip(sets, plans)):
        received = np.zeros(ls.local_idx.size, bool)
        for q, p in enumerate(pl.peers):
            # what r receives from p must be exactly what p sends to r, in the same order
            qp = list(plans[p].peers).index(r)
            sent = plans[p].send_idx[plans[p].send_offsets[qp]:plans[p].send_offsets[qp + 1]]
            sent_global = sets[p].local_idx[sent]
            rb, rc = int(pl.recv_begin[q]), int(pl.recv_count[q])
            assert np.array_equal(ls.local_idx[rb:rb + rc], sent_global)
            assert np.all((sent >= sets[p].first) & (sent < sets[p].last))  # only assigned particles are sent
            received[rb:rb + rc] = True
        # every halo slot is filled by exactly one peer, no assigned slot is overwritten
        assert received[:ls.first].all() and received[ls.last:].all() and not received[ls.first:ls.last].any()


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        boundary = [1, 1, 1]
        xs, ys, zs, hs, keys = sorted_global(12, boundary)
        splits = sdist.sfc_assignment(keys, world, 64)
        hl = sdist.find_halos(xs, ys, zs, hs, BOX, boundary, int(splits[rank]), int(splits[rank + 1]))
        ls = sdist.local_set(rank, world, splits, hl)
        all_halos = [None] * world
        dist.all_gather_object(all_halos, hl)
        pl = sdist.halo_plan(ls, all_halos)
        # the exchange semantics of sphx_halo_exchange, on CPU tensors over gloo: gather-pack, send, receive in place
        field = torch.full((ls.local_idx.size,), -1.0, dtype=torch.float64)
        field[ls.first:ls.last] = torch.from_numpy(ls.local_idx[ls.first:ls.last].astype(np.float64))
        reqs, bufs = [], []
        for qi, p in enumerate(pl.peers):
            idx = torch.from_numpy(pl.send_idx[pl.send_offsets[qi]:pl.send_offsets[qi + 1]].astype(np.int64))
            bufs.append(field[idx].contiguous())
            reqs.append(dist.isend(bufs[-1], int(p)))
            rb, rc = int(pl.recv_begin[qi]), int(pl.recv_count[qi])
            reqs.append(dist.irecv(field[rb:rb + rc], int(p)))
        for r_ in reqs:
            r_.wait()
        ok = bool(np.array_equal(field.numpy(), ls.local_idx.astype(np.float64)))
        q.put((rank, ok, int(ls.local_idx.size), int(ls.last - ls.first)))
    finally:
        dist.destroy_process_group()


def test_halo_plan_over_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _ in res)
    assert sum(na for *_, na in res) == 12 ** 3


def test_dist_symbols_exported():
    L = sx.load()
    for name in ["sphx_comm_unique_id", "sphx_comm_init", "sphx_comm_free", "sphx_halo_exchange", "sphx_allreduce_f64",
                 "sphx_hydro_step_dist", "sphx_sfc_assignment_host", "sphx_find_halos_host"]:
        assert hasattr(L, name)
