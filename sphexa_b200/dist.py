"""SFC domain decomposition over the GPUs of one node and the NCCL halo exchange plan.

Host logic only (numpy + the C ABI); the bytes move inside libsphx (csrc/dist.cu: gather-pack kernel + grouped
ncclSend/ncclRecv). Mirrors, for a static particle set, what cstone::Domain::sync produces for the hot path
(domain/include/cstone/domain/domain.hpp:181-234): the assigned SFC range of each rank (domaindecomp.hpp:33-110),
its halo particles (halos/halos.hpp:131-192), the local array layout [halos | assigned | halos] in SFC order
(domain/layout.hpp:150-163), and the per-peer send lists / receive ranges (halos/halos.hpp:234-254).

The reference discovers all of this with peer-to-peer MPI messages over a focus tree because no rank holds the global
particle set. Round 1 builds the decomposition of SYNTHETIC initial conditions, which every rank can generate in
full, so each rank derives its local set from the global arrays; the exchange plan is still assembled from the
halo sets the ranks publish to each other (gather over the host process group), as Domain does.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _cabi, host


def _p(a):
    return C.c_void_p(a.ctypes.data)


def sfc_assignment(sorted_keys: np.ndarray, nranks: int, bucket: int) -> np.ndarray:
    L = _cabi.load()
    keys = np.ascontiguousarray(sorted_keys, np.uint64)
    splits = np.zeros(nranks + 1, np.uint64)
    _cabi.check(L.sphx_sfc_assignment_host(_p(keys), keys.size, nranks, bucket, _p(splits)))
    return splits.astype(np.int64)


def find_halos(xs, ys, zs, hs, box_lim, boundary, owned_begin: int, owned_end: int, bucket: int = 64) -> np.ndarray:
    """global (SFC-sorted) indices of the halo particles of the rank owning [owned_begin, owned_end)"""
    L = _cabi.load()
    flags = np.zeros(xs.size, np.uint8)
    b = host.make_box(box_lim, boundary)
    _cabi.check(L.sphx_find_halos_host(_p(xs), _p(ys), _p(zs), _p(hs), xs.size, C.byref(b), bucket, owned_begin,
                                       owned_end, _p(flags)))
    return np.nonzero(flags)[0].astype(np.int64)


@dataclass
class LocalSet:
    """what one rank holds: indices into the global SFC-sorted arrays, ascending (= local SFC order)"""
    rank: int
    nranks: int
    splits: np.ndarray      # nranks + 1
    halos: np.ndarray       # global indices of this rank's halo particles
    local_idx: np.ndarray   # halos U assigned, ascending
    first: int              # Domain::startIndex
    last: int               # Domain::endIndex


def local_set(rank: int, nranks: int, splits: np.ndarray, halos: np.ndarray) -> LocalSet:
    b, e = int(splits[rank]), int(splits[rank + 1])
    lo = halos[halos < b]
    hi = halos[halos >= e]
    local_idx = np.concatenate([lo, np.arange(b, e, dtype=np.int64), hi])
    return LocalSet(rank, nranks, splits, halos, local_idx, int(lo.size), int(lo.size) + (e - b))


@dataclass
class HaloPlanHost:
    peers: np.ndarray        # int32
    send_offsets: np.ndarray  # uint32, numPeers + 1
    send_idx: np.ndarray     # uint32 local indices
    recv_begin: np.ndarray   # uint32
    recv_count: np.ndarray   # uint32


def halo_plan(ls: LocalSet, all_halos: list[np.ndarray]) -> HaloPlanHost:
    """send lists / receive ranges of rank ls.rank given every rank's halo set (global indices, ascending)"""
    r = ls.rank
    b, e = int(ls.splits[r]), int(ls.splits[r + 1])
    peers, send_chunks, recv_begin, recv_count = [], [], [], []
    for p in range(ls.nranks):
        if p == r:
            continue
        hp = all_halos[p]
        mine = hp[(hp >= b) & (hp < e)]                    # my assigned particles that are halos of p
        pb, pe = int(ls.splits[p]), int(ls.splits[p + 1])
        theirs = ls.halos[(ls.halos >= pb) & (ls.halos < pe)]  # my halos that p owns: contiguous in local order
        if mine.size == 0 and theirs.size == 0:
            continue
        peers.append(p)
        send_chunks.append((ls.first + (mine - b)).astype(np.uint32))
        if theirs.size:
            pos = int(np.searchsorted(ls.local_idx, theirs[0]))
            assert np.array_equal(ls.local_idx[pos:pos + theirs.size], theirs)
            recv_begin.append(pos)
        else:
            recv_begin.append(0)
        recv_count.append(theirs.size)
    offs = np.zeros(len(peers) + 1, np.uint32)
    if peers:
        offs[1:] = np.cumsum([c.size for c in send_chunks])
    send_idx = np.concatenate(send_chunks) if send_chunks else np.zeros(0, np.uint32)
    return HaloPlanHost(np.array(peers, np.int32), offs, send_idx.astype(np.uint32),
                        np.array(recv_begin, np.uint32), np.array(recv_count, np.uint32))


class DistributedHydro:
    """One rank of a multi-GPU hydro step: local fields + tree on the device, NCCL communicator, halo plan.

    `pg` is a torch.distributed process group used for host-side plumbing only (publishing halo sets, the NCCL
    unique id); the data path is sphx_hydro_step_dist in libsphx."""

    MAX_EXCHANGE_ARRAYS = 7  # c11..c33 + divv (ve_hydro.hpp:174)

    def __init__(self, sim, glob: dict, rank: int, nranks: int, device, pg=None, bucket_focus: int = 64):
        import torch
        import torch.distributed as dist

        self.L = _cabi.load()
        self.rank, self.nranks = rank, nranks
        x, y, z = glob["x"], glob["y"], glob["z"]
        box_lim, boundary, params = glob["box"], glob["boundary"], glob["params"]
        n = x.size
        # global SFC order (identical on every rank: same input, same code)
        t = host.build_tree(x, y, z, box_lim, boundary, bucket_focus)
        o = t.order
        xs, ys, zs = x[o], y[o], z[o]

        def sorted_field(v, dtype):
            return (v[o] if isinstance(v, np.ndarray) and v.shape == (n,) else np.full(n, v)).astype(dtype)

        hs = sorted_field(glob["fields"]["h"], np.float32)
        bucket_global = max(bucket_focus, n // (100 * nranks))
        self.splits = sfc_assignment(t.keys, nranks, bucket_global)
        halos = find_halos(xs, ys, zs, hs, box_lim, boundary, int(self.splits[rank]), int(self.splits[rank + 1]),
                           bucket_focus)
        self.ls = local_set(rank, nranks, self.splits, halos)
        if nranks > 1:
            all_halos = [None] * nranks
            dist.all_gather_object(all_halos, halos, group=pg)
        else:
            all_halos = [halos]
        self.plan_host = halo_plan(self.ls, all_halos)

        li = self.ls.local_idx
        self.global_id = o[li]  # original particle index (sphexa's `id` field) of every local particle
        nl = li.size
        hd = sim.HydroData(nl, self.ls.first, self.ls.last, box_lim, boundary, params, device=device)
        f = glob["fields"]
        loc = {k: sorted_field(v, np.float64 if k == "temp" else np.float32)[li] for k, v in f.items()}
        hd.set_fields(x=xs[li], y=ys[li], z=zs[li], **loc)
        lt = host.build_tree(xs[li], ys[li], zs[li], box_lim, boundary, bucket_focus)
        assert np.array_equal(lt.order, np.arange(nl, dtype=np.uint32)), "local particles must already be SFC-sorted"
        hd.set_tree(lt)
        self.hd = hd
        self.n_assigned = self.ls.last - self.ls.first
        self.n_global = n

        # device side of the plan + NCCL communicator
        dev = torch.device(device)
        ph = self.plan_host
        self._send_idx = torch.from_numpy(ph.send_idx.view(np.int32)).to(dev)
        nsend = int(ph.send_offsets[-1])
        buf_bytes = self.MAX_EXCHANGE_ARRAYS * ((nsend * 4 + 15) // 16 * 16) + 64
        self._send_buf = torch.empty(buf_bytes, dtype=torch.uint8, device=dev)
        self.plan = _cabi.SphxHaloPlan()
        self.plan.numPeers = ph.peers.size
        self.plan.peers = ph.peers.ctypes.data
        self.plan.sendOffsets = ph.send_offsets.ctypes.data
        self.plan.sendIdx = self._send_idx.data_ptr()
        self.plan.recvBegin = ph.recv_begin.ctypes.data
        self.plan.recvCount = ph.recv_count.ctypes.data
        self.plan.sendBuffer = self._send_buf.data_ptr()
        self.plan.sendBufferBytes = buf_bytes

        uid = np.zeros(_cabi.UNIQUE_ID_BYTES, np.uint8)
        if rank == 0:
            _cabi.check(self.L.sphx_comm_unique_id(_p(uid)))
        if nranks > 1:
            box_ = [uid.tobytes()]
            dist.broadcast_object_list(box_, src=0, group=pg)
            uid = np.frombuffer(box_[0], np.uint8).copy()
        self.comm = C.c_void_p()
        torch.cuda.set_device(dev)
        _cabi.check(self.L.sphx_comm_init(C.byref(self.comm), rank, nranks, _p(uid)))
        self.result = _cabi.SphxStepResult()

    def step(self):
        a = self.hd.args()
        _cabi.check(self.L.sphx_hydro_step_dist(C.byref(a), self.comm, C.byref(self.plan), C.byref(self.result)))
        return self.result

    def exchange(self, names):
        """one halo exchange of the listed fields (Domain::exchangeHalos call shape)"""
        arrs = (C.c_void_p * len(names))(*[self.hd.f[k].data_ptr() for k in names])
        eb = (C.c_int * len(names))(*[self.hd.f[k].element_size() for k in names])
        st = self.hd.stream.cuda_stream if self.hd.stream is not None else None
        _cabi.check(self.L.sphx_halo_exchange(self.comm, C.byref(self.plan), len(names), arrs, eb, st))

    def assigned(self, name) -> np.ndarray:
        return self.hd.get(name)[self.ls.first:self.ls.last]

    def assigned_ids(self) -> np.ndarray:
        return self.global_id[self.ls.first:self.ls.last]

    def close(self):
        if self.comm:
            self.L.sphx_comm_free(self.comm)
            self.comm = C.c_void_p()


# ---------------------------------------------------------------------------------------------------------------------
# dynamic decomposition (multi-rank Domain::sync): cell plan from the global histogram
# ---------------------------------------------------------------------------------------------------------------------
@dataclass
class CellPlan:
    """what sphx_cell_plan_build_host derives for one rank from the global per-cell particle counts"""
    level: int
    cell_splits: np.ndarray   # uint64, nranks + 1
    peers: np.ndarray         # int32
    send_offsets: np.ndarray  # uint32, numPeers + 1
    send_idx: np.ndarray      # uint32 local indices (valid in the layout [halos | assigned | halos])
    recv_begin: np.ndarray    # uint32
    recv_count: np.ndarray    # uint32
    recv_cells: np.ndarray    # uint32, sorted halo cells
    n_assigned: int
    n_halo_left: int
    n_halo_right: int
    n_global: int

    @property
    def n_local(self):
        return self.n_halo_left + self.n_assigned + self.n_halo_right


def cell_plan(global_counts: np.ndarray, level: int, boundary, rank: int, nranks: int) -> CellPlan:
    L = _cabi.load()
    g = np.ascontiguousarray(global_counts, np.uint32)
    assert g.size == 8 ** level
    per = np.array([int(b == 1) for b in boundary], np.int32)
    h = L.sphx_cell_plan_build_host(_p(g), level, _p(per), rank, nranks)
    if not h:
        raise ValueError("sphx_cell_plan_build_host: bad arguments")
    try:
        sz = np.zeros(8, np.uint64)
        L.sphx_cell_plan_sizes(h, _p(sz))
        npeer, nsend, ncells = int(sz[0]), int(sz[1]), int(sz[2])
        out = CellPlan(level, np.zeros(nranks + 1, np.uint64), np.zeros(npeer, np.int32),
                       np.zeros(npeer + 1, np.uint32), np.zeros(nsend, np.uint32), np.zeros(npeer, np.uint32),
                       np.zeros(npeer, np.uint32), np.zeros(ncells, np.uint32), int(sz[3]), int(sz[4]), int(sz[5]),
                       int(sz[6]))
        L.sphx_cell_plan_get(h, _p(out.cell_splits), _p(out.peers), _p(out.send_offsets), _p(out.send_idx),
                             _p(out.recv_begin), _p(out.recv_count), _p(out.recv_cells))
    finally:
        L.sphx_cell_plan_free(h)
    return out


def cell_level(box_lim, h_max: float, max_level: int = 7) -> int:
    """finest level whose cell edge is >= 2 h_max in every dimension (so the 26-neighbourhood covers every search
    sphere), capped at max_level (8^7 = 2 M cells: the histogram all-reduce and the host sweep stay cheap)"""
    ext = min(box_lim[1] - box_lim[0], box_lim[3] - box_lim[2], box_lim[5] - box_lim[4])
    lvl = 0
    while lvl < max_level and ext / (1 << (lvl + 1)) >= 2.0 * h_max * 1.0001:
        lvl += 1
    return lvl
