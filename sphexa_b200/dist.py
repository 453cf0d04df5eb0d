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
import math
import time
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


def cell_plan(global_counts: np.ndarray, level: int, boundary, rank: int, nranks: int, rings=None) -> CellPlan:
    """host builder (test oracle of the device plan). rings: per-cell reach in rings of cells (uint8), None = 1"""
    L = _cabi.load()
    g = np.ascontiguousarray(global_counts, np.uint32)
    assert g.size == 8 ** level
    per = np.array([int(b == 1) for b in boundary], np.int32)
    rg = None if rings is None else np.ascontiguousarray(rings, np.uint8)
    assert rg is None or rg.size == g.size
    h = L.sphx_cell_plan_build_host_rings(_p(g), _p(rg) if rg is not None else None, level, _p(per), rank, nranks)
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


def cell_plan_device(global_counts, local_counts, level: int, boundary, rank: int, nranks: int, scratch=None,
                     send_capacity: int | None = None, want_recv_cells: bool = False, rings=None, max_ring: int = 1):
    """sphx_cell_plan_build_device: the plan of `rank` from device histograms (torch int32 tensors of 8^level counts).
    Returns (CellPlan with host-side peer arrays, send_idx device tensor)."""
    import torch

    L = _cabi.load()
    dev = global_counts.device
    ncell = 8 ** level
    assert global_counts.numel() == ncell and local_counts.numel() == ncell
    need = L.sphx_cell_plan_device_bytes(level)
    if scratch is None or scratch.numel() < need:
        scratch = torch.empty(need, dtype=torch.uint8, device=dev)
    cap = int(send_capacity) if send_capacity is not None else 2 * int(local_counts.sum().item()) + 65536
    send_idx = torch.empty(max(cap, 1), dtype=torch.int32, device=dev)
    recv_cells = torch.empty(ncell, dtype=torch.int32, device=dev) if want_recv_cells else None
    per = np.array([int(b == 1) for b in boundary], np.int32)
    out = _cabi.SphxCellPlanSummary()
    assert rings is None or (rings.numel() == ncell and rings.dtype == torch.uint8)
    _cabi.check(L.sphx_cell_plan_build_device(global_counts.data_ptr(), local_counts.data_ptr(),
                                              rings.data_ptr() if rings is not None else None, max_ring, level,
                                              _p(per), rank, nranks, scratch.data_ptr(), scratch.numel(),
                                              send_idx.data_ptr(), cap,
                                              recv_cells.data_ptr() if recv_cells is not None else None,
                                              C.byref(out), None))
    recv = np.array(out.recvCount[:nranks], np.int64)
    send = np.array(out.sendCount[:nranks], np.int64)
    peers = [r for r in range(nranks) if r != rank and (recv[r] > 0 or send[r] > 0)]
    n_left, n_asg = int(out.nHaloLeft), int(out.nAssigned)
    begin, pos = {}, 0
    for r in range(nranks):
        if r == rank:
            pos = n_left + n_asg  # right halos follow the assigned range
            continue
        begin[r] = pos
        pos += int(recv[r])
    cp = CellPlan(level, np.array(out.cellSplits[:nranks + 1], np.uint64), np.array(peers, np.int32),
                  np.concatenate([[0], np.cumsum(send[peers])]).astype(np.uint32), np.zeros(0, np.uint32),
                  np.array([begin[r] for r in peers], np.uint32), np.array([recv[r] for r in peers], np.uint32),
                  (recv_cells[:int(out.numRecvCells)].cpu().numpy().view(np.uint32) if want_recv_cells
                   else np.zeros(0, np.uint32)),
                  n_asg, n_left, int(out.nHaloRight), int(out.nGlobal))
    cp.send_off_local = np.array(out.sendOffLocal[:nranks + 1], np.int64)
    cp.num_send = int(out.numSend)
    return cp, send_idx, scratch


def cell_level(box_lim, h_max: float, max_level: int = 7) -> int:
    """finest level whose cell edge is >= 2 h_max in every dimension (so the 26-neighbourhood covers every search
    sphere), capped at max_level (8^7 = 2 M cells: the histogram all-reduce and the host sweep stay cheap)"""
    ext = min(box_lim[1] - box_lim[0], box_lim[3] - box_lim[2], box_lim[5] - box_lim[4])
    lvl = 0
    while lvl < max_level and ext / (1 << (lvl + 1)) >= 2.0 * h_max * 1.0001:
        lvl += 1
    return lvl


#: the plan works on cells kRingLevels levels finer than the coarsest admissible ones and gives every cell its own reach
#: in rings of cells: thin halos where the smoothing lengths are small, wide ones only where they are large
RING_LEVELS = 2


def init_comm(L, rank: int, nranks: int, device, pg=None):
    """NCCL communicator of libsphx: rank 0 creates the unique id, the host process group distributes it"""
    import torch
    import torch.distributed as dist

    uid = np.zeros(_cabi.UNIQUE_ID_BYTES, np.uint8)
    if rank == 0:
        _cabi.check(L.sphx_comm_unique_id(_p(uid)))
    if nranks > 1:
        box_ = [uid.tobytes()]
        dist.broadcast_object_list(box_, src=0, group=pg)
        uid = np.frombuffer(box_[0], np.uint8).copy()
    comm = C.c_void_p()
    torch.cuda.set_device(torch.device(device))
    _cabi.check(L.sphx_comm_init(C.byref(comm), rank, nranks, _p(uid)))
    return comm


class DistributedSimulation:
    """One rank of the multi-GPU time-step loop: the reference's main loop (sphexa.cpp:141-170) with a dynamic SFC
    domain decomposition redone in every sync().

    sync() = multi-rank Domain::sync (domain/domain.hpp:181-234) re-designed around ONE global object, the particle
    count per Hilbert cell of a level whose cell edge is >= 2 max(h):
      keys + local radix sort (device) -> cell histogram (device) -> ncclAllReduce -> host plan (assignment, halo
      cells, send lists, layout: sphx_cell_plan_build_host, no request messages) -> particle migration as one slice
      per peer and field (sphx_exchange_slices) -> merge sort of the arrivals -> halo exchange of x, y, z, h, m ->
      octree over the local particles (sphx_domain_sync, presorted).
    The reference negotiates the same things through a focus octree with peer-to-peer messages.
    `pg`: torch.distributed group for host-side plumbing only (unique id, the R x R matrix of migration counts)."""

    SYNC_FIELDS = ("x", "y", "z", "h", "m", "vx", "vy", "vz", "x_m1", "y_m1", "z_m1", "du_m1", "temp", "alpha", "id")
    HALO_SYNC_FIELDS = ("x", "y", "z", "h", "m")
    MAX_EXCHANGE_ARRAYS = 7

    def __init__(self, sim, glob: dict, rank: int, nranks: int, device, pg=None, bucket: int = 64):
        import torch

        self.sim, self.L = sim, _cabi.load()
        self.rank, self.nranks, self.pg, self.bucket = rank, nranks, pg, bucket
        self.dev = torch.device(device)
        torch.cuda.set_device(self.dev)
        self.box_lim = [float(v) for v in glob["box"]]
        self.boundary = [int(b) for b in glob["boundary"]]
        self.p = glob["params"]
        # initial distribution: a contiguous slice of the particles in GENERATION order; the first sync migrates them.
        # `glob` holds either all particles or (glob["slice"] = (b, e), glob["n_global"]) just this rank's slice.
        if "slice" in glob:
            n, (b, e) = int(glob["n_global"]), glob["slice"]
            assert (b, e) == (rank * n // nranks, (rank + 1) * n // nranks) and glob["x"].size == e - b
            lo = b
        else:
            n = glob["x"].size
            b, e = rank * n // nranks, (rank + 1) * n // nranks
            lo = 0
        self.n_global = n
        f = glob["fields"]
        n_have = glob["x"].size

        def chunk(name, dtype):
            v = f.get(name, 0.0) if name not in ("x", "y", "z") else glob[name]
            a = v[b - lo:e - lo] if isinstance(v, np.ndarray) and v.shape == (n_have,) else np.full(e - b, v)
            return torch.from_numpy(np.ascontiguousarray(a, dtype)).to(self.dev)

        self.cur = {k: chunk(k, np.float64 if k in ("x", "y", "z", "temp") else np.float32)
                    for k in self.SYNC_FIELDS if k != "id"}
        if "vx" in f:  # x_m1 = v * minDt (noh_init.hpp:99-101); zero for Sedov
            for k, v in (("x_m1", "vx"), ("y_m1", "vy"), ("z_m1", "vz")):
                self.cur[k] = (self.cur[v].double() * self.p.minDt).float()
        self.cur["id"] = torch.arange(b, e, dtype=torch.int64, device=self.dev)
        self.comm = init_comm(self.L, rank, nranks, device, pg)
        self.hd = None
        self.plan = None
        self.result = _cabi.SphxStepResult()
        self.conserved = _cabi.SphxConserved()
        self.cons_scratch = torch.zeros(self.L.sphx_conserved_scratch_bytes(), dtype=torch.uint8, device=self.dev)
        self.iteration = 0
        self.timings = {}
        self._plan_scratch = None
        self.turbulence = None  # set to a sim.Turbulence for the turbulence-ve propagator
        self.profile = None  # set to a dict to accumulate per-stage wall times of sync() (synchronises the device)

    # -- helpers ------------------------------------------------------------------------------------------------------
    def _allreduce_host(self, values, op):
        a = np.ascontiguousarray(values, np.float64)
        if self.nranks > 1:
            _cabi.check(self.L.sphx_allreduce_f64(self.comm, _p(a), a.size, op, None))
        return a

    def _sfc_sort(self, x, y, z, presorted=False, tree=None):
        """keys (+ SFC permutation, + tree) of n particles through sphx_domain_sync"""
        import torch
        n = x.numel()
        keys = torch.empty(n, dtype=torch.int64, device=self.dev)
        order = torch.empty(n, dtype=torch.int32, device=self.dev)
        max_nodes = tree.max_nodes if tree is not None else 0
        scratch = torch.empty(self.L.sphx_domain_sync_bytes(n, max_nodes), dtype=torch.uint8, device=self.dev)
        a = _cabi.SphxSyncArgs()
        a.n, a.box, a.bucketSize = n, host.make_box(self.box_lim, self.boundary), self.bucket
        a.x, a.y, a.z = x.data_ptr(), y.data_ptr(), z.data_ptr()
        a.keys, a.order, a.maxNodes = keys.data_ptr(), order.data_ptr(), max_nodes
        if tree is not None:
            for k in ("prefixes", "childOffsets", "internalToLeaf", "levelRange", "leaves", "layout", "centers",
                      "sizes"):
                setattr(a, k, getattr(tree, k).data_ptr())
        a.scratch, a.scratchBytes = scratch.data_ptr(), scratch.numel()
        a.flags = (1 if presorted else 0) | (2 if tree is None else 0)
        nn, nl = C.c_int(0), C.c_int(0)
        rc = self.L.sphx_domain_sync(C.byref(a), None, C.byref(nn), C.byref(nl))
        if tree is not None and rc == 0:
            tree.num_nodes, tree.num_leaves = nn.value, nl.value
        return rc, keys, order

    def _reorder(self, order, src: dict, dst: dict, n: int):
        names = list(src)
        k = len(names)
        _cabi.check(self.L.sphx_reorder_fields(order.data_ptr(), n, k,
                                               (C.c_void_p * k)(*[src[m].data_ptr() for m in names]),
                                               (C.c_void_p * k)(*[dst[m].data_ptr() for m in names]),
                                               (C.c_int * k)(*[src[m].element_size() for m in names]), None))

    # -- Domain::sync -------------------------------------------------------------------------------------------------
    def sync(self):
        import torch

        L, R, me = self.L, self.nranks, self.rank
        cur = self.cur
        n_old = cur["x"].numel()
        prof = self.profile

        def tick(name):
            # stage timing (profile runs only: synchronises the device)
            if prof is not None:
                torch.cuda.synchronize()
                now = time.perf_counter()
                prof[name] = prof.get(name, 0.0) + now - self._t0
                self._t0 = now

        if prof is not None:
            torch.cuda.synchronize()
            self._t0 = time.perf_counter()
        # global box: open dimensions follow the particles (makeGlobalBox, box_mpi.hpp:66-109)
        if any(b != 1 for b in self.boundary):
            lo = [float(cur[k].min()) if n_old else np.inf for k in "xyz"]
            hi = [float(cur[k].max()) if n_old else -np.inf for k in "xyz"]
            lo, hi = self._allreduce_host(lo, 0), self._allreduce_host(hi, 1)
            for d in range(3):
                if self.boundary[d] != 1:
                    self.box_lim[2 * d], self.box_lim[2 * d + 1] = float(lo[d]), float(hi[d])
        h_max = float(self._allreduce_host([float(cur["h"].max()) if n_old else 0.0], 1)[0])
        h_sum = self._allreduce_host([float(cur["h"].sum()) if n_old else 0.0, float(n_old)], 2)
        h_mean = h_sum[0] / max(h_sum[1], 1.0)
        # cells RING_LEVELS finer than the coarsest level whose edge is >= 2 max(h); a cell reaches ceil(2 h_cell / edge)
        # rings of cells, h_cell = largest h in the cell over all ranks. With (nearly) uniform smoothing lengths the
        # finer cells buy nothing (same reach, 64 x more cells to scan): one ring of the coarse cells then.
        coarse = cell_level(self.box_lim, h_max)
        level = min(7, coarse + (RING_LEVELS if h_max > 1.25 * h_mean else 0))
        ncell = 8 ** level
        edge = min(self.box_lim[2 * d + 1] - self.box_lim[2 * d] for d in range(3)) / (1 << level)
        max_ring = max(1, min(16, int(math.ceil(2.0 * h_max * 1.0001 / edge))))
        tick("box_hmax")

        # 1. local SFC order, cell histogram, global histogram
        rc, keys, order = self._sfc_sort(cur["x"], cur["y"], cur["z"])
        _cabi.check(rc)
        tick("keys_sort")
        hist = torch.zeros(ncell, dtype=torch.int32, device=self.dev)
        _cabi.check(L.sphx_cell_histogram(keys.data_ptr(), n_old, level, hist.data_ptr(), None))
        local_hist = hist
        if R > 1:
            local_hist = hist.clone()
            _cabi.check(L.sphx_allreduce_device(self.comm, hist.data_ptr(), ncell, 0, 2, None))
        # largest h per cell (this rank, then all ranks) -> reach of the cell in rings
        hcell = torch.zeros(ncell, dtype=torch.float32, device=self.dev)
        if n_old:
            cell_of = (keys >> (3 * (21 - level))).to(torch.int64)
            hcell.scatter_reduce_(0, cell_of, cur["h"][order.to(torch.int64)], reduce="amax", include_self=True)
        if R > 1:
            _cabi.check(L.sphx_allreduce_device(self.comm, hcell.data_ptr(), ncell, 2, 1, None))
        rings = torch.ceil(hcell.double() * (2.0 * 1.0001 / edge)).clamp_(1, max_ring).to(torch.uint8)
        tick("histogram_allreduce")

        # 2. plan on the device: assignment, halo cells, send lists, layout; only the summary POD comes to the host
        cp, send_idx, self._plan_scratch = cell_plan_device(hist, local_hist, level, self.boundary, me, R,
                                                            scratch=self._plan_scratch,
                                                            send_capacity=4 * n_old + 65536, rings=rings,
                                                            max_ring=max_ring)
        tick("device_plan")
        send_off = cp.send_off_local  # my sorted particles [send_off[r], send_off[r+1]) belong to rank r
        send_cnt = np.diff(send_off)
        if R > 1:
            # R x R matrix of migration counts: every rank fills its row, one all-reduce completes it
            m = torch.zeros(R * R, dtype=torch.int64, device=self.dev)
            m[me * R:(me + 1) * R] = torch.from_numpy(send_cnt.astype(np.int64)).to(self.dev)
            _cabi.check(L.sphx_allreduce_device(self.comm, m.data_ptr(), R * R, 1, 2, None))
            recv_cnt = m.view(R, R)[:, me].cpu().numpy().astype(np.int64)
        else:
            recv_cnt = send_cnt.copy()
        recv_off = np.concatenate([[0], np.cumsum(recv_cnt)])
        n_new = int(recv_off[-1])
        assert n_new == cp.n_assigned, (n_new, cp.n_assigned)
        tick("counts_allgather")

        # 3. migration: sort my particles, ship one slice per peer and field, merge-sort what arrived
        sorted_ = {k: torch.empty_like(cur[k]) for k in self.SYNC_FIELDS}
        self._reorder(order, cur, sorted_, n_old)
        if R > 1:
            arrived = {k: torch.empty(n_new, dtype=cur[k].dtype, device=self.dev) for k in self.SYNC_FIELDS}
            so = np.ascontiguousarray(send_off, np.uint64)
            ro = np.ascontiguousarray(recv_off, np.uint64)
            names = list(self.SYNC_FIELDS)
            k = len(names)
            _cabi.check(L.sphx_exchange_slices(self.comm, _p(so), _p(ro), k,
                                               (C.c_void_p * k)(*[sorted_[m].data_ptr() for m in names]),
                                               (C.c_void_p * k)(*[arrived[m].data_ptr() for m in names]),
                                               (C.c_int * k)(*[sorted_[m].element_size() for m in names]), None))
            rc, _, order2 = self._sfc_sort(arrived["x"], arrived["y"], arrived["z"])
            _cabi.check(rc)
        else:
            arrived, order2 = sorted_, None
        tick("migrate_sort")

        # 4. local arrays [halos | assigned | halos]
        n_local, first, last = cp.n_local, cp.n_halo_left, cp.n_halo_left + cp.n_assigned
        if self.hd is None or n_local > self.cap_local or cp.n_assigned > self.cap_assigned:
            self.cap_local, self.cap_assigned = int(1.2 * n_local) + 1024, int(1.2 * cp.n_assigned) + 1024
            hd = self.sim.HydroData(self.cap_local, 0, self.cap_assigned, self.box_lim, self.boundary, self.p,
                                    device=self.dev)
            for name in ("x_m1", "y_m1", "z_m1", "du_m1"):
                hd.f[name] = torch.zeros(self.cap_local, dtype=torch.float32, device=self.dev)
            hd.f["id"] = torch.zeros(self.cap_local, dtype=torch.int64, device=self.dev)
            hd.tree = self.sim.DeviceTree.empty(max(4096, self.cap_local // 4), self.dev)
            self.hd = hd
        hd = self.hd
        hd.n, hd.first, hd.last = n_local, first, last
        hd.box_lim = list(self.box_lim)
        dstv = {k: hd.f[k][first:last] for k in self.SYNC_FIELDS}
        if order2 is not None:
            self._reorder(order2, arrived, dstv, n_new)
        else:
            for k in self.SYNC_FIELDS:
                dstv[k].copy_(arrived[k])
        self.cur = dstv  # views into the local arrays: integrate() updates them in place
        tick("layout_reorder")

        # 5. halo plan on the device, halo exchange of the fields the search and the first loop read
        self._send_idx = send_idx
        nsend = cp.num_send
        buf_bytes = self.MAX_EXCHANGE_ARRAYS * ((nsend * 8 + 15) // 16 * 16) + 64
        self._send_buf = torch.empty(buf_bytes, dtype=torch.uint8, device=self.dev)
        self._plan_host = cp  # keeps the host arrays alive
        pl = _cabi.SphxHaloPlan()
        pl.numPeers = cp.peers.size
        pl.peers, pl.sendOffsets = cp.peers.ctypes.data, cp.send_offsets.ctypes.data
        pl.sendIdx = self._send_idx.data_ptr()
        pl.recvBegin, pl.recvCount = cp.recv_begin.ctypes.data, cp.recv_count.ctypes.data
        pl.sendBuffer, pl.sendBufferBytes = self._send_buf.data_ptr(), buf_bytes
        self.plan = pl
        if R > 1:
            self.exchange(list(self.HALO_SYNC_FIELDS))
        tick("halo_exchange")

        # 6. octree over the local particles (already in SFC order)
        while True:
            rc, lkeys, _ = self._sfc_sort(hd.f["x"][:n_local], hd.f["y"][:n_local], hd.f["z"][:n_local], presorted=True,
                                          tree=hd.tree)
            if rc == 4 and hd.tree.max_nodes < 8 * n_local + 64:
                hd.tree = self.sim.DeviceTree.empty(2 * hd.tree.max_nodes, self.dev)
                continue
            _cabi.check(rc)
            break
        self.local_keys = lkeys
        self.level = level
        tick("local_tree")

    def exchange(self, names):
        arrs = (C.c_void_p * len(names))(*[self.hd.f[k].data_ptr() for k in names])
        eb = (C.c_int * len(names))(*[self.hd.f[k].element_size() for k in names])
        _cabi.check(self.L.sphx_halo_exchange(self.comm, C.byref(self.plan), len(names), arrs, eb, None))

    # -- the rest of the loop ------------------------------------------------------------------------------------------
    def compute_forces(self):
        """HydroVeProp::computeForces; with `turbulence` set (a sim.Turbulence), TurbVeProp::computeForces: every rank
        holds the same stirring state (same seed, same time steps) and stirs its assigned particles"""
        a = self.hd.args()
        if self.nranks > 1:
            _cabi.check(self.L.sphx_hydro_step_dist(C.byref(a), self.comm, C.byref(self.plan), C.byref(self.result)))
        else:
            _cabi.check(self.L.sphx_hydro_step(C.byref(a), None, None, C.byref(self.result)))
        if self.turbulence is not None:
            self.turbulence.drive(self.hd, self.p.minDt)
        return self.result

    def compute_conserved(self):
        f, hd = self.hd.f, self.hd
        _cabi.check(self.L.sphx_conserved_quantities(
            f["x"].data_ptr(), f["y"].data_ptr(), f["z"].data_ptr(), f["vx"].data_ptr(), f["vy"].data_ptr(),
            f["vz"].data_ptr(), f["m"].data_ptr(), f["temp"].data_ptr(), None, f["nc"].data_ptr(), hd.first, hd.last,
            self.p.gamma, self.p.muiConst, 0.0, self.cons_scratch.data_ptr(),
            self.comm if self.nranks > 1 else None, None, C.byref(self.conserved)))
        return self.conserved

    def integrate(self):
        hd = self.hd
        dt, dt1, tt = C.c_double(self.p.minDt), C.c_double(self.p.minDt_m1), C.c_double(self.p.ttot)
        _cabi.check(self.L.sphx_compute_timestep(self.result.minDtCourant, self.result.minDtRho, self.p.maxDtIncrease,
                                                 C.byref(dt), C.byref(dt1), C.byref(tt),
                                                 self.comm if self.nranks > 1 else None, None))
        self.p.minDt, self.p.minDt_m1, self.p.ttot = dt.value, dt1.value, tt.value
        a = _cabi.SphxIntegrateArgs()
        for k in _cabi.INTEGRATE_FIELDS:
            setattr(a, k, hd.f[k].data_ptr() if k in hd.f else None)
        a.first, a.last, a.box = hd.first, hd.last, host.make_box(self.box_lim, self.boundary)
        a.dt, a.dt_m1, a.gamma, a.muiConst, a.ng0 = self.p.minDt, self.p.minDt_m1, self.p.gamma, self.p.muiConst, self.p.ng0
        _cabi.check(self.L.sphx_integrate(C.byref(a)))

    def step(self):
        self.sync()
        self.compute_forces()
        c = self.compute_conserved()
        row = (self.iteration, self.p.ttot, self.p.minDt, c.etot, c.ecin, c.eint, c.linmom, c.angmom, c.totalNeighbors)
        self.integrate()
        self.iteration += 1
        return row

    def assigned(self, name) -> np.ndarray:
        a = self.hd.f[name][self.hd.first:self.hd.last].cpu().numpy()
        return a.view(np.uint32) if name == "nc" else a

    def close(self):
        if self.comm:
            self.L.sphx_comm_free(self.comm)
            self.comm = C.c_void_p()
