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


class _DomainTree:
    """OctreeNsView arrays owned by a SphxDomain (valid until its next sync)"""

    def __init__(self, view: _cabi.SphxTreeView):
        self._view = _cabi.SphxTreeView()
        C.memmove(C.byref(self._view), C.byref(view), C.sizeof(view))
        self.num_nodes, self.num_leaves = view.numNodes, view.numLeafNodes

    def view(self) -> _cabi.SphxTreeView:
        return self._view


class DistributedSimulation:
    """One rank of the multi-GPU time-step loop: the reference's main loop (sphexa.cpp:141-170) with a dynamic SFC
    domain decomposition redone in every sync().

    sync() = sphx_domain_sync_dist (csrc/domain_dist.cu): multi-rank Domain::sync (domain/domain.hpp:181-234) re-designed
    around ONE global object, the particle count per Hilbert cell of a level whose cell edge is >= 2 max(h):
      keys + local radix sort -> cell histogram -> ncclAllReduce -> decomposition plan on the device (assignment, halo
      cells, send lists, layout; no request messages) -> particle migration as one slice per peer and field -> merge of
      the arrivals -> halo exchange of x, y, z, h, m -> octree over the local particles.
    The reference negotiates the same things through a focus octree with peer-to-peer messages. This class only owns the
    torch tensors (fields and their spares) and swaps them as the C call says; nothing on the data path runs in Python.
    `pg`: torch.distributed group for host-side plumbing only (the NCCL unique id)."""

    SYNC_FIELDS = ("x", "y", "z", "h", "m", "vx", "vy", "vz", "x_m1", "y_m1", "z_m1", "du_m1", "temp", "alpha", "id")
    HALO_SYNC_FIELDS = ("x", "y", "z", "h", "m")

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

        init = {k: chunk(k, np.float64 if k in ("x", "y", "z", "temp") else np.float32)
                for k in self.SYNC_FIELDS if k != "id"}
        if "vx" in f:  # x_m1 = v * minDt (noh_init.hpp:99-101); zero for Sedov
            for k, v in (("x_m1", "vx"), ("y_m1", "vy"), ("z_m1", "vz")):
                init[k] = (init[v].double() * self.p.minDt).float()
        init["id"] = torch.arange(b, e, dtype=torch.int64, device=self.dev)
        self.comm = init_comm(self.L, rank, nranks, device, pg) if nranks > 1 else C.c_void_p()
        self.domain = C.c_void_p()
        box = host.make_box(self.box_lim, self.boundary)
        _cabi.check(self.L.sphx_domain_create(C.byref(self.domain), self.comm if nranks > 1 else None, C.byref(box),
                                              bucket))
        self.hd = None
        self._allocate(int(1.3 * (e - b)) + 4096, keep=(init, 0, e - b))
        self._in = (0, e - b)  # where this rank's particles sit in the field arrays
        self.plan = None
        self.result = _cabi.SphxStepResult()
        self.conserved = _cabi.SphxConserved()
        self.cons_scratch = torch.zeros(self.L.sphx_conserved_scratch_bytes(), dtype=torch.uint8, device=self.dev)
        self.iteration = 0
        self.level = 0
        self.turbulence = None  # set to a sim.Turbulence for the turbulence-ve propagator

    # -- memory -------------------------------------------------------------------------------------------------------
    def _allocate(self, capacity: int, keep=None):
        """field arrays + spares of `capacity` particles (and the step workspace for as many assigned particles); `keep`
        = (arrays, first, last): particles to carry over into [0, last - first) of the new arrays"""
        import torch
        hd = self.sim.HydroData(capacity, 0, capacity, self.box_lim, self.boundary, self.p, device=self.dev)
        for name in ("x_m1", "y_m1", "z_m1", "du_m1"):
            hd.f[name] = torch.zeros(capacity, dtype=torch.float32, device=self.dev)
        hd.f["id"] = torch.zeros(capacity, dtype=torch.int64, device=self.dev)
        if keep is not None:
            src, b, e = keep
            for k in self.SYNC_FIELDS:
                hd.f[k][:e - b].copy_(src[k][b:e])
        if self.hd is not None:
            hd.n, hd.first, hd.last, hd.tree = self.hd.n, self.hd.first, self.hd.last, self.hd.tree
        self.hd = hd
        self.spare = {k: torch.empty_like(hd.f[k]) for k in self.SYNC_FIELDS}
        self.capacity = capacity

    # -- Domain::sync -------------------------------------------------------------------------------------------------
    def sync(self):
        L, names = self.L, self.SYNC_FIELDS
        k = len(names)
        halo = (C.c_int * len(self.HALO_SYNC_FIELDS))(*[names.index(h) for h in self.HALO_SYNC_FIELDS])
        res = _cabi.SphxDomainResult()
        while True:
            f = self.hd.f
            a = _cabi.SphxDomainSyncArgs()
            a.count = k
            arrays = (C.c_void_p * k)(*[f[m].data_ptr() for m in names])
            spare = (C.c_void_p * k)(*[self.spare[m].data_ptr() for m in names])
            eb = (C.c_int * k)(*[f[m].element_size() for m in names])
            a.arrays, a.spare, a.elemBytes = C.cast(arrays, C.c_void_p), C.cast(spare, C.c_void_p), C.cast(eb, C.c_void_p)
            a.capacity, a.inFirst, a.inLast = self.capacity, self._in[0], self._in[1]
            a.numHaloFields, a.haloFields = len(self.HALO_SYNC_FIELDS), C.cast(halo, C.c_void_p)
            a.stream = None
            rc = L.sphx_domain_sync_dist(self.domain, C.byref(a), C.byref(res))
            if rc == 4:  # SPHX_ERR_WORKSPACE on every rank: nothing has moved, grow and call again
                self._allocate(int(1.25 * max(res.needCapacity, self.capacity)) + 4096,
                               keep=(f, self._in[0], self._in[1]))
                self._in = (0, self._in[1] - self._in[0])
                continue
            _cabi.check(rc)
            break
        if res.swapped:
            for m in names:
                self.hd.f[m], self.spare[m] = self.spare[m], self.hd.f[m]
        hd = self.hd
        hd.n, hd.first, hd.last = int(res.numLocal), int(res.first), int(res.last)
        self.box_lim = [float(v) for v in res.box.lim]
        hd.box_lim = list(self.box_lim)
        hd.tree = _DomainTree(res.tree)
        self._in = (hd.first, hd.last)
        self.level = int(res.level)
        self.plan = L.sphx_domain_halo_plan(self.domain)
        self._local_keys_ptr = res.localKeys

    @property
    def local_keys(self):
        """Hilbert keys of the local particles (copied out of the domain's buffer)"""
        import torch
        out = torch.empty(self.hd.n, dtype=torch.int64, device=self.dev)
        _cabi.check(self.L.sphx_domain_copy_local_keys(self.domain, out.data_ptr(), None))
        torch.cuda.synchronize(self.dev)
        return out

    def exchange(self, names):
        arrs = (C.c_void_p * len(names))(*[self.hd.f[k].data_ptr() for k in names])
        eb = (C.c_int * len(names))(*[self.hd.f[k].element_size() for k in names])
        _cabi.check(self.L.sphx_domain_exchange_halos(self.domain, len(names), arrs, eb, None))

    # -- the rest of the loop ------------------------------------------------------------------------------------------
    def compute_forces(self):
        """HydroVeProp::computeForces; with `turbulence` set (a sim.Turbulence), TurbVeProp::computeForces: every rank
        holds the same stirring state (same seed, same time steps) and stirs its assigned particles"""
        a = self.hd.args()
        if self.nranks > 1:
            _cabi.check(self.L.sphx_hydro_step_dist(C.byref(a), self.comm, self.plan, C.byref(self.result)))
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
        if self.domain:
            self.L.sphx_domain_destroy(self.domain)
            self.domain = C.c_void_p()
        if self.comm:
            self.L.sphx_comm_free(self.comm)
            self.comm = C.c_void_p()
