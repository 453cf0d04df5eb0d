"""Python plumbing around the C ABI: torch tensors as device memory for the ParticlesData fields, one call per loop.

Mirrors the field names of sphexa::ParticlesData (sph/include/sph/particles_data.hpp:220-248) and the call order of
HydroVeProp::computeForces (main/src/propagator/ve_hydro.hpp:130-204). Nothing here computes: every number comes
out of libsphx.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _cabi, host

F64_FIELDS = ("x", "y", "z", "temp", "du")
U32_FIELDS = ("nc",)
STEP_OUTPUTS = ("h", "nc", "xm", "kx", "gradh", "prho", "c", "c11", "c12", "c13", "c22", "c23", "c33", "divv", "curlv",
                "alpha", "ax", "ay", "az", "du")


@dataclass
class Params:
    """scalar attributes of ParticlesData with the reference defaults (particles_data.hpp:88-146)"""
    K: float = 0.0
    Kcour: float = 0.2
    Krho: float = 0.06
    gamma: float = 5.0 / 3.0
    minDt: float = 1e-12
    minDt_m1: float = 1e-12
    polytropic_const: float = 1.0
    polytropic_index: float = 5.0 / 3.0
    muiConst: float = 10.0
    soundSpeedConst: float = 1.0
    alphamin: float = 0.05
    alphamax: float = 1.0
    decay_constant: float = 0.2
    Atmin: float = 0.1
    Atmax: float = 0.2
    ramp: float = 10.0
    ng0: int = 100
    ngmax: int = 150
    eosChoice: int = 0
    avClean: int = 0
    maxDtIncrease: float = 1.1
    ttot: float = 0.0

    @staticmethod
    def from_dump(d: dict) -> "Params":
        v = d["params"]
        return Params(K=v[0], Kcour=v[1], Krho=v[2], gamma=v[3], muiConst=v[4], minDt=v[5], minDt_m1=v[6],
                      alphamin=v[7], alphamax=v[8], decay_constant=v[9], Atmin=v[10], Atmax=v[11], ramp=v[12],
                      ttot=v[13], eosChoice=int(v[14]), maxDtIncrease=v[15], ng0=int(d["ng0"][0]),
                      ngmax=int(d["ngmax"][0]), avClean=int("dV11" in d))

    def to_c(self) -> _cabi.SphxParams:
        p = _cabi.SphxParams()
        for name, _ in _cabi.SphxParams._fields_:
            setattr(p, name, getattr(self, name))
        return p


class DeviceTree:
    """OctreeNsView arrays on the device"""

    def __init__(self, t: host.HostTree | dict, device):
        if isinstance(t, host.HostTree):
            t = t.as_dump_dict()
        self.num_leaves = int(t["numLeafNodes"][0])
        self.num_nodes = int(t["numNodes"][0])
        up = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a).view(dt)).to(device)  # noqa: E731
        self.prefixes = up(t["tree_prefixes"], np.int64)
        self.childOffsets = up(t["tree_childOffsets"], np.int32)
        self.internalToLeaf = up(t["tree_internalToLeaf"], np.int32)
        self.levelRange = up(t["tree_levelRange"], np.int32)
        self.leaves = up(t["tree_leaves"], np.int64)
        self.layout = up(t["tree_layout"], np.int32)
        self.centers = up(t["tree_centers"], np.float64)
        self.sizes = up(t["tree_sizes"], np.float64)

    @classmethod
    def empty(cls, max_nodes: int, device) -> "DeviceTree":
        """uninitialised buffers for sphx_domain_sync (capacity max_nodes)"""
        t = cls.__new__(cls)
        t.num_leaves = t.num_nodes = 0
        t.max_nodes = int(max_nodes)
        mk = lambda n, dt: torch.zeros(n, dtype=dt, device=device)  # noqa: E731
        t.prefixes = mk(max_nodes, torch.int64)
        t.childOffsets = mk(max_nodes, torch.int32)
        t.internalToLeaf = mk(max_nodes, torch.int32)
        t.levelRange = mk(23, torch.int32)
        t.leaves = mk(max_nodes + 1, torch.int64)
        t.layout = mk(max_nodes + 1, torch.int32)
        t.centers = mk(3 * max_nodes, torch.float64)
        t.sizes = mk(3 * max_nodes, torch.float64)
        return t

    def to_host(self) -> dict:
        """the arrays in the reference-harness dump naming (tests), trimmed to the node counts"""
        nn, nl = self.num_nodes, self.num_leaves
        cpu = lambda t, k, dt: t[:k].cpu().numpy().view(dt)  # noqa: E731
        return dict(tree_prefixes=cpu(self.prefixes, nn, np.uint64), tree_childOffsets=cpu(self.childOffsets, nn, np.int32),
                    tree_internalToLeaf=cpu(self.internalToLeaf, nn, np.int32),
                    tree_levelRange=cpu(self.levelRange, 23, np.int32), tree_leaves=cpu(self.leaves, nl + 1, np.uint64),
                    tree_layout=cpu(self.layout, nl + 1, np.uint32), tree_centers=cpu(self.centers, 3 * nn, np.float64),
                    tree_sizes=cpu(self.sizes, 3 * nn, np.float64))

    def view(self) -> _cabi.SphxTreeView:
        v = _cabi.SphxTreeView()
        v.numLeafNodes, v.numNodes = self.num_leaves, self.num_nodes
        for k in ("prefixes", "childOffsets", "internalToLeaf", "levelRange", "leaves", "layout", "centers", "sizes"):
            setattr(v, k, getattr(self, k).data_ptr())
        v.searchExtFactor = 1.0
        return v


class HydroData:
    """Device-resident particle fields + tree + tables + workspace for one rank."""

    def __init__(self, n_local: int, first: int, last: int, box_lim, boundary, params: Params, device="cuda:0",
                 stream: torch.cuda.Stream | None = None):
        self.L = _cabi.load()
        _cabi.check(self.L.sphx_device_check())
        self.device = torch.device(device)
        self.n, self.first, self.last = n_local, first, last
        self.box_lim, self.boundary = [float(v) for v in box_lim], [int(b) for b in boundary]
        self.p = params
        self.f: dict[str, torch.Tensor] = {}
        for name in _cabi.FIELD_NAMES:
            if name in ("u", "rho", "p") or (name.startswith("dV") and not params.avClean):
                continue
            dt = torch.float64 if name in F64_FIELDS else (torch.int32 if name in U32_FIELDS else torch.float32)
            self.f[name] = torch.zeros(n_local, dtype=dt, device=self.device)
        wh, whd, K = host.make_tables()
        if self.p.K == 0.0:
            self.p.K = K
        self.wh = torch.from_numpy(wh).to(self.device)
        self.whd = torch.from_numpy(whd).to(self.device)
        self.tree: DeviceTree | None = None
        nbytes = self.L.sphx_workspace_bytes(last - first, self.p.ngmax)
        self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        self.stream = stream
        self.result = _cabi.SphxStepResult()

    # -- plumbing ---------------------------------------------------------------------------------------------------
    def set_fields(self, **arrays):
        for k, a in arrays.items():
            t = self.f[k]
            src = torch.from_numpy(np.ascontiguousarray(a).view(np.int32) if t.dtype == torch.int32 else
                                   np.ascontiguousarray(a))
            t.copy_(src.to(t.dtype), non_blocking=False)

    def get(self, name) -> np.ndarray:
        a = self.f[name].cpu().numpy()
        return a.view(np.uint32) if name in U32_FIELDS else a

    def set_tree(self, t):
        self.tree = t if isinstance(t, DeviceTree) else DeviceTree(t, self.device)

    def args(self) -> _cabi.SphxStepArgs:
        a = _cabi.SphxStepArgs()
        for name in _cabi.FIELD_NAMES:
            setattr(a.f, name, self.f[name].data_ptr() if name in self.f else None)
        a.numLocal, a.first, a.last = self.n, self.first, self.last
        a.p = self.p.to_c()
        a.box = host.make_box(self.box_lim, self.boundary)
        a.tree = self.tree.view()
        a.wh, a.whd = self.wh.data_ptr(), self.whd.data_ptr()
        a.workspace, a.workspaceBytes = self.workspace.data_ptr(), self.workspace.numel()
        a.stream = self.stream.cuda_stream if self.stream is not None else None
        return a

    # -- the six loops, one C-ABI call each --------------------------------------------------------------------------
    def find_neighbors_xmass(self, sync=True):
        a = self.args()
        _cabi.check(self.L.sphx_find_neighbors_xmass(C.byref(a), C.byref(self.result) if sync else None))

    def find_neighbors_sph(self, sync=True):
        a = self.args()
        _cabi.check(self.L.sphx_find_neighbors_sph(C.byref(a), C.byref(self.result) if sync else None))

    def xmass(self):
        a = self.args()
        _cabi.check(self.L.sphx_xmass(C.byref(a)))

    def ve_def_gradh(self):
        a = self.args()
        _cabi.check(self.L.sphx_ve_def_gradh(C.byref(a)))

    def eos(self):
        a = self.args()
        _cabi.check(self.L.sphx_eos(C.byref(a)))

    def iad_divv_curlv(self, sync=True):
        a = self.args()
        _cabi.check(self.L.sphx_iad_divv_curlv(C.byref(a), C.byref(self.result) if sync else None))

    def av_switches(self):
        a = self.args()
        _cabi.check(self.L.sphx_av_switches(C.byref(a)))

    def momentum_energy(self, sync=True):
        a = self.args()
        _cabi.check(self.L.sphx_momentum_energy(C.byref(a), C.byref(self.result) if sync else None))

    def hydro_step(self, halo=None, sync=True):
        """sphx_hydro_step: all loops in the reference order; `halo` is an optional Python callable
        (list[(ptr, elem_bytes)]) -> int used as the halo-exchange callback."""
        a = self.args()
        cb = None
        if halo is not None:
            def _tramp(user, count, arrays, elem_bytes):
                return int(halo([(arrays[i], elem_bytes[i]) for i in range(count)]) or 0)
            cb = _cabi.HALO_FN(_tramp)
        _cabi.check(self.L.sphx_hydro_step(C.byref(a), cb, None, C.byref(self.result) if sync else None))
        return self.result

    def block_stats(self) -> dict:
        """diagnostics: per-block candidate counts and flags read back from the workspace"""
        lay = np.zeros(8, np.uint64)
        self.L.sphx_workspace_layout(self.last - self.first, self.p.ngmax, lay.ctypes.data)
        nb = int(lay[5])
        raw = self.workspace[int(lay[1]): int(lay[1]) + nb * 40].cpu().numpy()
        desc = raw.view(np.dtype([("o", np.float64, 3), ("candBegin", np.uint32), ("numCand", np.uint32),
                                  ("flags", np.uint32), ("pad", np.uint32)]))
        scal = self.workspace[:64].cpu().numpy()
        pad = desc["pad"]
        return dict(numLeaves=pad & 0xfff, numTiles=(pad >> 12) & 0xff, hRepeats=(pad >> 20) & 0xf, precise=pad >> 31,
                    numCand=desc["numCand"].copy(), flags=desc["flags"].copy(), origin=desc["o"].copy(),
                    candBegin=desc["candBegin"].copy(), candTop=int(scal[28:32].view(np.uint32)[0]),
                    errFlags=int(scal[24:28].view(np.uint32)[0]), candCapacity=int(lay[7]))

    def export_neighbors_device(self) -> torch.Tensor:
        """the neighbour list in the reference CPU layout [(i - first) * ngmax + k], as an int32 device tensor"""
        out = torch.zeros((self.last - self.first) * self.p.ngmax, dtype=torch.int32, device=self.device)
        a = self.args()
        _cabi.check(self.L.sphx_export_neighbors(C.byref(a), C.c_void_p(out.data_ptr())))
        torch.cuda.synchronize(self.device)
        return out

    def export_neighbors(self) -> np.ndarray:
        return self.export_neighbors_device().cpu().numpy().view(np.uint32)


class HydroDataF64:
    """The all-double type set (every field fp64) behind sphx_*_f64: the precision path of libsphx (csrc/loops_f64.cu),
    one rank. Same call order as HydroData; fields are torch float64 tensors."""

    def __init__(self, n: int, box_lim, boundary, params: Params, device="cuda:0"):
        self.L = _cabi.load()
        _cabi.check(self.L.sphx_device_check())
        self.device = torch.device(device)
        self.n, self.first, self.last = n, 0, n
        self.box_lim, self.boundary = [float(v) for v in box_lim], [int(b) for b in boundary]
        self.p = params
        self.f: dict[str, torch.Tensor] = {}
        for name in _cabi.FIELD_NAMES:
            if name in ("u", "rho", "p") or (name.startswith("dV") and not params.avClean):
                continue
            self.f[name] = torch.zeros(n, dtype=torch.int32 if name == "nc" else torch.float64, device=self.device)
        wh, whd, K = np.zeros(20000), np.zeros(20000), C.c_double(0)
        _cabi.check(self.L.sphx_make_tables_host_f64(6.0, wh.ctypes.data, whd.ctypes.data, C.byref(K)))
        if self.p.K == 0.0:
            self.p.K = K.value
        self.wh, self.whd = torch.from_numpy(wh).to(self.device), torch.from_numpy(whd).to(self.device)
        self.tree: DeviceTree | None = None
        self.workspace = torch.empty(self.L.sphx_workspace_bytes_f64(n, self.p.ngmax), dtype=torch.uint8,
                                     device=self.device)
        self.result = _cabi.SphxStepResult()

    def set_fields(self, **arrays):
        for k, a in arrays.items():
            self.f[k].copy_(torch.from_numpy(np.ascontiguousarray(a, np.float64)))

    def get(self, name) -> np.ndarray:
        a = self.f[name].cpu().numpy()
        return a.view(np.uint32) if name == "nc" else a

    def set_tree(self, t):
        self.tree = t if isinstance(t, DeviceTree) else DeviceTree(t, self.device)

    def args(self) -> _cabi.SphxStepArgsF64:
        a = _cabi.SphxStepArgsF64()
        for name in _cabi.FIELD_NAMES:
            setattr(a.f, name, self.f[name].data_ptr() if name in self.f else None)
        a.numLocal, a.first, a.last = self.n, self.first, self.last
        for name, _ in _cabi.SphxParamsF64._fields_:
            setattr(a.p, name, getattr(self.p, name))
        a.box = host.make_box(self.box_lim, self.boundary)
        a.tree = self.tree.view()
        a.wh, a.whd = self.wh.data_ptr(), self.whd.data_ptr()
        a.workspace, a.workspaceBytes = self.workspace.data_ptr(), self.workspace.numel()
        return a

    def _call(self, name, with_result=False):
        a = self.args()
        fn = getattr(self.L, name)
        _cabi.check(fn(C.byref(a), C.byref(self.result)) if with_result else fn(C.byref(a)))

    def find_neighbors_sph(self):
        self._call("sphx_find_neighbors_sph_f64", True)

    def xmass(self):
        self._call("sphx_xmass_f64")

    def ve_def_gradh(self):
        self._call("sphx_ve_def_gradh_f64")

    def eos(self):
        self._call("sphx_eos_f64")

    def iad_divv_curlv(self):
        self._call("sphx_iad_divv_curlv_f64", True)

    def av_switches(self):
        self._call("sphx_av_switches_f64")

    def momentum_energy(self):
        self._call("sphx_momentum_energy_f64", True)

    def hydro_step(self):
        self._call("sphx_hydro_step_f64", True)
        return self.result

    def export_neighbors(self) -> np.ndarray:
        out = torch.zeros(self.n * self.p.ngmax, dtype=torch.int32, device=self.device)
        a = self.args()
        _cabi.check(self.L.sphx_export_neighbors_f64(C.byref(a), C.c_void_p(out.data_ptr())))
        torch.cuda.synchronize(self.device)
        return out.cpu().numpy().view(np.uint32)


class Turbulence:
    """sph::TurbulenceData + sph::driveTurbulence (sph/include/sph/hydro_turb/turbulence_data.hpp, driver.hpp:102-128)
    behind sphx_turbulence_* / sphx_drive_turbulence: host state (modes, OU phases, std::mt19937) and device tables
    live in libsphx."""

    def __init__(self, **settings):
        self.L = _cabi.load()
        s = _cabi.SphxTurbulenceSettings(**settings)
        self.handle = C.c_void_p()
        _cabi.check(self.L.sphx_turbulence_create(C.byref(s), C.byref(self.handle)))
        sz = (C.c_size_t * 4)()
        self.L.sphx_turbulence_sizes(self.handle, sz)
        self.num_modes, self.lattice, self.max_index, self._rng_bytes = int(sz[0]), bool(sz[1]), int(sz[2]), int(sz[3])

    def state(self) -> dict:
        nm = self.num_modes
        out = dict(modes=np.zeros(3 * nm), amplitudes=np.zeros(nm), phases=np.zeros(6 * nm), phasesReal=np.zeros(3 * nm),
                   phasesImag=np.zeros(3 * nm), scalars=np.zeros(4))
        sz = (C.c_size_t * 4)()
        self.L.sphx_turbulence_sizes(self.handle, sz)
        rng = C.create_string_buffer(int(sz[3]))
        self.L.sphx_turbulence_get(self.handle, *[out[k].ctypes.data for k in ("modes", "amplitudes", "phases",
                                                                                "phasesReal", "phasesImag", "scalars")],
                                   C.cast(rng, C.c_void_p))
        out["rng"] = rng.value.decode()
        return out

    def restore(self, phases=None, rng: str | None = None, modes=None, amplitudes=None, scalars=None):
        """TurbulenceData::loadOrStore: replace (parts of) the state"""
        arr = lambda a: None if a is None else np.ascontiguousarray(a, np.float64)  # noqa: E731
        ph, mo, am, sc = arr(phases), arr(modes), arr(amplitudes), arr(scalars)
        nm = am.size if am is not None else self.num_modes
        ptr = lambda a: None if a is None else a.ctypes.data  # noqa: E731
        _cabi.check(self.L.sphx_turbulence_restore(self.handle, nm, ptr(mo), ptr(am), ptr(ph), ptr(sc),
                                                   rng.encode() if rng is not None else None))
        sz = (C.c_size_t * 4)()
        self.L.sphx_turbulence_sizes(self.handle, sz)
        self.num_modes, self.lattice, self.max_index = int(sz[0]), bool(sz[1]), int(sz[2])

    def advance_host(self, min_dt: float):
        """updateNoise + computePhases (host half of driveTurbulence)"""
        _cabi.check(self.L.sphx_turbulence_advance_host(self.handle, min_dt))

    def drive(self, hd: "HydroData", min_dt: float):
        f = hd.f
        _cabi.check(self.L.sphx_drive_turbulence(self.handle, f["x"].data_ptr(), f["y"].data_ptr(), f["z"].data_ptr(),
                                                 f["ax"].data_ptr(), f["ay"].data_ptr(), f["az"].data_ptr(), hd.first,
                                                 hd.last, min_dt, hd.stream.cuda_stream if hd.stream is not None else None))

    def compute_stirring(self, hd: "HydroData"):
        f = hd.f
        _cabi.check(self.L.sphx_compute_stirring(self.handle, f["x"].data_ptr(), f["y"].data_ptr(), f["z"].data_ptr(),
                                                 f["ax"].data_ptr(), f["ay"].data_ptr(), f["az"].data_ptr(), hd.first,
                                                 hd.last, hd.stream.cuda_stream if hd.stream is not None else None))

    def close(self):
        if self.handle:
            self.L.sphx_turbulence_free(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


class Simulation(HydroData):
    """One rank of the reference's time-step loop (main/src/sphexa/sphexa.cpp:141-170 with HydroVeProp,
    main/src/propagator/ve_hydro.hpp:114-215), every stage a libsphx call on device-resident fields:
    sync (sphx_domain_sync + sphx_reorder_fields) -> computeForces (sphx_hydro_step) -> conserved quantities ->
    integrate (sphx_compute_timestep + sphx_integrate)."""

    #: fields Domain::sync keeps ordered (ve_hydro.hpp:72-76 conserved fields + coordinates, h, m and the particle id)
    SYNC_FIELDS = ("x", "y", "z", "h", "m", "vx", "vy", "vz", "x_m1", "y_m1", "z_m1", "du_m1", "temp", "alpha", "id")

    def __init__(self, n: int, box_lim, boundary, params: Params, device="cuda:0", bucket_size: int = 64):
        super().__init__(n, 0, n, box_lim, boundary, params, device=device)
        dev = self.device
        for name in ("x_m1", "y_m1", "z_m1", "du_m1"):
            self.f[name] = torch.zeros(n, dtype=torch.float32, device=dev)
        self.f["id"] = torch.arange(n, dtype=torch.int64, device=dev)
        self.spare = {k: torch.empty_like(self.f[k]) for k in self.SYNC_FIELDS}
        self.bucket_size = bucket_size
        self._synced_once = False
        self.keys = torch.zeros(n, dtype=torch.int64, device=dev)
        self.order = torch.zeros(n, dtype=torch.int32, device=dev)
        self._alloc_tree(max(4096, n // 4))
        self.cons_scratch = torch.zeros(self.L.sphx_conserved_scratch_bytes(), dtype=torch.uint8, device=dev)
        self.conserved = _cabi.SphxConserved()
        self.iteration = 0
        #: set to a Turbulence to run the reference's turbulence-ve propagator (TurbVeProp, turb_ve.hpp:67-72)
        self.turbulence: Turbulence | None = None

    def _alloc_tree(self, max_nodes: int):
        self.tree = DeviceTree.empty(max_nodes, self.device)
        nbytes = self.L.sphx_domain_sync_bytes(self.n, max_nodes)
        self.sync_scratch = torch.empty(nbytes, dtype=torch.uint8, device=self.device)

    def sync(self):
        """Domain::sync for one rank: keys, SFC order, octree (sphx_domain_sync), then the field reorder."""
        while True:
            t = self.tree
            a = _cabi.SphxSyncArgs()
            a.n, a.box, a.bucketSize = self.n, host.make_box(self.box_lim, self.boundary), self.bucket_size
            a.x, a.y, a.z = (self.f[k].data_ptr() for k in "xyz")
            a.keys, a.order, a.maxNodes = self.keys.data_ptr(), self.order.data_ptr(), t.max_nodes
            for k in ("prefixes", "childOffsets", "internalToLeaf", "levelRange", "leaves", "layout", "centers", "sizes"):
                setattr(a, k, getattr(t, k).data_ptr())
            a.scratch, a.scratchBytes = self.sync_scratch.data_ptr(), self.sync_scratch.numel()
            a.stream = self.stream.cuda_stream if self.stream is not None else None
            # every sync but the first limits how fast an open box may shrink (domain/assignment.hpp:80-82)
            a.flags = _cabi.SPHX_SYNC_LIMIT_SHRINK if self._synced_once else 0
            nn, nl, box = C.c_int(0), C.c_int(0), _cabi.SphxBox()
            rc = self.L.sphx_domain_sync(C.byref(a), C.byref(box), C.byref(nn), C.byref(nl))
            if rc == 4 and t.max_nodes < 8 * self.n + 64:  # SPHX_ERR_WORKSPACE: grow the tree buffers
                self._alloc_tree(2 * t.max_nodes)
                continue
            _cabi.check(rc)
            t.num_nodes, t.num_leaves = nn.value, nl.value
            self.box_lim = [float(v) for v in box.lim]  # open dimensions follow the particles (makeGlobalBox)
            self._synced_once = True
            break
        names = self.SYNC_FIELDS
        src = (C.c_void_p * len(names))(*[self.f[k].data_ptr() for k in names])
        dst = (C.c_void_p * len(names))(*[self.spare[k].data_ptr() for k in names])
        eb = (C.c_int * len(names))(*[self.f[k].element_size() for k in names])
        _cabi.check(self.L.sphx_reorder_fields(self.order.data_ptr(), self.n, len(names), src, dst, eb, a.stream))
        for k in names:
            self.f[k], self.spare[k] = self.spare[k], self.f[k]

    def compute_forces(self):
        """HydroVeProp::computeForces after sync (ve_hydro.hpp:130-204); with `turbulence` set, TurbVeProp::computeForces
        (turb_ve.hpp:67-72): the same followed by driveTurbulence"""
        r = self.hydro_step()
        if self.turbulence is not None:
            self.turbulence.drive(self, self.p.minDt)
        return r

    def compute_conserved(self) -> _cabi.SphxConserved:
        f = self.f
        _cabi.check(self.L.sphx_conserved_quantities(
            f["x"].data_ptr(), f["y"].data_ptr(), f["z"].data_ptr(), f["vx"].data_ptr(), f["vy"].data_ptr(),
            f["vz"].data_ptr(), f["m"].data_ptr(), f["temp"].data_ptr(), None, f["nc"].data_ptr(), self.first, self.last,
            self.p.gamma, self.p.muiConst, 0.0, self.cons_scratch.data_ptr(), None,
            self.stream.cuda_stream if self.stream is not None else None, C.byref(self.conserved)))
        return self.conserved

    def integrate_args(self) -> _cabi.SphxIntegrateArgs:
        a = _cabi.SphxIntegrateArgs()
        for k in _cabi.INTEGRATE_FIELDS:
            setattr(a, k, self.f[k].data_ptr() if k in self.f else None)
        a.first, a.last, a.box = self.first, self.last, host.make_box(self.box_lim, self.boundary)
        a.dt, a.dt_m1, a.gamma, a.muiConst, a.ng0 = self.p.minDt, self.p.minDt_m1, self.p.gamma, self.p.muiConst, self.p.ng0
        a.stream = self.stream.cuda_stream if self.stream is not None else None
        return a

    def compute_timestep(self):
        dt, dt1, tt = C.c_double(self.p.minDt), C.c_double(self.p.minDt_m1), C.c_double(self.p.ttot)
        _cabi.check(self.L.sphx_compute_timestep(self.result.minDtCourant, self.result.minDtRho, self.p.maxDtIncrease,
                                                 C.byref(dt), C.byref(dt1), C.byref(tt), None, None))
        self.p.minDt, self.p.minDt_m1, self.p.ttot = dt.value, dt1.value, tt.value

    def integrate(self, fused=True):
        """HydroVeProp::integrate (ve_hydro.hpp:206-215)"""
        self.compute_timestep()
        a = self.integrate_args()
        if fused:
            _cabi.check(self.L.sphx_integrate(C.byref(a)))
        else:
            _cabi.check(self.L.sphx_compute_positions(C.byref(a)))
            _cabi.check(self.L.sphx_update_smoothing_length(C.byref(a)))

    def step(self):
        """one iteration of the main loop; returns the conserved quantities of the state the step started from"""
        self.sync()
        self.compute_forces()
        c = self.compute_conserved()
        row = (self.iteration, self.p.ttot, self.p.minDt, c.etot, c.ecin, c.eint, c.linmom, c.angmom, c.totalNeighbors)
        self.integrate()
        self.iteration += 1
        return row


def from_dump(d: dict, device="cuda:0") -> HydroData:
    """HydroData holding the inputs of one reference-harness dump (single rank: no halos)."""
    n = int(d["n"][0])
    hd = HydroData(n, 0, n, d["box"], d["boundary"], Params.from_dump(d), device=device)
    hd.set_fields(x=d["x"], y=d["y"], z=d["z"], h=d["h_in"], m=d["m"], vx=d["vx"], vy=d["vy"], vz=d["vz"],
                  temp=d["temp"], alpha=d["alpha_in"])
    hd.set_tree(d)
    return hd


def simulation_from_dump(d: dict, device="cuda:0") -> Simulation:
    """Simulation holding the state a reference-harness dump starts its step from (single rank)."""
    n = int(d["n"][0])
    s = Simulation(n, d["box"], d["boundary"], Params.from_dump(d), device=device)
    s.set_fields(x=d["x"], y=d["y"], z=d["z"], h=d["h_in"], m=d["m"], vx=d["vx"], vy=d["vy"], vz=d["vz"],
                 temp=d["temp"], alpha=d["alpha_in"], x_m1=d["x_m1"], y_m1=d["y_m1"], z_m1=d["z_m1"],
                 du_m1=d["du_m1"])
    s.f["id"].copy_(torch.from_numpy(d["id"].view(np.int64)))
    return s


def find_neighbors(x, y, z, h, tree: DeviceTree, box_lim, boundary, ngmax: int, first=0, last=None, device="cuda:0"):
    """cstone::findNeighbors call shape through sphx_find_neighbors: returns (neighbors[n*ngmax], counts[n])"""
    L = _cabi.load()
    dev = torch.device(device)
    xd, yd, zd = (torch.from_numpy(np.ascontiguousarray(a, np.float64)).to(dev) for a in (x, y, z))
    hd = torch.from_numpy(np.ascontiguousarray(h, np.float32)).to(dev)
    last = xd.numel() if last is None else last
    n = last - first
    nb = torch.zeros(max(n * ngmax, 1), dtype=torch.int32, device=dev)
    cnt = torch.zeros(max(n, 1), dtype=torch.int32, device=dev)
    box = host.make_box(box_lim, boundary)
    tv = tree.view()
    _cabi.check(L.sphx_find_neighbors(xd.data_ptr(), yd.data_ptr(), zd.data_ptr(), hd.data_ptr(), first, last,
                                      C.byref(box), C.byref(tv), ngmax, nb.data_ptr(), cnt.data_ptr(), None))
    return nb.cpu().numpy().view(np.uint32)[: n * ngmax], cnt.cpu().numpy().view(np.uint32)[:n]
