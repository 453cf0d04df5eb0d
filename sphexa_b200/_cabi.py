"""ctypes mirror of include/sphx.h. Loads the in-tree libsphx.so; there is no fallback implementation."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "libsphx.so"


class SphxBox(C.Structure):
    _fields_ = [("lim", C.c_double * 6), ("boundary", C.c_int * 3)]


class SphxTreeView(C.Structure):
    _fields_ = [("numLeafNodes", C.c_int), ("numNodes", C.c_int), ("prefixes", C.c_void_p),
                ("childOffsets", C.c_void_p), ("internalToLeaf", C.c_void_p), ("levelRange", C.c_void_p),
                ("leaves", C.c_void_p), ("layout", C.c_void_p), ("centers", C.c_void_p), ("sizes", C.c_void_p),
                ("searchExtFactor", C.c_float)]


FIELD_NAMES = ("x y z h m vx vy vz temp u nc xm kx gradh prho c rho p c11 c12 c13 c22 c23 c33 divv curlv alpha "
               "ax ay az du dV11 dV12 dV13 dV22 dV23 dV33").split()


class SphxFields(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in FIELD_NAMES]


class SphxParams(C.Structure):
    _fields_ = [("K", C.c_double), ("Kcour", C.c_double), ("Krho", C.c_double), ("gamma", C.c_double),
                ("minDt", C.c_double), ("polytropic_const", C.c_double), ("polytropic_index", C.c_double),
                ("muiConst", C.c_float), ("soundSpeedConst", C.c_float), ("alphamin", C.c_float),
                ("alphamax", C.c_float), ("decay_constant", C.c_float), ("Atmin", C.c_float), ("Atmax", C.c_float),
                ("ramp", C.c_float), ("ng0", C.c_uint), ("ngmax", C.c_uint), ("eosChoice", C.c_int),
                ("avClean", C.c_int)]


class SphxStepArgs(C.Structure):
    _fields_ = [("f", SphxFields), ("numLocal", C.c_size_t), ("first", C.c_size_t), ("last", C.c_size_t),
                ("p", SphxParams), ("box", SphxBox), ("tree", SphxTreeView), ("wh", C.c_void_p), ("whd", C.c_void_p),
                ("workspace", C.c_void_p), ("workspaceBytes", C.c_size_t), ("stream", C.c_void_p)]


class SphxFieldsF64(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in FIELD_NAMES]


class SphxParamsF64(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("K", "Kcour", "Krho", "gamma", "minDt", "polytropic_const", "polytropic_index",
                                          "muiConst", "soundSpeedConst", "alphamin", "alphamax", "decay_constant", "Atmin",
                                          "Atmax", "ramp")] + \
               [("ng0", C.c_uint), ("ngmax", C.c_uint), ("eosChoice", C.c_int), ("avClean", C.c_int)]


class SphxStepArgsF64(C.Structure):
    _fields_ = [("f", SphxFieldsF64), ("numLocal", C.c_size_t), ("first", C.c_size_t), ("last", C.c_size_t),
                ("p", SphxParamsF64), ("box", SphxBox), ("tree", SphxTreeView), ("wh", C.c_void_p), ("whd", C.c_void_p),
                ("workspace", C.c_void_p), ("workspaceBytes", C.c_size_t), ("stream", C.c_void_p)]


class SphxStepResult(C.Structure):
    _fields_ = [("minDtCourant", C.c_double), ("minDtRho", C.c_double), ("totalNeighbors", C.c_ulong),
                ("maxNc", C.c_uint), ("numHIterated", C.c_uint)]


class SphxHaloPlan(C.Structure):
    _fields_ = [("numPeers", C.c_int), ("peers", C.c_void_p), ("sendOffsets", C.c_void_p), ("sendIdx", C.c_void_p),
                ("recvBegin", C.c_void_p), ("recvCount", C.c_void_p), ("sendBuffer", C.c_void_p),
                ("sendBufferBytes", C.c_size_t)]


class SphxDomainSyncArgs(C.Structure):
    _fields_ = [("count", C.c_int), ("arrays", C.c_void_p), ("spare", C.c_void_p), ("elemBytes", C.c_void_p),
                ("capacity", C.c_size_t), ("inFirst", C.c_size_t), ("inLast", C.c_size_t), ("numHaloFields", C.c_int),
                ("haloFields", C.c_void_p), ("stream", C.c_void_p)]


class SphxDomainResult(C.Structure):
    _fields_ = [("first", C.c_size_t), ("last", C.c_size_t), ("numLocal", C.c_size_t), ("numGlobal", C.c_size_t),
                ("needCapacity", C.c_size_t), ("box", SphxBox), ("tree", SphxTreeView), ("localKeys", C.c_void_p),
                ("level", C.c_int), ("swapped", C.c_int)]


class SphxSyncArgs(C.Structure):
    _fields_ = [("n", C.c_size_t), ("box", SphxBox), ("bucketSize", C.c_uint), ("x", C.c_void_p), ("y", C.c_void_p),
                ("z", C.c_void_p), ("keys", C.c_void_p), ("order", C.c_void_p), ("maxNodes", C.c_int),
                ("prefixes", C.c_void_p), ("childOffsets", C.c_void_p), ("internalToLeaf", C.c_void_p),
                ("levelRange", C.c_void_p), ("leaves", C.c_void_p), ("layout", C.c_void_p), ("centers", C.c_void_p),
                ("sizes", C.c_void_p), ("scratch", C.c_void_p), ("scratchBytes", C.c_size_t), ("stream", C.c_void_p),
                ("flags", C.c_int)]


INTEGRATE_FIELDS = "x y z x_m1 y_m1 z_m1 vx vy vz ax ay az temp u du du_m1 h nc".split()


class SphxIntegrateArgs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in INTEGRATE_FIELDS] + [
        ("first", C.c_size_t), ("last", C.c_size_t), ("box", SphxBox), ("dt", C.c_double), ("dt_m1", C.c_double),
        ("gamma", C.c_double), ("muiConst", C.c_float), ("ng0", C.c_uint), ("stream", C.c_void_p)]


class SphxConserved(C.Structure):
    _fields_ = [("ecin", C.c_double), ("eint", C.c_double), ("egrav", C.c_double), ("etot", C.c_double),
                ("linmom", C.c_double), ("angmom", C.c_double), ("linmom3", C.c_double * 3),
                ("angmom3", C.c_double * 3), ("totalNeighbors", C.c_ulong)]


MAX_RANKS = 64


class SphxCellPlanSummary(C.Structure):
    _fields_ = [("cellSplits", C.c_uint64 * (MAX_RANKS + 1)), ("sendOffLocal", C.c_uint64 * (MAX_RANKS + 1)),
                ("nGlobal", C.c_uint64), ("nAssigned", C.c_uint64), ("nHaloLeft", C.c_uint64),
                ("nHaloRight", C.c_uint64), ("recvCount", C.c_uint32 * MAX_RANKS), ("sendCount", C.c_uint32 * MAX_RANKS),
                ("numRecvCells", C.c_uint32), ("numSend", C.c_uint32), ("overflow", C.c_uint32), ("pad", C.c_uint32)]


class SphxTurbulenceSettings(C.Structure):
    """defaults = sphexa::TurbulenceConstants() (main/src/init/turbulence_init.hpp:47-72)"""
    _fields_ = [("solWeight", C.c_double), ("Lbox", C.c_double), ("stEnergyPrefac", C.c_double),
                ("stMachVelocity", C.c_double), ("epsilon", C.c_double), ("powerLawExp", C.c_double),
                ("anglesExp", C.c_double), ("stMaxModes", C.c_uint64), ("rngSeed", C.c_uint64),
                ("stSpectForm", C.c_int)]

    def __init__(self, **kw):
        d = dict(solWeight=0.5, Lbox=1.0, stEnergyPrefac=5.0e-3, stMachVelocity=0.3, epsilon=1e-15,
                 powerLawExp=5.0 / 3, anglesExp=2.0, stMaxModes=100000, rngSeed=251299, stSpectForm=1)
        d.update(kw)
        super().__init__(**d)


UNIQUE_ID_BYTES = 128
SPHX_SYNC_PRESORTED, SPHX_SYNC_NO_TREE, SPHX_SYNC_LIMIT_SHRINK = 1, 2, 4

HALO_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int))

STATUS = {0: "SPHX_OK", 1: "SPHX_ERR_NO_DEVICE", 2: "SPHX_ERR_CUDA", 3: "SPHX_ERR_INVALID", 4: "SPHX_ERR_WORKSPACE",
          5: "SPHX_ERR_H_CONVERGENCE", 6: "SPHX_ERR_NGMAX_OVERFLOW", 7: "SPHX_ERR_TRAVERSAL", 8: "SPHX_ERR_NCCL", 9: "SPHX_ERR_TABLE"}

# every symbol include/sphx.h declares
EXPORTS = ["sphx_last_error", "sphx_abi_version", "sphx_debug_candidate_chunk", "sphx_device_check", "sphx_workspace_bytes", "sphx_workspace_layout",
           "sphx_make_tables_host", "sphx_invalidate_tables", "sphx_table_mode", "sphx_find_neighbors_xmass", "sphx_find_neighbors_sph", "sphx_xmass", "sphx_ve_def_gradh", "sphx_eos",
           "sphx_iad_divv_curlv", "sphx_av_switches", "sphx_momentum_energy", "sphx_hydro_step",
           "sphx_find_neighbors", "sphx_export_neighbors", "sphx_host_tree_build", "sphx_host_tree_free",
           "sphx_host_tree_sizes", "sphx_host_tree_get", "sphx_hilbert_keys_host", "sphx_update_h_host",
           "sphx_powf_host", "sphx_sfc_assignment_host", "sphx_find_halos_host", "sphx_comm_unique_id",
           "sphx_comm_init", "sphx_comm_free", "sphx_halo_exchange", "sphx_allreduce_f64", "sphx_hydro_step_dist",
           "sphx_allreduce_device", "sphx_exchange_slices", "sphx_reduce_step_result", "sphx_comm_rank", "sphx_domain_create", "sphx_domain_destroy", "sphx_domain_sync_dist", "sphx_domain_exchange_halos", "sphx_domain_halo_plan", "sphx_domain_copy_local_keys", "sphx_workspace_bytes_f64", "sphx_make_tables_host_f64", "sphx_find_neighbors_sph_f64", "sphx_xmass_f64", "sphx_ve_def_gradh_f64", "sphx_eos_f64", "sphx_iad_divv_curlv_f64", "sphx_av_switches_f64", "sphx_momentum_energy_f64", "sphx_hydro_step_f64", "sphx_export_neighbors_f64", "sphx_cell_plan_build_host", "sphx_cell_plan_free", "sphx_cell_plan_sizes", "sphx_cell_plan_get", "sphx_cell_plan_build_host_rings",
           "sphx_cell_plan_device_bytes", "sphx_cell_plan_build_device",
           "sphx_domain_sync_bytes", "sphx_domain_sync", "sphx_cell_histogram", "sphx_reorder_fields", "sphx_compute_timestep",
           "sphx_compute_positions", "sphx_update_smoothing_length", "sphx_integrate", "sphx_conserved_scratch_bytes",
           "sphx_conserved_quantities", "sphx_turbulence_create", "sphx_turbulence_free", "sphx_turbulence_sizes",
           "sphx_turbulence_get", "sphx_turbulence_restore", "sphx_turbulence_advance_host", "sphx_drive_turbulence", "sphx_compute_stirring"]


class SphxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{STATUS.get(code, code)}: {msg}")
        self.code = code


_lib = None


def load():
    """Load libsphx.so. Raises if the library was not built: the product has no Python/CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(make -C sphexa_b200/csrc). There is no fallback implementation.")
    L = C.CDLL(str(LIB_PATH))
    L.sphx_last_error.restype = C.c_char_p
    L.sphx_workspace_bytes.restype = C.c_size_t
    L.sphx_workspace_bytes.argtypes = [C.c_size_t, C.c_uint]
    L.sphx_invalidate_tables.restype = None
    L.sphx_invalidate_tables.argtypes = []
    L.sphx_table_mode.restype = C.c_int
    L.sphx_table_mode.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.sphx_workspace_layout.restype = None
    L.sphx_workspace_layout.argtypes = [C.c_size_t, C.c_uint, C.c_void_p]
    L.sphx_host_tree_build.restype = C.c_void_p
    L.sphx_host_tree_build.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint]
    L.sphx_host_tree_free.argtypes = [C.c_void_p]
    L.sphx_host_tree_sizes.argtypes = [C.c_void_p, C.c_void_p]
    L.sphx_host_tree_get.argtypes = [C.c_void_p] * 11
    L.sphx_hilbert_keys_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    L.sphx_update_h_host.restype = C.c_float
    L.sphx_update_h_host.argtypes = [C.c_uint, C.c_uint, C.c_float]
    L.sphx_powf_host.restype = C.c_float
    L.sphx_powf_host.argtypes = [C.c_float, C.c_float]
    L.sphx_make_tables_host.argtypes = [C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
    L.sphx_find_neighbors.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t,
                                      C.c_void_p, C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p]
    for name in ("sphx_find_neighbors_xmass", "sphx_find_neighbors_sph", "sphx_iad_divv_curlv", "sphx_momentum_energy"):
        getattr(L, name).argtypes = [C.c_void_p, C.c_void_p]
    for name in ("sphx_ve_def_gradh", "sphx_eos", "sphx_av_switches", "sphx_xmass"):
        getattr(L, name).argtypes = [C.c_void_p]
    L.sphx_export_neighbors.argtypes = [C.c_void_p, C.c_void_p]
    L.sphx_hydro_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.sphx_sfc_assignment_host.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_uint, C.c_void_p]
    L.sphx_find_halos_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint,
                                       C.c_size_t, C.c_size_t, C.c_void_p]
    L.sphx_comm_unique_id.argtypes = [C.c_void_p]
    L.sphx_comm_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.sphx_comm_free.argtypes = [C.c_void_p]
    L.sphx_halo_exchange.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.sphx_allreduce_f64.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.sphx_hydro_step_dist.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.sphx_reduce_step_result.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.sphx_workspace_bytes_f64.restype = C.c_size_t
    L.sphx_workspace_bytes_f64.argtypes = [C.c_size_t, C.c_uint]
    L.sphx_make_tables_host_f64.argtypes = [C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
    for name in ("sphx_find_neighbors_sph_f64", "sphx_iad_divv_curlv_f64", "sphx_momentum_energy_f64",
                 "sphx_hydro_step_f64", "sphx_export_neighbors_f64"):
        getattr(L, name).argtypes = [C.c_void_p, C.c_void_p]
    for name in ("sphx_xmass_f64", "sphx_ve_def_gradh_f64", "sphx_eos_f64", "sphx_av_switches_f64"):
        getattr(L, name).argtypes = [C.c_void_p]
    L.sphx_comm_rank.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.sphx_domain_create.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint]
    L.sphx_domain_destroy.argtypes = [C.c_void_p]
    L.sphx_domain_destroy.restype = None
    L.sphx_domain_sync_dist.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.sphx_domain_exchange_halos.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.sphx_domain_halo_plan.argtypes = [C.c_void_p]
    L.sphx_domain_halo_plan.restype = C.c_void_p
    L.sphx_domain_copy_local_keys.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.sphx_domain_sync_bytes.restype = C.c_size_t
    L.sphx_domain_sync_bytes.argtypes = [C.c_size_t, C.c_int]
    L.sphx_domain_sync.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.sphx_allreduce_device.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p]
    L.sphx_exchange_slices.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p]
    L.sphx_cell_plan_build_host.restype = C.c_void_p
    L.sphx_cell_plan_build_host.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]
    L.sphx_cell_plan_free.argtypes = [C.c_void_p]
    L.sphx_cell_plan_sizes.argtypes = [C.c_void_p, C.c_void_p]
    L.sphx_cell_plan_get.argtypes = [C.c_void_p] * 8
    L.sphx_debug_candidate_chunk.restype = None
    L.sphx_debug_candidate_chunk.argtypes = [C.c_uint]
    L.sphx_cell_plan_device_bytes.restype = C.c_size_t
    L.sphx_cell_plan_device_bytes.argtypes = [C.c_int]
    L.sphx_cell_plan_build_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                              C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p,
                                              C.c_void_p, C.c_void_p]
    L.sphx_cell_plan_build_host_rings.restype = C.c_void_p
    L.sphx_cell_plan_build_host_rings.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]
    L.sphx_cell_histogram.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p]
    L.sphx_reorder_fields.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.sphx_compute_timestep.argtypes = [C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p]
    for name in ("sphx_compute_positions", "sphx_update_smoothing_length", "sphx_integrate"):
        getattr(L, name).argtypes = [C.c_void_p]
    L.sphx_conserved_scratch_bytes.restype = C.c_size_t
    L.sphx_conserved_quantities.argtypes = [C.c_void_p] * 10 + [C.c_size_t, C.c_size_t, C.c_double, C.c_float,
                                                                 C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                                                                 C.c_void_p]
    L.sphx_turbulence_create.argtypes = [C.c_void_p, C.c_void_p]
    L.sphx_turbulence_free.argtypes = [C.c_void_p]
    L.sphx_turbulence_free.restype = None
    L.sphx_turbulence_sizes.argtypes = [C.c_void_p, C.c_void_p]
    L.sphx_turbulence_sizes.restype = None
    L.sphx_turbulence_get.argtypes = [C.c_void_p] * 8
    L.sphx_turbulence_get.restype = None
    L.sphx_turbulence_restore.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_char_p]
    L.sphx_turbulence_advance_host.argtypes = [C.c_void_p, C.c_double]
    L.sphx_drive_turbulence.argtypes = [C.c_void_p] * 7 + [C.c_size_t, C.c_size_t, C.c_double, C.c_void_p]
    L.sphx_compute_stirring.argtypes = [C.c_void_p] * 7 + [C.c_size_t, C.c_size_t, C.c_void_p]
    _lib = L
    return L


def check(rc: int):
    if rc != 0:
        raise SphxError(rc, load().sphx_last_error().decode())
