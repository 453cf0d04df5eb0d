"""ctypes binding of oracle/liboracle.so (the CPU restatement). TEST INFRASTRUCTURE ONLY.

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs only.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
LIB = REPO / "oracle" / "liboracle.so"


class OrcBox(C.Structure):
    _fields_ = [("lim", C.c_double * 6), ("boundary", C.c_int * 3)]


class OrcTree(C.Structure):
    _fields_ = [("numLeafNodes", C.c_int), ("numNodes", C.c_int), ("childOffsets", C.c_void_p),
                ("internalToLeaf", C.c_void_p), ("layout", C.c_void_p), ("centers", C.c_void_p),
                ("sizes", C.c_void_p), ("searchExtFactor", C.c_float)]


class OrcParams(C.Structure):
    _fields_ = [("K", C.c_double), ("Kcour", C.c_double), ("Krho", C.c_double), ("gamma", C.c_double),
                ("minDt", C.c_double), ("muiConst", C.c_float), ("alphamin", C.c_float), ("alphamax", C.c_float),
                ("decay_constant", C.c_float), ("Atmin", C.c_float), ("Atmax", C.c_float), ("ramp", C.c_float),
                ("ng0", C.c_uint), ("ngmax", C.c_uint)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not LIB.exists():
            subprocess.run(["make", "-C", str(REPO / "oracle"), "restate"], check=True)
        _lib = C.CDLL(str(LIB))
        _lib.orc_distance_sq.restype = C.c_double
        _lib.orc_distance_sq.argtypes = [C.c_int] + [C.c_double] * 6 + [C.c_void_p]
        _lib.orc_sphynx_3D_k.restype = C.c_double
        _lib.orc_sphynx_3D_k.argtypes = [C.c_double]
        _lib.orc_update_h_f.restype = C.c_float
        _lib.orc_update_h_f.argtypes = [C.c_uint, C.c_uint, C.c_float]
        _lib.orc_find_neighbors_sph_f.restype = C.c_ulong
        _lib.orc_xmass_jloop_d.restype = C.c_double
        _lib.orc_av_switches_jloop_d.restype = C.c_double
    return _lib


def P(a):
    """pointer to a C-contiguous numpy array (None -> NULL)"""
    if a is None:
        return C.c_void_p(0)
    assert a.flags["C_CONTIGUOUS"]
    return C.c_void_p(a.ctypes.data)


def make_box(lim, boundary) -> OrcBox:
    b = OrcBox()
    for i in range(6):
        b.lim[i] = float(lim[i])
    for i in range(3):
        b.boundary[i] = int(boundary[i])
    return b


def make_tree(d: dict, keep: list) -> OrcTree:
    """OrcTree view over the tree_* arrays of a reference dump; `keep` receives the arrays to keep them alive."""
    t = OrcTree()
    t.numLeafNodes = int(d["numLeafNodes"][0])
    t.numNodes = int(d["numNodes"][0])
    arrs = [np.ascontiguousarray(d[k]) for k in ("tree_childOffsets", "tree_internalToLeaf", "tree_layout",
                                                 "tree_centers", "tree_sizes")]
    keep.extend(arrs)
    t.childOffsets, t.internalToLeaf, t.layout, t.centers, t.sizes = [a.ctypes.data for a in arrs]
    t.searchExtFactor = 1.0
    return t


def make_params(d: dict) -> OrcParams:
    """params vector layout as dumped by ref_harness: K Kcour Krho gamma muiConst minDt minDt_m1 alphamin alphamax
    decay_constant Atmin Atmax ramp ttot eosChoice maxDtIncrease"""
    v = d["params"]
    p = OrcParams()
    p.K, p.Kcour, p.Krho, p.gamma, p.muiConst, p.minDt = v[0], v[1], v[2], v[3], v[4], v[5]
    p.alphamin, p.alphamax, p.decay_constant, p.Atmin, p.Atmax, p.ramp = v[7], v[8], v[9], v[10], v[11], v[12]
    p.ng0, p.ngmax = int(d["ng0"][0]), int(d["ngmax"][0])
    return p


def tables_f(sinc_index=6.0):
    wh = np.zeros(20000, np.float32)
    whd = np.zeros(20000, np.float32)
    K = C.c_double(0)
    lib().orc_tables_f(C.c_double(sinc_index), P(wh), P(whd), C.byref(K))
    return wh, whd, K.value


def tables_d(sinc_index=6.0):
    wh = np.zeros(20000, np.float64)
    whd = np.zeros(20000, np.float64)
    K = C.c_double(0)
    lib().orc_tables_d(C.c_double(sinc_index), P(wh), P(whd), C.byref(K))
    return wh, whd, K.value


def hydro_step_f(d: dict, wh=None, whd=None) -> dict:
    """Run the restated step (neighbour search + h-iteration + six loops) on the inputs of a reference dump.

    Returns a dict with the same output names as the dump (h, nc, neighbors, xm, kx, gradh, prho, c, c11.., divv,
    curlv, alpha, ax, ay, az, du, dts)."""
    L = lib()
    n = int(d["n"][0])
    first, last = 0, n
    keep = []
    box = make_box(d["box"], d["boundary"])
    tree = make_tree(d, keep)
    prm = make_params(d)
    if wh is None:
        wh, whd, K = tables_f()
        prm.K = K
    u32 = C.c_uint
    x, y, z = (np.ascontiguousarray(d[k]) for k in "xyz")
    h = d["h_in"].copy()
    m, vx, vy, vz, temp = (np.ascontiguousarray(d[k]) for k in ("m", "vx", "vy", "vz", "temp"))
    ngmax = prm.ngmax
    nb = np.zeros(n * ngmax, np.uint32)
    nc = np.zeros(n, np.uint32)
    fails = L.orc_find_neighbors_sph_f(P(x), P(y), P(z), P(h), u32(first), u32(last), C.byref(box), C.byref(tree),
                                       u32(prm.ng0), u32(ngmax), P(nb), P(nc))
    f32 = lambda: np.zeros(n, np.float32)  # noqa: E731
    xm, kx, gradh, prho, c = f32(), f32(), f32(), f32(), f32()
    c11, c12, c13, c22, c23, c33, divv, curlv = (f32() for _ in range(8))
    ax, ay, az = f32(), f32(), f32()
    du = np.zeros(n, np.float64)
    alpha = d["alpha_in"].copy()
    dtRho, dtCour = C.c_double(0), C.c_double(0)
    a = (u32(first), u32(last), C.byref(prm), C.byref(box), P(nb), P(nc), P(x), P(y), P(z))
    L.orc_xmass_f(*a, P(h), P(m), P(wh), P(xm))
    L.orc_ve_def_gradh_f(*a, P(h), P(m), P(wh), P(whd), P(xm), P(kx), P(gradh))
    L.orc_eos_ideal_temp_f(u32(first), u32(last), C.byref(prm), P(temp), P(m), P(kx), P(xm), P(gradh), P(prho), P(c))
    L.orc_iad_divv_curlv_f(*a, P(vx), P(vy), P(vz), P(h), P(wh), P(xm), P(kx), P(c11), P(c12), P(c13), P(c22), P(c23),
                           P(c33), P(divv), P(curlv), C.byref(dtRho))
    L.orc_av_switches_f(*a, P(vx), P(vy), P(vz), P(h), P(c), P(c11), P(c12), P(c13), P(c22), P(c23), P(c33), P(wh),
                        P(kx), P(xm), P(divv), P(alpha))
    L.orc_momentum_energy_f(*a, P(vx), P(vy), P(vz), P(h), P(m), P(prho), P(c), P(c11), P(c12), P(c13), P(c22),
                            P(c23), P(c33), P(wh), P(kx), P(xm), P(alpha), P(ax), P(ay), P(az), P(du),
                            C.byref(dtCour))
    return dict(h=h, nc=nc, neighbors=nb, xm=xm, kx=kx, gradh=gradh, prho=prho, c=c, c11=c11, c12=c12, c13=c13,
                c22=c22, c23=c23, c33=c33, divv=divv, curlv=curlv, alpha=alpha, ax=ax, ay=ay, az=az, du=du,
                dts=np.array([dtCour.value, dtRho.value]), fails=fails, wh=wh, whd=whd, K=prm.K)


def momentum_fields_d(d: dict, double_table: bool = False) -> dict:
    """ax, ay, az, du of a reference dump re-evaluated with EVERY operation in fp64 (the all-double instantiation of
    momentum_energy_kern.hpp:65-222) from the dump's own inputs of that loop (h, nc, neighbours, prho, c, c11..c33, kx,
    xm, alpha as the reference stored them, widened to double). |dump - this| is the reference's own fp32 rounding and
    summation noise on these fields."""
    L = lib()
    n = int(d["n"][0])
    ngmax = int(d["ngmax"][0])
    v = d["params"]
    box = make_box(d["box"], d["boundary"])
    # the production table is float: by default use ITS values (widened), so that only the arithmetic differs, not the
    # inputs; double_table: the table of the all-double instantiation (createWharmonicTable<double>)
    if double_table:
        wh, _, _ = tables_d()
    else:
        wh = d["wh"].astype(np.float64) if "wh" in d else tables_f()[0].astype(np.float64)
    dd = {k: np.ascontiguousarray(d[k], np.float64) for k in ("x", "y", "z", "vx", "vy", "vz", "h", "m", "prho", "c", "c11",
                                                             "c12", "c13", "c22", "c23", "c33", "kx", "xm", "alpha")}
    nb = np.ascontiguousarray(d["neighbors"], np.uint32)
    nc = np.ascontiguousarray(d["nc"], np.uint32)
    out = {k: np.zeros(n, np.float64) for k in ("ax", "ay", "az", "du")}
    D = C.c_double
    L.orc_momentum_energy_fields_d(C.c_uint(0), C.c_uint(n), C.c_uint(ngmax), D(v[0]), C.byref(box), P(nb), P(nc),
                                   *[P(dd[k]) for k in ("x", "y", "z", "vx", "vy", "vz", "h", "m", "prho", "c", "c11", "c12",
                                                        "c13", "c22", "c23", "c33")],
                                   D(np.float32(v[10])), D(np.float32(v[11])), D(np.float32(v[12])), P(wh), P(dd["kx"]),
                                   P(dd["xm"]), P(dd["alpha"]), P(out["ax"]), P(out["ay"]), P(out["az"]), P(out["du"]))
    return out
