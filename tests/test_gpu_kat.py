"""The reference's own known-answer vectors for the six J-loops (sph/test/ve.cpp:112-233 on sph/test/example_data.txt:
particle 0 with its 98 neighbours) pushed through the CUDA kernels of libsphx, through the C ABI, in the PRODUCTION type
set (x, y, z double; everything else float) at <= 1e-4 relative.

The data set is in cgs units (m = 3.8e26 g, x ~ 1e8 cm): fp32 intermediates such as m^2 overflow, in the reference's own
float instantiation as much as here. The SPH-VE equations are scale covariant, so the test feeds the same particles in
units of (L, M, V) = (h_0, m, c_0), and scales the answers back before comparing them with the reference's all-double
answers (tests/golden/ve_kat.npz: `ref_*` = the unmodified reference run on the file, `exp_*` = the literals of
ve.cpp).  The neighbour list is NOT given to the kernels: the block search has to find exactly particles 1..98 for
particle 0 (all lie inside 2 h_0, the farthest at 0.9983 x 2 h_0).
"""
import ctypes as C

import numpy as np
import pytest

from refdata import load_golden

pytestmark = pytest.mark.gpu

NAMES = ("x y z vx vy vz h c c11 c12 c13 c22 c23 c33 p gradh rho0 sumwhrho0 sumwh dvxdx dvxdy dvxdz dvydx dvydy "
         "dvydz dvzdx dvzdy dvzdz alpha u divv").split()
TOL = 1e-4


@pytest.fixture(scope="module")
def sx():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device: the product has no CPU fallback")
    import sphexa_b200
    sphexa_b200.load()
    return sphexa_b200


def _setup(sx, av_clean):
    from sphexa_b200.sim import HydroData, Params
    g = load_golden("ve_kat.npz")
    cols = np.ascontiguousarray(g["example_data"].T)
    f = {k: np.ascontiguousarray(cols[i]) for i, k in enumerate(NAMES)}
    K = float(g["ref_K"])
    mpart = 3.781038064465603e26
    # units
    L, M, V = f["h"][0], mpart, f["c"][0]
    T = L / V
    n = 99
    xm = (mpart / f["rho0"]) / L ** 3                       # volume element
    h = f["h"] / L
    kx = K * xm / h ** 3                                    # ve.cpp fixture: kx = K xm / h^3
    p = f["p"] * L * T * T / M
    s = dict(x=f["x"] / L, y=f["y"] / L, z=f["z"] / L, h=h, m=np.ones(n), vx=f["vx"] / V, vy=f["vy"] / V, vz=f["vz"] / V,
             c=f["c"] / V, xm=xm, kx=kx, gradh=f["gradh"], prho=p / (kx * 1.0 * 1.0 * f["gradh"]),
             c11=f["c11"] * L * L, c12=f["c12"] * L * L, c13=f["c13"] * L * L, c22=f["c22"] * L * L,
             c23=f["c23"] * L * L, c33=f["c33"] * L * L, divv=f["divv"] * T, alpha=f["alpha"],
             dV11=f["dvxdx"] * T, dV12=(f["dvxdy"] + f["dvydx"]) * T, dV13=(f["dvxdz"] + f["dvzdx"]) * T,
             dV22=f["dvydy"] * T, dV23=(f["dvydz"] + f["dvzdy"]) * T, dV33=f["dvzdz"] * T)
    lo = min(s["x"].min(), s["y"].min(), s["z"].min()) - 1.0
    hi = max(s["x"].max(), s["y"].max(), s["z"].max()) + 1.0
    box, boundary = [lo, hi] * 3, [0, 0, 0]
    t = sx.host.build_tree(s["x"], s["y"], s["z"], box, boundary, bucket_size=64)
    o = t.order
    i0 = int(np.nonzero(o == 0)[0][0])                      # where particle 0 of the file sits in SFC order
    prm = Params(K=K, ng0=100, ngmax=150, minDt=0.3 / T, alphamin=0.05, alphamax=1.0, decay_constant=0.2, Atmin=0.1,
                 Atmax=0.2, ramp=1.0 / (0.2 - 0.1), avClean=int(av_clean))
    hd = HydroData(n, 0, n, box, boundary, prm, device="cuda:0")
    hd.set_tree(t)
    sorted_ = {k: v[o] for k, v in s.items()}
    return g, hd, sorted_, i0, (L, M, V, T), o


def _put(hd, s, *names):
    hd.set_fields(**{k: s[k] for k in names if k in hd.f})


def _rel(a, b):
    return abs(a - b) / max(abs(a), abs(b))


@pytest.mark.parametrize("av_clean", [False, True])
def test_reference_known_answers_through_the_cuda_kernels(sx, oracle, av_clean):
    g, hd, s, i0, (L, M, V, T), order = _setup(sx, av_clean)
    ref = lambda k: float(g["ref_" + k])  # noqa: E731
    _put(hd, s, "x", "y", "z", "h", "m", "vx", "vy", "vz")
    # --- search: particle 0 sees exactly the 98 others, no h-iteration (ngmin = ng0 / 4) -------------------------------
    hd.find_neighbors_sph()
    assert hd.get("nc")[i0] == 99 and hd.get("h")[i0] == np.float32(s["h"][i0])
    nb = hd.export_neighbors().reshape(99, 150)[i0, :98]
    assert sorted(order[nb].tolist()) == list(range(1, 99))
    # --- XMass (ve.cpp:214-233): xm_0, rho0 = m / xm --------------------------------------------------------------------
    hd.xmass()
    assert _rel(hd.get("xm")[i0] * L ** 3, ref("xmass")) <= TOL
    val, tol = g["exp_xmass_rho0"]
    assert abs(M / (hd.get("xm")[i0] * L ** 3) - val) <= max(tol, TOL * abs(val))
    # --- VeDefGradh (ve.cpp:196-212) with the fixture's xm -----------------------------------------------------------------
    _put(hd, s, "xm")
    hd.ve_def_gradh()
    assert _rel(hd.get("kx")[i0], ref("gradh_kx")) <= TOL
    assert _rel(hd.get("gradh")[i0], ref("gradh_gradh")) <= TOL
    val, tol = g["exp_gradh_gradh"]
    assert abs(hd.get("gradh")[i0] - val) <= max(tol, TOL * abs(val))
    # --- IAD (ve.cpp:152-169) and divv / curlv (ve.cpp:123-150) -------------------------------------------------------
    _put(hd, s, "xm", "kx")
    hd.iad_divv_curlv()
    got_c = [hd.get(k)[i0] / (L * L) for k in ("c11", "c12", "c13", "c22", "c23", "c33")]
    cscale = max(abs(ref(f"iad_{k}")) for k in range(6))
    for k in range(6):
        assert abs(got_c[k] - ref(f"iad_{k}")) <= TOL * max(abs(ref(f"iad_{k}")), 1e-2 * cscale), k
        val, tol = g[f"exp_iad_{k}"]
        assert abs(got_c[k] - val) <= max(tol, TOL * max(abs(val), 1e-2 * cscale)), k
    # The reference's divv/curlv test feeds c_0 from the file, which is NOT the IAD result of the file's particles (it was
    # produced with another kernel: 5.7e-17 vs 1.9e-18), while the CUDA kernel computes both passes in one launch. The
    # expected values therefore come from the pinned all-double restatement (tests/test_oracle.py::test_ve_known_answers)
    # fed with the IAD answers above as c_0 - c_i is the only tensor that loop reads.
    f64 = {k: np.ascontiguousarray(v, np.float64) for k, v in s.items()}
    for k, name in enumerate(("c11", "c12", "c13", "c22", "c23", "c33")):
        f64[name] = f64[name].copy()
        f64[name][i0] = ref(f"iad_{k}") * L * L
    Lo, P = oracle.lib(), oracle.P
    wh, whd, _ = oracle.tables_d()
    obox = oracle.make_box(hd.box_lim, hd.boundary)
    nbu = np.ascontiguousarray(nb, np.uint32)
    out = np.zeros(8)
    Lo.orc_divv_curlv_jloop_d(C.c_uint(i0), C.c_double(hd.p.K), C.byref(obox), P(nbu), C.c_uint(98), P(f64["x"]),
                              P(f64["y"]), P(f64["z"]), P(f64["vx"]), P(f64["vy"]), P(f64["vz"]), P(f64["h"]),
                              *[P(f64[k]) for k in ("c11", "c12", "c13", "c22", "c23", "c33")], P(wh), P(f64["kx"]),
                              P(f64["xm"]), P(out))
    assert _rel(hd.get("divv")[i0], out[0]) <= TOL
    assert _rel(hd.get("curlv")[i0], out[1]) <= TOL
    if av_clean:
        dscale = np.abs(out[2:]).max()
        for k, name in enumerate(("dV11", "dV12", "dV13", "dV22", "dV23", "dV33")):
            assert abs(hd.get(name)[i0] - out[2 + k]) <= TOL * max(abs(out[2 + k]), 1e-2 * dscale), name
    # --- AV switches (ve.cpp:112-121) with the fixture's c_ij, divv, alpha --------------------------------------------------
    _put(hd, s, "c", "c11", "c12", "c13", "c22", "c23", "c33", "divv", "alpha")
    hd.av_switches()
    assert _rel(hd.get("alpha")[i0], ref("av_alpha")) <= TOL
    val, tol = g["exp_av_alpha"]
    assert abs(hd.get("alpha")[i0] - val) <= max(tol, TOL * abs(val))
    # --- momentum + energy (ve.cpp:171-194), avClean as parametrised ------------------------------------------------------
    _put(hd, s, "alpha", "prho", "dV11", "dV12", "dV13", "dV22", "dV23", "dV33")
    hd.momentum_energy()
    tag = f"mom{int(av_clean)}_"
    acc = V * V / L
    ascale = max(abs(ref(tag + k)) for k in ("ax", "ay", "az"))
    for k in ("ax", "ay", "az"):
        assert abs(hd.get(k)[i0] * acc - ref(tag + k)) <= TOL * max(abs(ref(tag + k)), 1e-2 * ascale), k
    assert _rel(hd.get("du")[i0] * V ** 3 / L, ref(tag + "du")) <= TOL
