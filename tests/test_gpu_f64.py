"""The all-double type set of libsphx (sphx_*_f64, csrc/loops_f64.cu) against the reference's all-double answers at
<= 1e-10 (BASELINE.json north_star: "<= 1e-10 fp64"):

  * the reference's own known-answer vectors (sph/test/ve.cpp:112-233 on sph/test/example_data.txt, `using T = double`) in
    their ORIGINAL cgs units - the fp64 kernels have the range for m^2 ~ 1e53 that the production type set lacks
    (tests/test_gpu_kat.py runs the same vectors through the production kernels in rescaled units at 1e-4);
  * the all-double momentum / energy loop of the pinned restatement on live dumps of the compiled reference;
  * and, as the on-device yardstick it is meant to be, the production path itself: the two share no pair code.
"""
import ctypes as C

import numpy as np
import pytest

from refdata import csr_sorted_neighbors, load_golden

pytestmark = pytest.mark.gpu

NAMES = ("x y z vx vy vz h c c11 c12 c13 c22 c23 c33 p gradh rho0 sumwhrho0 sumwh dvxdx dvxdy dvxdz dvydx dvydy "
         "dvydz dvzdx dvzdy dvzdz alpha u divv").split()
TOL = 1e-10


@pytest.fixture(scope="module")
def sx():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device: the product has no CPU fallback")
    import sphexa_b200
    sphexa_b200.load()
    return sphexa_b200


def _rel(a, b):
    return abs(a - b) / max(abs(a), abs(b))


@pytest.mark.parametrize("av_clean", [False, True])
def test_reference_known_answers_all_double(sx, oracle, av_clean):
    from sphexa_b200.sim import HydroDataF64, Params
    g = load_golden("ve_kat.npz")
    cols = np.ascontiguousarray(g["example_data"].T)
    f = {k: np.ascontiguousarray(cols[i]) for i, k in enumerate(NAMES)}
    K = float(g["ref_K"])
    mpart = 3.781038064465603e26
    n = 99
    s = dict(f)
    s["m"] = np.full(n, mpart)
    s["xm"] = mpart / f["rho0"]
    s["kx"] = K * s["xm"] / f["h"] ** 3
    s["prho"] = f["p"] / (s["kx"] * s["m"] * s["m"] * f["gradh"])
    s["dV11"], s["dV12"], s["dV13"] = f["dvxdx"], f["dvxdy"] + f["dvydx"], f["dvxdz"] + f["dvzdx"]
    s["dV22"], s["dV23"], s["dV33"] = f["dvydy"], f["dvydz"] + f["dvzdy"], f["dvzdz"]
    lo = min(f["x"].min(), f["y"].min(), f["z"].min()) - 1e9
    hi = max(f["x"].max(), f["y"].max(), f["z"].max()) + 1e9
    box, boundary = [lo, hi] * 3, [0, 0, 0]
    t = sx.host.build_tree(f["x"], f["y"], f["z"], box, boundary, bucket_size=64)
    o = t.order
    i0 = int(np.nonzero(o == 0)[0][0])
    prm = Params(K=K, ng0=100, ngmax=150, minDt=0.3, alphamin=0.05, alphamax=1.0, decay_constant=0.2, Atmin=0.1, Atmax=0.2,
                 ramp=1.0 / (0.2 - 0.1), avClean=int(av_clean))
    hd = HydroDataF64(n, box, boundary, prm)
    hd.set_tree(t)
    srt = {k: v[o] for k, v in s.items() if k in hd.f}
    put = lambda *names: hd.set_fields(**{k: srt[k] for k in names if k in hd.f})  # noqa: E731
    ref = lambda k: float(g["ref_" + k])  # noqa: E731

    put("x", "y", "z", "h", "m", "vx", "vy", "vz")
    hd.find_neighbors_sph()
    assert hd.get("nc")[i0] == 99 and hd.get("h")[i0] == srt["h"][i0]
    nb = hd.export_neighbors().reshape(n, 150)[i0, :98]
    assert sorted(o[nb].tolist()) == list(range(1, 99))
    hd.xmass()
    assert _rel(hd.get("xm")[i0], ref("xmass")) <= TOL
    put("xm")
    hd.ve_def_gradh()
    assert _rel(hd.get("kx")[i0], ref("gradh_kx")) <= TOL and _rel(hd.get("gradh")[i0], ref("gradh_gradh")) <= TOL
    put("xm", "kx")
    hd.iad_divv_curlv()
    cscale = max(abs(ref(f"iad_{k}")) for k in range(6))
    for k, name in enumerate(("c11", "c12", "c13", "c22", "c23", "c33")):
        assert abs(hd.get(name)[i0] - ref(f"iad_{k}")) <= TOL * cscale, name
    # divv / curlv of the fused kernel: expected from the pinned restatement fed with the IAD answers as c_0 (the file's
    # own c_0 is not the IAD result of its particles, see tests/test_gpu_kat.py)
    f64 = {k: np.ascontiguousarray(v, np.float64).copy() for k, v in srt.items()}
    for k, name in enumerate(("c11", "c12", "c13", "c22", "c23", "c33")):
        f64[name][i0] = ref(f"iad_{k}")
    Lo, P = oracle.lib(), oracle.P
    wh, whd, _ = oracle.tables_d()
    obox = oracle.make_box(box, boundary)
    nbu = np.ascontiguousarray(nb, np.uint32)
    out = np.zeros(8)
    Lo.orc_divv_curlv_jloop_d(C.c_uint(i0), C.c_double(K), C.byref(obox), P(nbu), C.c_uint(98), P(f64["x"]), P(f64["y"]),
                              P(f64["z"]), P(f64["vx"]), P(f64["vy"]), P(f64["vz"]), P(f64["h"]),
                              *[P(f64[k]) for k in ("c11", "c12", "c13", "c22", "c23", "c33")], P(wh), P(f64["kx"]),
                              P(f64["xm"]), P(out))
    assert _rel(hd.get("divv")[i0], out[0]) <= 1e-9 and _rel(hd.get("curlv")[i0], out[1]) <= 1e-9
    if av_clean:
        dscale = np.abs(out[2:]).max()
        for k, name in enumerate(("dV11", "dV12", "dV13", "dV22", "dV23", "dV33")):
            assert abs(hd.get(name)[i0] - out[2 + k]) <= 1e-9 * dscale, name
    put("c", "c11", "c12", "c13", "c22", "c23", "c33", "divv", "alpha")
    hd.av_switches()
    assert _rel(hd.get("alpha")[i0], ref("av_alpha")) <= TOL
    val, tol = g["exp_av_alpha"]
    assert abs(hd.get("alpha")[i0] - val) <= tol          # the literal of ve.cpp:120 within ITS tolerance (2e-9)
    put("alpha", "prho", "dV11", "dV12", "dV13", "dV22", "dV23", "dV33")
    hd.momentum_energy()
    tag = f"mom{int(av_clean)}_"
    ascale = max(abs(ref(tag + k)) for k in ("ax", "ay", "az"))
    for k in ("ax", "ay", "az"):
        assert abs(hd.get(k)[i0] - ref(tag + k)) <= TOL * ascale, k
        val, tol = g["exp_" + tag + k]
        assert abs(hd.get(k)[i0] - val) <= tol, k          # ve.cpp:171-194
    assert _rel(hd.get("du")[i0], ref(tag + "du")) <= TOL


@pytest.mark.parametrize("fname", ["turb12h_step0.npz", "noh14_step0.npz", "sedov12_step2.npz"])
def test_all_double_step_is_the_yardstick_of_the_production_path(sx, oracle, fname):
    """(1) momentum / energy in all-double on the reference's own loop inputs == the pinned all-double restatement to
    1e-10 of the field scale; (2) a whole all-double step from the same initial state agrees with the production step of
    libsphx (mixed precision, block search, staged candidates) to the production tolerance: two implementations that share
    no pair code."""
    from sphexa_b200.sim import HydroDataF64, Params
    from test_gpu_parity import assert_fields_close, F32_FIELDS
    d = load_golden(fname)
    n, ngmax = int(d["n"][0]), int(d["ngmax"][0])
    prm = Params.from_dump(d)
    hd = HydroDataF64(n, d["box"], d["boundary"], prm)
    hd.set_tree(d)
    # (1) the reference's converged h, then the momentum loop on the reference's own intermediate fields
    hd.set_fields(x=d["x"], y=d["y"], z=d["z"], h=d["h"], m=d["m"], vx=d["vx"], vy=d["vy"], vz=d["vz"])
    hd.find_neighbors_sph()
    np.testing.assert_array_equal(hd.get("nc"), d["nc"])
    off, idx = csr_sorted_neighbors(hd.export_neighbors(), hd.get("nc"), ngmax)
    np.testing.assert_array_equal(idx, d["nb_sorted"])
    hd.set_fields(**{k: d[k] for k in ("prho", "c", "c11", "c12", "c13", "c22", "c23", "c33", "kx", "xm", "alpha")})
    hd.momentum_energy()
    dd = dict(d)
    dd["neighbors"] = hd.export_neighbors()
    exact = oracle.momentum_fields_d(dd, double_table=True)
    fam = max(np.abs(exact[k]).max() for k in ("ax", "ay", "az"))
    for k in ("ax", "ay", "az", "du"):
        scale = np.abs(exact[k]).max() if k == "du" else fam
        assert np.abs(hd.get(k) - exact[k]).max() <= 1e-10 * max(scale, 1e-300), k
    # (2) whole step, all-double vs production
    hd.set_fields(h=d["h_in"], temp=d["temp"], alpha=d["alpha_in"])
    hd.hydro_step()
    prod = sx.sim.from_dump(d)
    prod.hydro_step()
    np.testing.assert_array_equal(hd.get("nc"), prod.get("nc"))
    np.testing.assert_allclose(hd.get("h"), prod.get("h"), rtol=1e-6)  # up to 10 float roundings of the h-iteration
    assert_fields_close({k: prod.get(k) for k in F32_FIELDS}, {k: hd.get(k) for k in F32_FIELDS})
