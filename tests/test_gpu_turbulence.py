"""Turbulence stirring on the device, through the C ABI: sphx_compute_stirring / sphx_drive_turbulence
against dumps of the reference's turbulence-ve propagator (TurbVeProp, main/src/propagator/turb_ve.hpp:67-72;
oracle/ref_harness.cpp stir=1) and against the oracle's restatement of sph::computeStirring.

Tolerance: the host half (modes, OU phases, projection) is bit-exact (tests/test_turbulence.py). The device half
evaluates the reference's expressions with separately rounded fp64 operations; what differs is the last bit of CUDA's
sin/cos against glibc's, i.e. ~1e-16 relative per mode term before the fp32 accumulation, which can flip the rounding of
an fp32 partial sum: |a - b| <= 2^-22 * max|a| over the particle set (a few fp32 ulps of the largest acceleration).
"""
import ctypes as C

import numpy as np
import pytest

from refdata import GOLDEN, load_golden

pytestmark = pytest.mark.gpu

STIR_TOL = 2.0 ** -22


@pytest.fixture(scope="module")
def sx():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device: the product has no CPU fallback")
    import sphexa_b200
    sphexa_b200.load()
    return sphexa_b200


def _stir_on_device(sx, t, x, y, z, ax, ay, az, first=0, last=None):
    import torch
    L = sx.load()
    dev = torch.device("cuda:0")
    X, Y, Z = (torch.from_numpy(np.ascontiguousarray(a, np.float64)).to(dev) for a in (x, y, z))
    A = [torch.from_numpy(np.ascontiguousarray(a, np.float32)).to(dev) for a in (ax, ay, az)]
    last = X.numel() if last is None else last
    from sphexa_b200 import _cabi
    _cabi.check(L.sphx_compute_stirring(t.handle, X.data_ptr(), Y.data_ptr(), Z.data_ptr(), A[0].data_ptr(),
                                        A[1].data_ptr(), A[2].data_ptr(), first, last, None))
    torch.cuda.synchronize()
    return [a.cpu().numpy() for a in A]


def _close(got, exp, scale):
    err = np.abs(got.astype(np.float64) - exp.astype(np.float64)).max()
    assert err <= STIR_TOL * scale, (err, scale)


def test_stirring_step0_matches_reference(sx):
    """one driveTurbulence on the reference's step-0 state: phases bit-exact, accelerations within STIR_TOL"""
    g = load_golden("turb12s_step0.npz")
    e = load_golden("turb12s_energies.npz")["series"]
    t = sx.sim.Turbulence()
    t.advance_host(float(e[0, 2]))
    np.testing.assert_array_equal(t.state()["phasesReal"], g["turb_phasesReal"])
    got = _stir_on_device(sx, t, g["x"], g["y"], g["z"], g["ax"], g["ay"], g["az"])
    scale = max(np.abs(g[k]).max() for k in ("stir_ax", "stir_ay", "stir_az"))
    for a, k in zip(got, ("stir_ax", "stir_ay", "stir_az")):
        _close(a, g[k], scale)
    # the stirring term is not a rounding-level change of the accelerations
    assert np.abs(g["stir_ax"] - g["ax"]).max() > 1e3 * STIR_TOL * scale


def test_stirring_step2_matches_reference(sx):
    """third call: OU sequence advanced with the reference's time steps, particles have moved"""
    g = load_golden("turb12s_step2.npz")
    e = load_golden("turb12s_energies.npz")["series"]
    t = sx.sim.Turbulence()
    for k in range(3):
        t.advance_host(float(e[k, 2]))
    np.testing.assert_array_equal(t.state()["phasesImag"], g["turb_phasesImag"])
    got = _stir_on_device(sx, t, g["x"], g["y"], g["z"], g["ax"], g["ay"], g["az"])
    scale = max(np.abs(g[k]).max() for k in ("stir_ax", "stir_ay", "stir_az"))
    for a, k in zip(got, ("stir_ax", "stir_ay", "stir_az")):
        _close(a, g[k], scale)


@pytest.mark.parametrize("form", ["lattice", "off_lattice", "power_law"])
def test_stirring_kernels_vs_oracle(sx, oracle, form):
    """both device kernels (lattice table form, mode-by-mode form) against the restatement of sph::computeStirring on
    random particles, a sub-range [first, last) only, non-zero accelerations coming in"""
    rng = np.random.default_rng(5)
    n, first, last = 5000, 37, 4901
    x, y, z = (rng.uniform(-0.5, 0.5, n) for _ in range(3))
    a0 = [rng.normal(0, 30.0, n).astype(np.float32) for _ in range(3)]
    t = sx.sim.Turbulence(stSpectForm=2) if form == "power_law" else sx.sim.Turbulence()
    if form == "off_lattice":
        s = t.state()
        t.restore(modes=s["modes"] * 1.03, amplitudes=s["amplitudes"], phases=s["phases"])
        assert not t.lattice
    else:
        assert t.lattice
    t.advance_host(2e-4)
    s = t.state()
    exp = [a.copy() for a in a0]
    L = oracle.lib()
    L.orc_compute_stirring(C.c_uint(first), C.c_uint(last), oracle.P(x), oracle.P(y), oracle.P(z), oracle.P(exp[0]),
                           oracle.P(exp[1]), oracle.P(exp[2]), C.c_uint(t.num_modes), oracle.P(s["modes"]),
                           oracle.P(s["phasesReal"]), oracle.P(s["phasesImag"]), oracle.P(s["amplitudes"]),
                           C.c_double(s["scalars"][3]))
    got = _stir_on_device(sx, t, x, y, z, *a0, first=first, last=last)
    scale = max(np.abs(e_).max() for e_ in exp)
    for a, e_, o in zip(got, exp, a0):
        _close(a, e_, scale)
        assert np.array_equal(a[:first], o[:first]) and np.array_equal(a[last:], o[last:])
        assert np.abs(e_[first:last] - o[first:last]).max() > 0


def test_stirring_empty_range_and_restore_after_upload(sx):
    """empty [first, last) is a no-op; replacing the mode set after the device tables exist rebuilds them"""
    g = load_golden("turb12s_step0.npz")
    t = sx.sim.Turbulence()
    t.advance_host(1e-4)
    same = _stir_on_device(sx, t, g["x"], g["y"], g["z"], g["ax"], g["ay"], g["az"], first=5, last=5)
    assert np.array_equal(same[0], g["ax"])
    full = _stir_on_device(sx, t, g["x"], g["y"], g["z"], g["ax"], g["ay"], g["az"])
    s = t.state()
    keep = 40
    t.restore(modes=s["modes"][: 3 * keep], amplitudes=s["amplitudes"][:keep], phases=s["phases"][: 6 * keep])
    t.advance_host(0.0)
    part = _stir_on_device(sx, t, g["x"], g["y"], g["z"], g["ax"], g["ay"], g["az"])
    assert t.num_modes == keep and not np.array_equal(part[0], full[0])


def test_turbulence_ve_loop_energy_series(sx):
    """the reference's turbulence-ve propagator over 40 steps (particles start at rest; all kinetic energy comes from the
    stirring): native loop (sync, hydro step, driveTurbulence, conserved, integrate) vs the reference series"""
    from sphexa_b200 import sim
    d = load_golden("turb12s_step0.npz")
    with np.load(GOLDEN / "turb12s_energies.npz") as z:
        ref = z["series"]  # step ttot minDt etot ecin eint linmom angmom totalNeighbors
    s = sim.simulation_from_dump(d)
    s.turbulence = sim.Turbulence()
    got = np.array([s.step() for _ in range(ref.shape[0])], dtype=np.float64)
    np.testing.assert_allclose(got[:, 2], ref[:, 2], rtol=2e-5)               # minDt
    np.testing.assert_allclose(got[:, 3], ref[:, 3], rtol=1e-6)               # etot
    np.testing.assert_allclose(got[1:, 4], ref[1:, 4], rtol=1e-4)             # ecin: grows from 0 through the stirring
    assert got[0, 4] == 0.0 and ref[0, 4] == 0.0
    np.testing.assert_allclose(got[:, 5], ref[:, 5], rtol=1e-6)               # eint
    np.testing.assert_allclose(got[1:, 6], ref[1:, 6], rtol=1e-3)             # |linear momentum|
    assert np.array_equal(got[:3, 8], ref[:3, 8])
    np.testing.assert_allclose(got[:, 8], ref[:, 8], rtol=2e-3)
    # OU state after the loop equals the host-only sequence driven by the same time steps
    t = sim.Turbulence()
    for dt in got[:, 2]:  # the row's minDt is the one its step's driveTurbulence used
        t.advance_host(float(dt))
    np.testing.assert_array_equal(s.turbulence.state()["phases"], t.state()["phases"])
