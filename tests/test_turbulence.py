"""Turbulence stirring, CPU part: the host half of libsphx's driver (stirring modes, Ornstein-Uhlenbeck
phases, projection) and the oracle's restatement against dumps of the reference's own TurbulenceData / driveTurbulence
(tests/golden/turb12s_*.npz, turb_form2.npz; oracle/ref_harness.cpp stir=1|2).

The random numbers come from libstdc++ (std::mt19937, std::normal_distribution<double>: GCC 13, bits/random.tcc), a
dependency outside the reference tree; `std_normal_stream` below restates its published algorithm (Marsaglia polar
method on generate_canonical<double, 53>) so that the phase sequence is pinned independently of the library under test.
"""
import ctypes as C
import math

import numpy as np
import pytest

from refdata import load_golden

sx = pytest.importorskip("sphexa_b200")


def std_normal_stream(state: np.random.MT19937, n: int) -> np.ndarray:
    """n values of std::normal_distribution<double>(0, 1) of a FRESH distribution object driven by std::mt19937"""
    def canonical():
        lo, hi = (int(v) for v in state.random_raw(2))
        r = (lo + hi * 4294967296.0) / 18446744073709551616.0
        return r if r < 1.0 else np.nextafter(1.0, 0.0)
    out, saved = [], None
    while len(out) < n:
        if saved is not None:
            out.append(saved)
            saved = None
            continue
        while True:
            x = 2.0 * canonical() - 1.0
            y = 2.0 * canonical() - 1.0
            r2 = x * x + y * y
            if not (r2 > 1.0 or r2 == 0.0):
                break
        mult = math.sqrt(-2 * math.log(r2) / r2)  # libm, as libstdc++ calls it
        saved = x * mult
        out.append(y * mult)
    return np.array(out)


def std_mt19937(seed: int) -> np.random.MT19937:
    """std::mt19937(seed): Knuth's initialisation, the one numpy's legacy seeding uses"""
    rs = np.random.RandomState(seed)
    bg = np.random.MT19937()
    bg.state = rs.get_state(legacy=False)
    return bg


def test_modes_and_initial_phases_match_reference():
    """createStirringModes + initial Gaussian phases, parabolic spectrum (TurbulenceConstants as they are)"""
    g = load_golden("turb12s_step0.npz")
    t = sx.sim.Turbulence()
    s = t.state()
    assert t.num_modes == 112 and t.lattice and t.max_index == 3
    np.testing.assert_array_equal(s["modes"], g["turb_modes"])
    np.testing.assert_array_equal(s["amplitudes"], g["turb_amplitudes"])
    np.testing.assert_array_equal(s["phases"], g["turb_phases_in"])
    np.testing.assert_array_equal(s["scalars"], g["turb_scalars"])


def test_power_law_modes_match_reference():
    """stSpectForm = 2: the mode directions are drawn from the engine before the phases"""
    g = load_golden("turb_form2.npz")
    t = sx.sim.Turbulence(stSpectForm=2)
    s = t.state()
    np.testing.assert_array_equal(s["modes"], g["turb_modes"])
    np.testing.assert_array_equal(s["amplitudes"], g["turb_amplitudes"])
    np.testing.assert_array_equal(s["phases"], g["turb_phases_in"])
    # one OU step with the harness's minDt reproduces the reference's phases and projections
    t.advance_host(1e-4)
    s = t.state()
    np.testing.assert_array_equal(s["phases"], g["turb_phases"])
    np.testing.assert_array_equal(s["phasesReal"], g["turb_phasesReal"])
    np.testing.assert_array_equal(s["phasesImag"], g["turb_phasesImag"])


def test_ou_sequence_over_steps_matches_reference():
    """three driveTurbulence calls with the reference's time steps: phases at step 0 and step 2"""
    g0, g2 = load_golden("turb12s_step0.npz"), load_golden("turb12s_step2.npz")
    e = load_golden("turb12s_energies.npz")["series"]
    t = sx.sim.Turbulence()
    t.advance_host(float(e[0, 2]))
    s = t.state()
    for k in ("phases", "phasesReal", "phasesImag"):
        np.testing.assert_array_equal(s[k], g0["turb_" + k])
    t.advance_host(float(e[1, 2]))
    t.advance_host(float(e[2, 2]))
    s = t.state()
    for k in ("phases", "phasesReal", "phasesImag"):
        np.testing.assert_array_equal(s[k], g2["turb_" + k])


def test_restore_round_trip():
    """TurbulenceData::loadOrStore: phases + engine text restore the sequence"""
    a, b = sx.sim.Turbulence(), sx.sim.Turbulence(rngSeed=7)
    a.advance_host(1e-4)
    sa = a.state()
    b.restore(phases=sa["phases"], rng=sa["rng"])
    a.advance_host(2e-4), b.advance_host(2e-4)
    np.testing.assert_array_equal(a.state()["phases"], b.state()["phases"])
    np.testing.assert_array_equal(a.state()["phasesImag"], b.state()["phasesImag"])
    # replacing the mode set: off-lattice modes switch the device kernel to the mode-by-mode form
    b.restore(modes=sa["modes"] * 1.03, amplitudes=sa["amplitudes"], phases=sa["phases"])
    assert not b.lattice and b.num_modes == a.num_modes
    b.restore(modes=sa["modes"][:30], amplitudes=sa["amplitudes"][:10], phases=sa["phases"][:60])
    assert b.lattice and b.num_modes == 10


def test_oracle_phase_pipeline_matches_reference(oracle):
    """orc_update_noise (with libstdc++'s Gaussian stream restated above) and orc_compute_phases reproduce the reference
    dump bit for bit"""
    g = load_golden("turb12s_step0.npz")
    L = oracle.lib()
    nm = g["turb_amplitudes"].size
    gen = std_mt19937(251299)
    init = std_normal_stream(gen, 6 * nm) * g["turb_scalars"][0] + 0.0  # normal_distribution(0, variance)
    np.testing.assert_array_equal(init, g["turb_phases_in"])
    ph = init.copy()
    z = std_normal_stream(gen, 6 * nm)
    L.orc_update_noise(C.c_uint(ph.size), oracle.P(ph), C.c_double(g["turb_scalars"][0]), C.c_double(g["params"][5]),
                       C.c_double(g["turb_scalars"][1]), oracle.P(z))
    np.testing.assert_array_equal(ph, g["turb_phases"])
    pr, pi = np.zeros(3 * nm), np.zeros(3 * nm)
    L.orc_compute_phases(C.c_uint(nm), oracle.P(ph), C.c_double(g["turb_scalars"][2]), oracle.P(g["turb_modes"]),
                         oracle.P(pr), oracle.P(pi))
    np.testing.assert_array_equal(pr, g["turb_phasesReal"])
    np.testing.assert_array_equal(pi, g["turb_phasesImag"])


@pytest.mark.parametrize("fname", ["turb12s_step0.npz", "turb12s_step2.npz"])
def test_oracle_stirring_matches_reference(oracle, fname):
    """orc_compute_stirring == sph::computeStirring of the reference, bit for bit (same libm)"""
    g = load_golden(fname)
    L = oracle.lib()
    n = g["x"].size
    a = [g[k].copy() for k in ("ax", "ay", "az")]
    L.orc_compute_stirring(C.c_uint(0), C.c_uint(n), oracle.P(g["x"]), oracle.P(g["y"]), oracle.P(g["z"]),
                           oracle.P(a[0]), oracle.P(a[1]), oracle.P(a[2]), C.c_uint(g["turb_amplitudes"].size),
                           oracle.P(g["turb_modes"]), oracle.P(g["turb_phasesReal"]), oracle.P(g["turb_phasesImag"]),
                           oracle.P(g["turb_amplitudes"]), C.c_double(g["turb_scalars"][3]))
    for got, k in zip(a, ("stir_ax", "stir_ay", "stir_az")):
        np.testing.assert_array_equal(got, g[k])
    assert np.abs(g["stir_ax"] - g["ax"]).max() > 0


def test_stirring_needs_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    t = sx.sim.Turbulence()
    L = sx.load()
    assert L.sphx_drive_turbulence(t.handle, None, None, None, None, None, None, 0, 0, 1e-4, None) == 1
