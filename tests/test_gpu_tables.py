"""How the loops evaluate the kernel tables (include/sphx.h: sphx_table_mode, sphx_invalidate_tables): polynomial fits
of the caller's tables by default, shared-memory copies of the tables (lt::lookup as written,
sph/include/sph/table_lookup.hpp:13-26) when the polynomials do not reproduce them. Both instantiations are held to the
reference goldens; a table rewritten in place is reported, not silently evaluated through a stale fit."""
import ctypes as C
import os

import numpy as np
import pytest

from refdata import load_golden
from test_gpu_parity import STEP_FILES, check_against_reference, run_step_by_loops, sx  # noqa: F401

pytestmark = pytest.mark.gpu


def table_mode(sx, hd):
    ew, ed = C.c_double(), C.c_double()
    m = hd.L.sphx_table_mode(hd.wh.data_ptr(), hd.whd.data_ptr(), None, C.byref(ew), C.byref(ed))
    return m, ew.value, ed.value


@pytest.fixture
def forced_tables(sx):
    """every table pair first seen inside the test runs the shared-memory-table instantiations"""
    L = sx.load()
    os.environ["SPHX_FORCE_TABLE"] = "1"
    L.sphx_invalidate_tables()
    yield
    del os.environ["SPHX_FORCE_TABLE"]
    L.sphx_invalidate_tables()


def test_default_tables_run_as_polynomials(sx):
    ref = load_golden(STEP_FILES[0])
    got, hd = run_step_by_loops(sx, ref)
    mode, ew, ed = table_mode(sx, hd)
    # sinc^6: reproduced to the fp32 noise of the Horner evaluation, well inside the 1e-6 the fit accepts
    assert mode == 1 and ew < 1e-6 and ed < 1e-6, (mode, ew, ed)


@pytest.mark.parametrize("fname", STEP_FILES)
def test_shared_memory_tables_vs_reference_golden(sx, forced_tables, fname):
    ref = load_golden(fname)
    got, hd = run_step_by_loops(sx, ref)
    assert table_mode(sx, hd)[0] == 0
    check_against_reference(got, ref)


def test_polynomials_and_tables_agree(sx):
    """the two instantiations differ by the rounding noise of lt::lookup (1e-7 of the table maximum) and of the Horner
    chain (3e-7): sums of ~100 kernel values agree to a few 1e-6 of the field scale (the terms of divv and of the
    accelerations cancel, so their sums carry the noise of the individual terms), far inside the 1e-4 parity tolerance"""
    ref = load_golden("turb12h_step0.npz")
    got_p, hd = run_step_by_loops(sx, ref)
    assert table_mode(sx, hd)[0] == 1
    L = sx.load()
    os.environ["SPHX_FORCE_TABLE"] = "1"
    L.sphx_invalidate_tables()
    try:
        got_t, hd_t = run_step_by_loops(sx, ref)
        assert table_mode(sx, hd_t)[0] == 0
    finally:
        del os.environ["SPHX_FORCE_TABLE"]
        L.sphx_invalidate_tables()
    np.testing.assert_array_equal(got_p["nc"], got_t["nc"])
    for k in ("xm", "kx", "gradh", "c11", "c22", "c33", "divv", "alpha", "ax", "ay", "az", "du"):
        a, b = got_p[k].astype(np.float64), got_t[k].astype(np.float64)
        scale = np.abs(b).max()
        assert np.abs(a - b).max() <= 2e-5 * scale, (k, np.abs(a - b).max() / scale)


def test_sharp_kernel_keeps_the_tables(sx):
    """sinc^9 is not reproduced to 1e-6 by the degree-13 polynomial: the loops must fall back to the tables"""
    ref = load_golden(STEP_FILES[0])
    hd = sx.sim.from_dump(ref)
    import torch
    wh, whd, _ = sx.host.make_tables(9.0)
    hd.wh, hd.whd = torch.from_numpy(wh).to(hd.device), torch.from_numpy(whd).to(hd.device)
    hd.L.sphx_invalidate_tables()  # (the allocator may hand out addresses an earlier test's tables had)
    mode, ew, ed = table_mode(sx, hd)
    assert mode == 0 and max(ew, ed) > 1e-6, (mode, ew, ed)
    hd.hydro_step()  # runs (shared-memory tables)
    assert np.isfinite(hd.get("ax")).all()
    hd.L.sphx_invalidate_tables()


def test_table_rewritten_in_place_is_reported(sx):
    import torch
    ref = load_golden(STEP_FILES[0])

    def fresh(tables):
        hd = sx.sim.from_dump(ref)
        if tables is not None:
            hd.wh, hd.whd = tables
        return hd

    hd = fresh(None)
    tables = (hd.wh, hd.whd)
    hd.L.sphx_invalidate_tables()
    hd.hydro_step()
    ax0 = hd.get("ax").copy()
    wh5, whd5, _ = sx.host.make_tables(5.0)
    wh6, whd6 = hd.wh.clone(), hd.whd.clone()
    hd.wh.copy_(torch.from_numpy(wh5).to(hd.device))   # same addresses, another kernel
    hd.whd.copy_(torch.from_numpy(whd5).to(hd.device))
    with pytest.raises(sx._cabi.SphxError, match="SPHX_ERR_TABLE|kernel tables"):
        fresh(tables).hydro_step()
    hd.L.sphx_invalidate_tables()
    hd5 = fresh(tables)
    hd5.hydro_step()                                   # refitted to the sinc^5 tables
    assert table_mode(sx, hd5)[0] == 1
    assert np.abs(hd5.get("ax") - ax0).max() > 0
    tables[0].copy_(wh6), tables[1].copy_(whd6)
    hd.L.sphx_invalidate_tables()
    hd6 = fresh(tables)
    hd6.hydro_step()
    np.testing.assert_array_equal(hd6.get("ax"), ax0)
