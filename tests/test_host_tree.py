"""CPU tests of the host-side callers of the path: kernel tables, Hilbert keys, SFC order, octree view, and that
libsphx.so loads and exports every symbol of include/sphx.h (no compute call needs a GPU here)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from refdata import REPO, have_ref_harness, load_golden, run_ref_harness

sx = pytest.importorskip("sphexa_b200")

STEP_FILES = ["sedov12_step0.npz", "sedov12_step2.npz", "noh14_step0.npz", "turb12_step0.npz", "turb12h_step0.npz"]


def test_library_exports_every_declared_symbol():
    L = sx.load()
    header = (REPO / "include" / "sphx.h").read_text()
    declared = set(re.findall(r"\b(sphx_[a-z0-9_]+)\s*\(", header))
    assert declared == set(sx._cabi.EXPORTS)
    for name in declared:
        assert hasattr(L, name), name
    assert L.sphx_abi_version() == 1


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = sx.load()
    assert L.sphx_device_check() == 1  # SPHX_ERR_NO_DEVICE
    a = sx._cabi.SphxStepArgs()
    assert L.sphx_find_neighbors_xmass(C.byref(a), None) == 1
    assert b"no CPU fallback" in L.sphx_last_error()


def test_struct_sizes_match_header():
    """ctypes mirrors must have the C layout: compile a tiny probe against include/sphx.h"""
    import subprocess, tempfile
    src = '#include "sphx.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(SphxBox),' \
          'sizeof(SphxTreeView),sizeof(SphxFields),sizeof(SphxParams),sizeof(SphxStepArgs),sizeof(SphxStepResult),' \
          'sizeof(SphxGroups));return 0;}'
    with tempfile.TemporaryDirectory() as tmp:
        (Path(tmp) / "p.c").write_text(src)
        subprocess.run(["/usr/bin/gcc", "-I", str(REPO / "include"), str(Path(tmp) / "p.c"), "-o", str(Path(tmp) / "p")],
                       check=True)
        out = subprocess.run([str(Path(tmp) / "p")], check=True, stdout=subprocess.PIPE, text=True).stdout.split()
    c = sx._cabi
    got = [C.sizeof(t) for t in (c.SphxBox, c.SphxTreeView, c.SphxFields, c.SphxParams, c.SphxStepArgs,
                                 c.SphxStepResult)]
    assert got == [int(v) for v in out[:6]]


def test_tables_match_reference():
    g = load_golden("sedov12_step0.npz")
    wh, whd, K = sx.host.make_tables()
    assert K == g["params"][0]
    np.testing.assert_array_equal(wh, g["wh"])
    np.testing.assert_array_equal(whd, g["whd"])


def _check_tree(g):
    keys = sx.host.hilbert_keys(g["x"], g["y"], g["z"], g["box"])
    np.testing.assert_array_equal(keys, g["keys"])
    t = sx.host.build_tree(g["x"], g["y"], g["z"], g["box"], g["boundary"])
    # inputs are already SFC-sorted by the reference: our order must be the identity (stable sort)
    np.testing.assert_array_equal(t.order, np.arange(t.order.size))
    assert t.num_nodes == int(g["numNodes"][0]) and t.num_leaves == int(g["numLeafNodes"][0])
    for k in ("prefixes", "childOffsets", "levelRange", "leaves", "layout", "centers", "sizes"):
        np.testing.assert_array_equal(getattr(t, k), g["tree_" + k], err_msg=k)
    leaf = t.childOffsets == 0
    np.testing.assert_array_equal(t.internalToLeaf[leaf], g["tree_internalToLeaf"][leaf])


@pytest.mark.parametrize("fname", STEP_FILES)
def test_tree_matches_reference_golden(fname):
    _check_tree(load_golden(fname))


@pytest.mark.skipif(not have_ref_harness(), reason="oracle/_ref/ref_harness not present")
@pytest.mark.parametrize("case,n", [("sedov", 30), ("noh", 36), ("turb", 25)])
def test_tree_matches_compiled_reference(tmp_path, case, n):
    """non-uniform trees (several levels); step 0 = reference's converged focus tree"""
    d = run_ref_harness(case, n, 1, tmp_path / "o", dump_neighbors=False)[0]
    _check_tree(d)


def test_hilbert_is_a_bijection_with_unit_steps():
    """consecutive Hilbert keys are face neighbours (the defining property of the curve)"""
    m = 8  # 8^3 grid at level 3: keys are multiples of 8^(21-3)
    g = np.arange(m)
    ix, iy, iz = [a.ravel() for a in np.meshgrid(g, g, g, indexing="ij")]
    x, y, z = [(a + 0.5) / m for a in (ix, iy, iz)]
    keys = sx.host.hilbert_keys(x, y, z, [0, 1] * 3)
    assert np.unique(keys).size == m ** 3
    o = np.argsort(keys)
    step = np.abs(np.diff(ix[o])) + np.abs(np.diff(iy[o])) + np.abs(np.diff(iz[o]))
    assert (step == 1).all()
    assert (keys[o] >> np.uint64(3 * 18) == np.arange(m ** 3)).all()


def test_shuffled_input_is_sorted_back():
    g = load_golden("noh14_step0.npz")
    rng = np.random.default_rng(1)
    perm = rng.permutation(g["x"].size)
    t = sx.host.build_tree(g["x"][perm], g["y"][perm], g["z"][perm], g["box"], g["boundary"])
    np.testing.assert_array_equal(g["x"][perm][t.order], g["x"])
    np.testing.assert_array_equal(t.keys, g["keys"])
    np.testing.assert_array_equal(t.layout, g["tree_layout"])


def test_powf_emulation_matches_libm():
    """csrc/sphx_powf.h restates glibc's powf (FMA variant); the device uses the same code for sph::updateH.
    Compare with the libm of this box on the argument range updateH produces and on a wide random range."""
    L = sx.load()
    libm = C.CDLL("libm.so.6")
    libm.powf.restype = C.c_float
    libm.powf.argtypes = [C.c_float, C.c_float]
    rng = np.random.default_rng(0)
    ng0 = np.float32(100)
    bases = np.float32(1) + np.float32(1023.0) * ng0 / np.arange(1, 20000, dtype=np.float32)
    xs = np.concatenate([bases, rng.uniform(1, 2e5, 20000).astype(np.float32),
                         np.exp(rng.uniform(np.log(1e-30), np.log(1e30), 10000)).astype(np.float32)])
    for x in xs:
        assert L.sphx_powf_host(float(x), 0.1) == libm.powf(float(x), np.float32(0.1)), x
    # sph::updateH itself against the oracle restatement (which calls libm)
    import oracle_lib
    O = oracle_lib.lib()
    for nc in list(range(1, 400)) + [1000, 12345]:
        for h in (0.0123456, 0.5, 3.3e-4):
            assert L.sphx_update_h_host(100, nc, h) == O.orc_update_h_f(100, nc, h), (nc, h)
