"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path through the C ABI (libsphx.so) against
(1) golden fixtures produced by the unmodified reference, (2) the CPU restatement (oracle/), (3) the compiled
reference itself (oracle/_ref/ref_harness) on larger cases, and (4) size-independent properties at larger sizes.

Tolerances (BASELINE.json north_star): neighbour counts and sorted neighbour sets bit-exact; fp32 fields <= 1e-4
relative (measured against the field's max-norm floor, SURVEY §8c: |a-b| <= 1e-4 max(|a|,|b|,floor)).
"""
import numpy as np
import pytest

from refdata import csr_sorted_neighbors, have_ref_harness, load_golden, run_ref_harness

pytestmark = pytest.mark.gpu

REL_TOL_F32 = 1e-4
F32_FIELDS = ["xm", "kx", "gradh", "prho", "c", "c11", "c12", "c13", "c22", "c23", "c33", "divv", "curlv", "alpha",
              "ax", "ay", "az", "du"]
STEP_FILES = ["sedov12_step0.npz", "sedov12_step2.npz", "noh14_step0.npz", "turb12_step0.npz", "turb12h_step0.npz"]


@pytest.fixture(scope="module")
def sx():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device: the product has no CPU fallback")
    import sphexa_b200
    sphexa_b200.load()
    return sphexa_b200


# Fields that vanish by symmetry or cancellation (off-diagonal IAD terms on a lattice, divv of a solenoidal flow,
# accelerations in a uniform region) carry only fp32 summation noise of their ~100 pair terms, so "relative" error is
# measured against the magnitude of the field FAMILY: floor = 1e-2 x the family's max-norm, i.e. the absolute error
# allowed for a vanishing value is 1e-6 of the family scale.
FAMILIES = [("c11", "c12", "c13", "c22", "c23", "c33"), ("divv", "curlv"), ("ax", "ay", "az")]


# du = K prho_i sum_j m_j a_mom (v_ij . A_ij) and a = -K sum_j (pressure-gradient pair terms): the ~100 pair terms of a
# particle cancel to a few per cent of their magnitude (steep but smooth pressure field of the Sedov blob, near-uniform
# subsonic turbulence), so ANY fp32 evaluation carries an absolute error that is a fixed fraction of the field scale,
# whatever the value it belongs to. The reference's OWN production evaluation shows it: against the same loop carried
# out in all-double on the same inputs (oracle.momentum_fields_d) it is off by up to 2.2e-6 x max|a| (Sedov 64^3),
# 1.7e-6 (turbulence), 4.7e-6 x max|du|, i.e. up to 1.5e-2 RELATIVE for small values; the CUDA path is off by up to
# 4.4e-6 x max|a| against the same all-double values (fp32 positions relative to the block origin instead of fp64
# differences), and the two fp32 evaluations differ from each other by up to 5.5e-6 x scale
# (profiles/r02_floor_analysis.json, tools/floor_analysis.py; test_floors_are_the_references_own_fp32_noise below keeps
# the numbers honest). The floor of these four fields is therefore 1e-1 x the family max-norm: the absolute error allowed
# for a nearly cancelled value is 1e-5 of the field scale, twice what two correct fp32 evaluations are observed to
# differ by, and every value above a tenth of the scale is held to the plain 1e-4 relative tolerance.
FLOOR_FRACTION = {"du": 1e-1, "ax": 1e-1, "ay": 1e-1, "az": 1e-1}


def field_floor(ref: dict, k: str) -> float:
    fam = next((f for f in FAMILIES if k in f), (k,))
    scale = max(float(np.abs(ref[m]).max()) for m in fam)
    return max(scale, 1e-300) * FLOOR_FRACTION.get(k, 1e-2)


def assert_fields_close(got: dict, ref: dict, fields=F32_FIELDS, tol=REL_TOL_F32):
    for k in fields:
        a, b = got[k].astype(np.float64), ref[k].astype(np.float64)
        denom = np.maximum(np.maximum(np.abs(a), np.abs(b)), field_floor(ref, k))
        err = np.abs(a - b) / denom
        assert err.max() <= tol, f"{k}: max rel err {err.max():.3e} at {err.argmax()} ({a[err.argmax()]} vs {b[err.argmax()]})"
        assert np.sqrt((err ** 2).mean()) <= tol / 4, f"{k}: rms rel err {np.sqrt((err ** 2).mean()):.3e}"


def run_step_by_loops(sx, d):
    """one C-ABI call per loop, in the reference order (ve_hydro.hpp:147-190)"""
    hd = sx.sim.from_dump(d)
    hd.find_neighbors_xmass()
    nb = hd.export_neighbors()
    hd.ve_def_gradh()
    hd.eos()
    hd.iad_divv_curlv()
    hd.av_switches()
    hd.momentum_energy()
    out = {k: hd.get(k) for k in sx.sim.STEP_OUTPUTS}
    out["neighbors"] = nb
    out["dts"] = np.array([hd.result.minDtCourant, hd.result.minDtRho])
    out["totalNeighbors"] = hd.result.totalNeighbors
    return out, hd


CANCELLING = ("ax", "ay", "az", "du")


def assert_within_reference_noise(got: dict, d: dict, oracle, tol=REL_TOL_F32, factor=3.0):
    """du, ax, ay, az at resolutions where the pair terms cancel ever more strongly (the fp32 noise of these sums is a
    fraction of the field scale that grows like width / h: 2e-6 of max|a| at Sedov 64^3, 4e-6 at 100^3 for the
    reference itself): the yardstick is the all-double evaluation of the same loop on the reference's own inputs. The
    CUDA value must be within 1e-4 relative of it, or within `factor` x the LARGEST error the reference's own fp32
    evaluation makes against it on that field."""
    exact = oracle.momentum_fields_d(d)
    out = {}
    for k in CANCELLING:
        e, r32, g = exact[k], d[k].astype(np.float64), got[k].astype(np.float64)
        noise = np.abs(r32 - e).max()
        err = np.abs(g - e)
        bad = err > np.maximum(tol * np.abs(e), factor * noise)
        out[k] = (float(noise), float(err.max()))
        assert not bad.any(), f"{k}: {int(bad.sum())} values off by more than {factor} x the reference's own fp32 noise " \
                              f"{noise:.3e} (max err {err.max():.3e}, scale {np.abs(e).max():.3e})"
    return out


def check_against_reference(got, ref, oracle=None):
    """oracle given: du / a are judged against the all-double values and the reference's own noise instead of the
    fixed floors (assert_within_reference_noise)"""
    ngmax = int(ref["ngmax"][0])
    np.testing.assert_array_equal(got["h"], ref["h"])        # h-iteration trajectory bit-exact
    np.testing.assert_array_equal(got["nc"], ref["nc"])      # neighbour counts bit-exact
    off, idx = csr_sorted_neighbors(got["neighbors"], got["nc"], ngmax)
    np.testing.assert_array_equal(off, ref["nb_offsets"])
    np.testing.assert_array_equal(idx, ref["nb_sorted"])     # sorted neighbour sets bit-exact
    if oracle is None:
        assert_fields_close(got, ref)
    else:
        assert_fields_close(got, ref, [k for k in F32_FIELDS if k not in CANCELLING])
        assert_within_reference_noise(got, ref, oracle)
    np.testing.assert_allclose(got["dts"], ref["dts"], rtol=1e-4)
    assert got["totalNeighbors"] == int(ref["nc"].astype(np.int64).sum())


@pytest.mark.parametrize("fname", STEP_FILES)
def test_step_vs_reference_golden(sx, fname):
    ref = load_golden(fname)
    got, _ = run_step_by_loops(sx, ref)
    check_against_reference(got, ref)


DV_FIELDS = ["dV11", "dV12", "dV13", "dV22", "dV23", "dV33"]


def test_av_clean_step_vs_reference_golden(sx):
    """avClean = true (HydroVeProp<true>, `--avclean`): the velocity-gradient fields dV11..dV33 of the divv/curlv pass
    (divv_curlv_kern.hpp:114-122) and computeMomentumEnergy<true> with avRvCorrection (momentum_energy_kern.hpp:43-63)"""
    ref = load_golden("turb12av_step0.npz")
    plain = load_golden("turb12_step0.npz")
    got, hd = run_step_by_loops(sx, ref)
    assert hd.p.avClean == 1
    check_against_reference(got, ref)
    dv = {k: hd.get(k) for k in DV_FIELDS}
    scale = max(float(np.abs(ref[k]).max()) for k in DV_FIELDS)
    for k in DV_FIELDS:
        err = np.abs(dv[k].astype(np.float64) - ref[k]) / np.maximum(np.abs(ref[k]), 1e-2 * scale)
        assert err.max() <= REL_TOL_F32, (k, err.max())
    # the correction really changes the accelerations (the golden pair differs), and we follow the right one
    assert np.abs(ref["ax"] - plain["ax"]).max() > 1e-3 * np.abs(plain["ax"]).max()
    hd2 = sx.sim.from_dump(ref)
    calls = []
    hd2.hydro_step(halo=lambda arrs: calls.append(len(arrs)) or 0)
    assert calls == [1, 6, 7, 7]  # with avClean the last exchange carries dV11..dV33 + alpha (ve_hydro.hpp:180-184)
    for k in sx.sim.STEP_OUTPUTS:
        np.testing.assert_array_equal(hd2.get(k), got[k], err_msg=k)


@pytest.mark.parametrize("fname", ["turb12_step0.npz", "noh14_step0.npz"])
def test_fused_step_equals_loop_calls(sx, fname):
    ref = load_golden(fname)
    a, _ = run_step_by_loops(sx, ref)
    hd = sx.sim.from_dump(ref)
    calls = []
    hd.hydro_step(halo=lambda arrs: calls.append(len(arrs)) or 0)
    assert calls == [1, 6, 7, 1]  # the four halo exchanges of ve_hydro.hpp:154,165,174,185
    for k in sx.sim.STEP_OUTPUTS:
        np.testing.assert_array_equal(hd.get(k), a[k], err_msg=k)


@pytest.mark.parametrize("fname", STEP_FILES)
def test_step_vs_oracle_restatement(sx, oracle, fname):
    """same inputs through the CPU restatement (bit-identical to the reference, tests/test_oracle.py)"""
    d = load_golden(fname)
    ref = oracle.hydro_step_f(d)
    ngmax = int(d["ngmax"][0])
    ref["nb_offsets"], ref["nb_sorted"] = csr_sorted_neighbors(ref["neighbors"], ref["nc"], ngmax)
    ref["ngmax"] = d["ngmax"]
    got, _ = run_step_by_loops(sx, d)
    check_against_reference(got, ref)


@pytest.mark.skipif(not have_ref_harness(), reason="oracle/_ref/ref_harness not present")
@pytest.mark.parametrize("case,n,steps,hs", [("sedov", 50, 3, 1.0), ("noh", 40, 2, 1.0), ("turb", 32, 2, 1.4),
                                             ("noh", 30, 1, 1.7)])
def test_step_vs_compiled_reference(sx, tmp_path, case, n, steps, hs):
    """the reference itself (BASELINE config 0 is `sedov -n 50`), every dumped step"""
    dumps = run_ref_harness(case, n, steps, tmp_path / "o", hscale=hs)
    for d in dumps:
        ngmax = int(d["ngmax"][0])
        d["nb_offsets"], d["nb_sorted"] = csr_sorted_neighbors(d["neighbors"], d["nc"], ngmax)
        got, _ = run_step_by_loops(sx, d)
        check_against_reference(got, d)


@pytest.mark.skipif(not have_ref_harness(), reason="oracle/_ref/ref_harness not present")
@pytest.mark.parametrize("case,n,steps", [("sedov", 100, 2), ("noh", 80, 1), ("turb", 64, 2)])
def test_step_vs_compiled_reference_at_scale(sx, oracle, tmp_path, case, n, steps):
    """the reference itself on 0.26 - 1 M particles: every target block runs in SHIFT mode (candidates moved to the
    periodic image next to the block, the path the benchmark times; the 12^3 / 14^3 fixtures run in fold mode), and for
    Sedov and turbulence the compared step is the one after the first integrate (v != 0, x off the lattice)"""
    d = run_ref_harness(case, n, steps, tmp_path / "o", dump_every=1000)[-1]
    assert d["_step"] == steps - 1
    ngmax = int(d["ngmax"][0])
    d["nb_offsets"], d["nb_sorted"] = csr_sorted_neighbors(d["neighbors"], d["nc"], ngmax)
    got, hd = run_step_by_loops(sx, d)
    assert int((hd.block_stats()["flags"] & 1).sum()) == 0  # no fold-mode block
    check_against_reference(got, d, oracle)


@pytest.mark.skipif(not have_ref_harness(), reason="oracle/_ref/ref_harness not present")
def test_evolved_state_vs_compiled_reference(sx, tmp_path):
    """BASELINE config 0 (sedov -n 50) after 100 steps of the REFERENCE: the blast wave has left the lattice behind
    (neighbour counts vary, density contrasts at the shock, v of order one); the 101st hydro step field by field"""
    d = run_ref_harness("sedov", 50, 101, tmp_path / "o", dump_every=100)[-1]
    assert d["_step"] == 100
    assert d["nc"].min() < 93 < d["nc"].max() and np.abs(d["vx"]).max() > 0.1
    ngmax = int(d["ngmax"][0])
    d["nb_offsets"], d["nb_sorted"] = csr_sorted_neighbors(d["neighbors"], d["nc"], ngmax)
    got, _ = run_step_by_loops(sx, d)
    check_against_reference(got, d)


@pytest.mark.skipif(not have_ref_harness(), reason="oracle/_ref/ref_harness not present")
@pytest.mark.parametrize("case,n,steps,hs", [("sedov", 64, 1, 1.0), ("turb", 32, 2, 1.4), ("noh", 40, 2, 1.0)])
def test_floors_are_the_references_own_fp32_noise(sx, oracle, tmp_path, case, n, steps, hs):
    """FLOOR_FRACTION of du / ax / ay / az is not a fudge: the all-double instantiation of the momentum loop on the
    reference's own inputs (every operation in fp64) is the yardstick; the reference's production fp32 evaluation
    misses it by `noise` (a fixed fraction of the field scale), the CUDA path by no more than 3 x that, and the absolute
    error the floor admits (1e-4 x 1e-1 x scale) is within [2, 30] x the reference's own noise."""
    d = run_ref_harness(case, n, steps, tmp_path / "o", dump_every=1000, hscale=hs)[-1]
    exact = oracle.momentum_fields_d(d)
    hd = sx.sim.from_dump(d)
    hd.hydro_step()
    fam = max(np.abs(exact[k]).max() for k in ("ax", "ay", "az"))
    for k in ("ax", "ay", "az", "du"):
        scale = np.abs(exact[k]).max() if k == "du" else fam
        if scale == 0.0:
            continue  # Sedov at t = 0: no energy rate
        noise = np.abs(d[k].astype(np.float64) - exact[k]).max() / scale
        err = np.abs(hd.get(k).astype(np.float64) - exact[k]).max() / scale
        admitted = REL_TOL_F32 * FLOOR_FRACTION[k]
        assert err <= 3.0 * noise, (case, k, err, noise)
        assert 2.0 * noise <= admitted <= 30.0 * noise, (case, k, noise, admitted)


@pytest.mark.parametrize("fname,chunk", [("noh14_step0.npz", 256), ("turb12_step0.npz", 320)])
def test_candidate_chunks(sx, fname, chunk):
    """blocks whose candidate set does not fit the shared-memory buffer of a loop (evolved states reach 1500 candidates
    against 1280 - 1792 records) are processed in candidate chunks: forced here with small chunks (test hook), the step
    agrees with the reference and is bit-identical to the unchunked one"""
    d = load_golden(fname)
    ref, _ = run_step_by_loops(sx, d)
    sx.load().sphx_debug_candidate_chunk(chunk)
    try:
        got, hd = run_step_by_loops(sx, d)
    finally:
        sx.load().sphx_debug_candidate_chunk(0)
    assert (hd.block_stats()["numCand"] > chunk).any()
    check_against_reference(got, d)
    # a target's list is ascending in the candidate index, so chunking by candidate index keeps the order of every
    # partial sum: the chunked step is bit-identical to the unchunked one
    for k in sx.sim.STEP_OUTPUTS:
        np.testing.assert_array_equal(got[k], ref[k], err_msg=k)


def _random_points(n, box, seed, gaussian):
    rng = np.random.default_rng(seed)
    lo, hi = np.array(box[0::2]), np.array(box[1::2])
    p = rng.normal(0.5, 0.15, size=(n, 3)).clip(0.0, 1.0 - 1e-12) if gaussian else rng.random((n, 3))
    return (lo + p * (hi - lo)).T.copy()


@pytest.mark.parametrize("boundary", [0, 1])
@pytest.mark.parametrize("box", [[0., 1., 0., 1., 0., 1.], [-1.2, 0.23, -0.213, 3.213, -5.1, 1.23]])
@pytest.mark.parametrize("radius,n,gaussian", [(0.124, 2500, False), (0.0624, 2500, True), (3.0, 500, False)])
def test_find_neighbors_equals_all_to_all(sx, oracle, radius, n, gaussian, box, boundary):
    """domain/test/unit/neighbors/findneighbors.cpp:43-133 with the GPU search (cstone::findNeighbors call shape)."""
    import ctypes as C
    x, y, z = _random_points(n, box, seed=n + boundary, gaussian=gaussian)
    t = sx.host.build_tree(x, y, z, box, [boundary] * 3, bucket_size=64)
    x, y, z = x[t.order], y[t.order], z[t.order]
    h = np.full(n, radius / 2, np.float32)
    ngmax = n
    nb, cnt = sx.sim.find_neighbors(x, y, z, h, sx.sim.DeviceTree(t, "cuda:0"), box, [boundary] * 3, ngmax)
    L, P = oracle.lib(), oracle.P
    obox = oracle.make_box(box, [boundary] * 3)
    nb_a = np.zeros(n * ngmax, np.uint32)
    nc_a = np.zeros(n, np.uint32)
    L.orc_all2all_neighbors_f(P(x), P(y), P(z), P(h), C.c_uint(n), P(nb_a), P(nc_a), C.c_uint(ngmax), C.byref(obox))
    np.testing.assert_array_equal(cnt, nc_a)
    _, ia = csr_sorted_neighbors(nb_a, nc_a + 1, ngmax)
    _, ig = csr_sorted_neighbors(nb, cnt + 1, ngmax)
    np.testing.assert_array_equal(ig, ia)


def _block_search_vs_all_to_all(sx, oracle, x, y, z, h, box, boundary, ngmax=384, bucket=64):
    """block search of the hydro step (sphx_find_neighbors_sph, no h-iteration: ngmin = 1) vs the O(N^2) reference"""
    import ctypes as C
    from sphexa_b200.sim import HydroData, Params
    n = x.size
    t = sx.host.build_tree(x, y, z, box, boundary, bucket_size=bucket)
    o = t.order
    x, y, z, h = x[o], y[o], z[o], h[o]
    p = Params(ng0=4, ngmax=ngmax)
    hd = HydroData(n, 0, n, box, boundary, p, device="cuda:0")
    hd.set_fields(x=x, y=y, z=z, h=h, m=np.full(n, 1.0 / n, np.float32))
    hd.set_tree(t)
    hd.find_neighbors_sph()
    L, P = oracle.lib(), oracle.P
    nb_a = np.zeros(n * ngmax, np.uint32)
    nc_a = np.zeros(n, np.uint32)
    obox = oracle.make_box(box, boundary)
    L.orc_all2all_neighbors_f(P(x), P(y), P(z), P(h), C.c_uint(n), P(nb_a), P(nc_a), C.c_uint(ngmax), C.byref(obox))
    assert nc_a.max() <= ngmax, "test set-up: more neighbours than the list holds"
    np.testing.assert_array_equal(hd.get("h"), h)
    np.testing.assert_array_equal(hd.get("nc"), nc_a + 1)
    _, ia = csr_sorted_neighbors(nb_a, nc_a + 1, ngmax)
    _, ig = csr_sorted_neighbors(hd.export_neighbors(), hd.get("nc"), ngmax)
    np.testing.assert_array_equal(ig, ia)
    return t, hd


@pytest.mark.parametrize("kind", ["uniform_pbc", "clustered_open", "mixed_h_pbc", "small_buckets"])
def test_block_search_equals_all_to_all(sx, oracle, kind):
    """the block search on irregular particle sets (quad cull, interleaved tiles, precise walk): bit-exact sets"""
    rng = np.random.default_rng(11)
    n = 6000
    if kind == "clustered_open":
        pts = np.concatenate([rng.normal(0.3, 0.05, (n // 2, 3)), rng.normal(0.7, 0.12, (n // 2, 3))]).clip(0.0, 1.0 - 1e-9)
        box, boundary, h = [0., 1.] * 3, [0, 0, 0], np.full(n, 0.018, np.float32)
    elif kind == "mixed_h_pbc":
        pts = rng.random((n, 3))
        box, boundary = [0., 1.] * 3, [1, 0, 1]
        h = (0.06 * (0.5 + rng.random(n))).astype(np.float32)  # radii differ by 3x between neighbours
    else:
        pts = rng.random((n, 3)) * np.array([1.43, 3.4, 6.33]) + np.array([-1.2, -0.2, -5.1])
        box, boundary = [-1.2, 0.23, -0.2, 3.2, -5.1, 1.23], [1, 1, 1]
        # 2h = 0.5 against a box length of 1.43: the blocks run in fold mode; bucket 16: many small leaves, ~15 neighbours
        h = np.full(n, 0.25 if kind == "uniform_pbc" else 0.13, np.float32)
    x, y, z = (np.ascontiguousarray(pts[:, d]) for d in range(3))
    _block_search_vs_all_to_all(sx, oracle, x, y, z, h, box, boundary, bucket=16 if kind == "small_buckets" else 64)


def test_block_search_overflow_blocks_use_the_big_tables(sx, oracle):
    """an octree far too fine for the search radii (bucket 8 in an anisotropic box: 1.8 particles per leaf, > 800 leaves
    in reach of one target block) exceeds the standard per-block tables (512 leaves): those blocks go on the overflow
    list and are redone by the big instantiation of the search; the result is still the all-to-all one"""
    rng = np.random.default_rng(11)
    pts = rng.random((6000, 3)) * np.array([1.43, 3.4, 6.33]) + np.array([-1.2, -0.2, -5.1])
    x, y, z = (np.ascontiguousarray(pts[:, d]) for d in range(3))
    _, hd = _block_search_vs_all_to_all(sx, oracle, x, y, z, np.full(6000, 0.13, np.float32),
                                        [-1.2, 0.23, -0.2, 3.2, -5.1, 1.23], [1, 1, 1], bucket=8)
    bs = hd.block_stats()
    assert bs["numLeaves"].max() > 512 and bs["errFlags"] == 0


def test_block_search_with_coincident_particles(sx, oracle):
    """two clumps of 200 coincident particles each inside ONE cell of the deepest tree level: the leaf cannot be split
    (400 particles > half a tile), which takes the serial tile numbering of the search; pairs at distance 0 count"""
    rng = np.random.default_rng(3)
    n_bg = 3000
    pts = rng.random((2 * n_bg, 3))
    pts = pts[np.linalg.norm(pts - 0.5, axis=1) > 0.15][:n_bg]  # the background does not see the clumps
    n_bg = pts.shape[0]
    c = np.array([0.5 + 1e-7, 0.5 + 1e-7, 0.5 + 1e-7])
    clumps = np.concatenate([np.tile(c, (200, 1)), np.tile(c + 1.5e-7, (200, 1))])
    pts = np.concatenate([pts, clumps])
    h = np.concatenate([np.full(n_bg, 0.06, np.float32), np.full(400, 2e-8, np.float32)])
    x, y, z = (np.ascontiguousarray(pts[:, d]) for d in range(3))
    t, hd = _block_search_vs_all_to_all(sx, oracle, x, y, z, h, [0., 1.] * 3, [0, 0, 0])
    counts = np.diff(np.asarray(t.layout))
    assert counts.max() >= 400
    nc = hd.get("nc")
    assert (nc == 200).sum() == 400  # 199 coincident partners + self


def test_eos_prefers_u_over_temp(sx):
    """a dataset holding BOTH u and temp takes pressure and sound speed from u (hydro_ve/eos.hpp:71 `d.u.empty()`,
    eos_gpu.cu:55 `u == nullptr`); temp is only used when u is absent"""
    import torch
    d = load_golden("turb12_step0.npz")
    got, hd = run_step_by_loops(sx, d)
    prho_t, c_t = hd.get("prho").copy(), hd.get("c").copy()
    cv = np.float32(np.float64(np.float32(8.317e7) / np.float32(hd.p.muiConst)) / (hd.p.gamma - 1.0))
    u = np.float64(cv) * hd.get("temp")
    hd.f["u"] = torch.from_numpy(2.0 * u).to(hd.device)           # u says "twice as hot" as temp
    hd.eos()
    np.testing.assert_allclose(hd.get("prho"), 2.0 * prho_t, rtol=2e-7)
    np.testing.assert_allclose(hd.get("c"), np.sqrt(2.0) * c_t, rtol=2e-7)
    hd.f["u"] = torch.from_numpy(u).to(hd.device)
    hd.eos()
    np.testing.assert_allclose(hd.get("prho"), prho_t, rtol=2e-7)
    del hd.f["u"]
    hd.eos()
    np.testing.assert_array_equal(hd.get("prho"), prho_t)
    np.testing.assert_array_equal(hd.get("c"), c_t)


def test_edge_cases(sx):
    """empty range, ragged last group, ngmax truncation semantics, error codes"""
    d = load_golden("turb12_step0.npz")
    n = int(d["n"][0])
    t = sx.sim.DeviceTree(d, "cuda:0")
    # empty range
    nb, cnt = sx.sim.find_neighbors(d["x"], d["y"], d["z"], d["h"], t, d["box"], d["boundary"], 150, first=5, last=5)
    assert nb.size == 0 and cnt.size == 0
    # ragged sub-range not aligned to 32: same counts as the full search
    _, cnt_full = sx.sim.find_neighbors(d["x"], d["y"], d["z"], d["h"], t, d["box"], d["boundary"], 150)
    _, cnt_sub = sx.sim.find_neighbors(d["x"], d["y"], d["z"], d["h"], t, d["box"], d["boundary"], 150, first=7,
                                       last=n - 13)
    np.testing.assert_array_equal(cnt_sub, cnt_full[7:n - 13])
    np.testing.assert_array_equal(cnt_full + 1, d["nc"])
    # ngmax smaller than the count: count keeps counting, list is truncated (findneighbors.hpp:119-121)
    nb8, cnt8 = sx.sim.find_neighbors(d["x"], d["y"], d["z"], d["h"], t, d["box"], d["boundary"], 8)
    np.testing.assert_array_equal(cnt8, cnt_full)
    # ng0 > ngmax -> error as in sph/find_neighbors.hpp:52
    hd = sx.sim.from_dump(d)
    hd.p.ng0, hd.p.ngmax = 200, 150
    with pytest.raises(sx.SphxError) as e:
        hd.find_neighbors_xmass()
    assert e.value.code == 3
    # workspace too small
    hd = sx.sim.from_dump(d)
    hd.workspace = hd.workspace[:1024]
    with pytest.raises(sx.SphxError) as e:
        hd.find_neighbors_xmass()
    assert e.value.code == 4


def test_non_compact_blocks_use_the_precise_walk(sx):
    """Noh sphere, octree of bucket 16: for 36 of the 263 target blocks the SFC leaves the sphere and re-enters it
    elsewhere, their bounding boxes overlap up to 636 leaves (> the 512 the shared-memory tables hold); the search
    must fall back to the per-sphere node test (csrc/search.cu, "precise") instead of reporting SPHX_ERR_TRAVERSAL,
    and the result must not depend on the tree: identical nc, h, neighbour sets as with the bucket-64 tree."""
    from sphexa_b200 import cases
    a = cases.make_noh(sx, 40)
    b = cases.make_noh(sx, 40, bucket_size=16)
    assert b.tree.num_leaves > 2 * a.tree.num_leaves
    a.hydro_step()
    b.hydro_step()
    np.testing.assert_array_equal(a.get("nc"), b.get("nc"))
    np.testing.assert_array_equal(a.get("h"), b.get("h"))
    ngmax = a.p.ngmax
    na = csr_sorted_neighbors(a.export_neighbors(), a.get("nc"), ngmax)
    nb = csr_sorted_neighbors(b.export_neighbors(), b.get("nc"), ngmax)
    np.testing.assert_array_equal(na[0], nb[0])
    np.testing.assert_array_equal(na[1], nb[1])
    assert_fields_close({k: b.get(k) for k in F32_FIELDS}, {k: a.get(k) for k in F32_FIELDS})


def test_properties_at_scale(sx):
    """Sedov 100^3 lattice (1M particles), no oracle needed: every particle of the periodic lattice has exactly 92
    neighbours + self, neighbour relation is symmetric for equal h, the lattice at rest has zero
    divv/curlv and uniform kx, and total energy rate sum(m du) vanishes."""
    from sphexa_b200 import cases
    n = 100
    hd = cases.make_sedov(sx, n)
    r = hd.hydro_step()
    N = n ** 3
    nc = hd.get("nc")
    assert nc.min() == 93 and nc.max() == 93
    assert r.totalNeighbors == 93 * N
    kx = hd.get("kx")
    assert np.abs(kx / kx.mean() - 1).max() < 1e-5
    assert np.abs(hd.get("divv")).max() == 0.0 or np.abs(hd.get("divv")).max() < 1e-3
    # symmetry of the neighbour relation on a sample of particles
    nb = hd.export_neighbors().reshape(N, hd.p.ngmax)[:, :92]
    rng = np.random.default_rng(0)
    for i in rng.integers(0, N, 200):
        for j in nb[i, :8]:
            assert i in nb[j]
    # momentum conservation: sum of m*a vanishes to rounding
    ax = hd.get("ax").astype(np.float64)
    assert abs(ax.sum()) < 1e-6 * np.abs(ax).sum() + 1e-30
