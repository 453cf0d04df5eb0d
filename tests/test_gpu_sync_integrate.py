"""GPU tests of the callers of the hot path (SURVEY §8f "next" rows), all through the C ABI:
  rank 1  sphx_domain_sync / sphx_reorder_fields   vs the host builder (csrc/host_domain.cpp, itself pinned to the
          reference's Domain dumps in tests/test_host_tree.py) and vs the trees inside the golden dumps
  rank 2  sphx_compute_timestep / sphx_compute_positions / sphx_update_smoothing_length / sphx_integrate
          vs the post-integrate state of the reference harness dumps: BIT-EXACT when fed the reference's accelerations
  rank 3  sphx_conserved_quantities vs a numpy fp64 evaluation of the reference formulas
  loop    Simulation.step() over 100 steps of Sedov 50^3 (BASELINE config 0) vs the reference CPU energy series
"""
import ctypes as C

import numpy as np
import pytest

from refdata import GOLDEN, load_golden

pytestmark = pytest.mark.gpu

STEP_FILES = ["sedov12_step0.npz", "sedov12_step2.npz", "noh14_step0.npz", "turb12_step0.npz", "turb12h_step0.npz"]
TREE_KEYS = ["tree_prefixes", "tree_childOffsets", "tree_levelRange", "tree_leaves", "tree_layout", "tree_centers",
             "tree_sizes"]


@pytest.fixture(scope="module")
def sx():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device: the product has no CPU fallback")
    import sphexa_b200
    sphexa_b200.load()
    return sphexa_b200


def device_sync(sx, x, y, z, box, boundary, bucket=64, update_box=False):
    import torch
    from sphexa_b200 import _cabi, host
    from sphexa_b200.sim import DeviceTree
    L = _cabi.load()
    dev = torch.device("cuda:0")
    n = x.size
    xd, yd, zd = (torch.from_numpy(np.ascontiguousarray(a, np.float64)).to(dev) for a in (x, y, z))
    max_nodes = max(64, n)
    t = DeviceTree.empty(max_nodes, dev)
    keys = torch.zeros(n, dtype=torch.int64, device=dev)
    order = torch.zeros(n, dtype=torch.int32, device=dev)
    scratch = torch.empty(L.sphx_domain_sync_bytes(n, max_nodes), dtype=torch.uint8, device=dev)
    a = _cabi.SphxSyncArgs()
    a.n, a.box, a.bucketSize = n, host.make_box(box, boundary), bucket
    a.x, a.y, a.z = xd.data_ptr(), yd.data_ptr(), zd.data_ptr()
    a.keys, a.order, a.maxNodes = keys.data_ptr(), order.data_ptr(), max_nodes
    for k in ("prefixes", "childOffsets", "internalToLeaf", "levelRange", "leaves", "layout", "centers", "sizes"):
        setattr(a, k, getattr(t, k).data_ptr())
    a.scratch, a.scratchBytes = scratch.data_ptr(), scratch.numel()
    nn, nl, bo = C.c_int(0), C.c_int(0), _cabi.SphxBox()
    _cabi.check(L.sphx_domain_sync(C.byref(a), C.byref(bo) if update_box else None, C.byref(nn), C.byref(nl)))
    t.num_nodes, t.num_leaves = nn.value, nl.value
    out = t.to_host()
    out["keys"] = keys.cpu().numpy().view(np.uint64)
    out["order"] = order.cpu().numpy().view(np.uint32)
    out["box"] = np.array(list(bo.lim)) if update_box else np.array(box, np.float64)
    return out


def test_domain_sync_shrink_limit_and_empty_rank(sx):
    """open box: SPHX_SYNC_LIMIT_SHRINK keeps each side within 5 % of the previous extent (limitBoxShrinking,
    sfc/box.hpp:397-414; domain/assignment.hpp:80-82); n == 0 is a no-op with the empty root leaf as tree"""
    import torch
    from sphexa_b200 import _cabi, host
    from sphexa_b200.sim import DeviceTree
    L = _cabi.load()
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(5)
    n = 4000
    pts = 0.3 + 0.4 * rng.random((3, n))
    pts[0] = -0.2 + 1.4 * rng.random(n)  # x leaves the previous box on both sides: the box grows freely
    xd, yd, zd = (torch.from_numpy(pts[d].copy()).to(dev) for d in range(3))
    t = DeviceTree.empty(n, dev)
    keys = torch.zeros(n, dtype=torch.int64, device=dev)
    order = torch.zeros(n, dtype=torch.int32, device=dev)
    scratch = torch.empty(L.sphx_domain_sync_bytes(n, n), dtype=torch.uint8, device=dev)

    def run(count, flags, boundary):
        a = _cabi.SphxSyncArgs()
        a.n, a.box, a.bucketSize = count, host.make_box([0., 1.] * 3, boundary), 64
        a.x, a.y, a.z = xd.data_ptr(), yd.data_ptr(), zd.data_ptr()
        a.keys, a.order, a.maxNodes = keys.data_ptr(), order.data_ptr(), n
        for k in ("prefixes", "childOffsets", "internalToLeaf", "levelRange", "leaves", "layout", "centers", "sizes"):
            setattr(a, k, getattr(t, k).data_ptr())
        a.scratch, a.scratchBytes, a.flags = scratch.data_ptr(), scratch.numel(), flags
        nn, nl, bo = C.c_int(0), C.c_int(0), _cabi.SphxBox()
        _cabi.check(L.sphx_domain_sync(C.byref(a), C.byref(bo), C.byref(nn), C.byref(nl)))
        return list(bo.lim), nn.value, nl.value

    fit = [pts[0].min(), pts[0].max(), pts[1].min(), pts[1].max(), pts[2].min(), pts[2].max()]
    box, _, _ = run(n, 0, [0, 0, 1])
    assert box == [fit[0], fit[1], fit[2], fit[3], 0.0, 1.0]          # first call: the fitting box; z is periodic
    box, _, _ = run(n, _cabi.SPHX_SYNC_LIMIT_SHRINK, [0, 0, 1])
    assert box == [fit[0], fit[1], 0.05, 0.95, 0.0, 1.0]              # y may shrink by 5 % per side only
    box, nn, nl = run(0, 0, [0, 0, 0])
    assert (nn, nl) == (1, 1) and box == [0., 1.] * 3
    assert t.childOffsets[0].item() == 0 and t.layout[:2].tolist() == [0, 0]


def assert_tree_equal(got: dict, ht):
    exp = ht.as_dump_dict()
    assert got["tree_childOffsets"].size == exp["tree_childOffsets"].size
    for k in TREE_KEYS:
        assert np.array_equal(got[k], exp[k]), k
    leaf = exp["tree_childOffsets"] == 0
    assert np.array_equal(got["tree_internalToLeaf"][leaf], exp["tree_internalToLeaf"][leaf])


@pytest.mark.parametrize("case", ["lattice20", "uniform", "clustered", "tiny", "duplicates", "open_clustered"])
def test_device_sync_equals_host_builder(sx, case):
    rng = np.random.default_rng(7)
    boundary = [1, 1, 1]
    box = [-0.5, 0.5] * 3
    bucket = 64
    if case == "lattice20":
        from sphexa_b200.cases import regular_grid
        x, y, z = regular_grid(0.5, 20)
    elif case == "uniform":
        x, y, z = rng.uniform(-0.5, 0.5, (3, 50000))
    elif case == "clustered":
        p = np.concatenate([rng.normal(0.1, 0.01, (3, 30000)), rng.uniform(-0.5, 0.5, (3, 5000))], axis=1)
        x, y, z = np.clip(p, -0.5, 0.5 - 1e-12)
        bucket = 16
    elif case == "tiny":
        x, y, z = rng.uniform(-0.5, 0.5, (3, 40))  # root stays a leaf
    elif case == "duplicates":
        x, y, z = rng.uniform(-0.5, 0.5, (3, 3000))
        x[:200], y[:200], z[:200] = 0.123, -0.2, 0.3  # 200 coincident particles: the tree descends to level 21
        bucket = 8
    else:
        p = np.concatenate([rng.normal(0.0, 0.05, (3, 20000)), rng.uniform(-2.0, 3.0, (3, 2000))], axis=1)
        x, y, z = p
        boundary = [0, 0, 0]
        box = [x.min(), x.max(), y.min(), y.max(), z.min(), z.max()]
    from sphexa_b200 import host
    ht = host.build_tree(x, y, z, box, boundary, bucket)
    got = device_sync(sx, x, y, z, box, boundary, bucket)
    assert np.array_equal(got["keys"], ht.keys)
    assert np.array_equal(got["order"], ht.order)  # both sorts are stable
    assert_tree_equal(got, ht)
    assert np.all(np.diff(got["keys"].astype(np.uint64)) >= 0)
    if case == "open_clustered":  # makeGlobalBox: extrema of the non-periodic dimensions
        got2 = device_sync(sx, x, y, z, [0, 1] * 3, boundary, bucket, update_box=True)
        assert np.array_equal(got2["box"], np.array(box))
        assert np.array_equal(got2["keys"], ht.keys)


@pytest.mark.parametrize("fname", STEP_FILES)
def test_device_sync_reproduces_reference_domain(sx, fname):
    """particles of a reference dump, shuffled: the device sync must restore the reference's SFC order (checked through
    keys and particle ids) and, for converged reference trees, the same OctreeNsView arrays"""
    d = load_golden(fname)
    n = int(d["n"][0])
    perm = np.random.default_rng(3).permutation(n)
    got = device_sync(sx, d["x"][perm], d["y"][perm], d["z"][perm], d["box"], d["boundary"], 64)
    assert np.array_equal(got["keys"], d["keys"])
    assert np.array_equal(d["id"][perm][got["order"]], d["id"]) or np.unique(d["keys"]).size < n
    if got["tree_childOffsets"].size == d["tree_childOffsets"].size:
        for k in TREE_KEYS:
            assert np.array_equal(got[k], d[k]), k


def test_reorder_fields(sx):
    import torch
    from sphexa_b200 import _cabi
    L = _cabi.load()
    dev = torch.device("cuda:0")
    n = 100003
    g = torch.Generator(device="cpu").manual_seed(1)
    order = torch.randperm(n, generator=g).to(torch.int32).to(dev)
    srcs = [torch.rand(n, dtype=torch.float64, generator=g).to(dev), torch.rand(n, generator=g).to(dev),
            torch.arange(n, dtype=torch.int64).to(dev), torch.arange(n, dtype=torch.int16).to(dev),
            (torch.arange(n) % 251).to(torch.uint8).to(dev)]
    dsts = [torch.empty_like(s) for s in srcs]
    k = len(srcs)
    _cabi.check(L.sphx_reorder_fields(order.data_ptr(), n, k, (C.c_void_p * k)(*[s.data_ptr() for s in srcs]),
                                      (C.c_void_p * k)(*[t.data_ptr() for t in dsts]),
                                      (C.c_int * k)(*[s.element_size() for s in srcs]), None))
    torch.cuda.synchronize()
    for s, t in zip(srcs, dsts):
        assert torch.equal(t, s[order.long()])
    # aliased arrays are refused
    rc = L.sphx_reorder_fields(order.data_ptr(), n, 1, (C.c_void_p * 1)(srcs[0].data_ptr()),
                               (C.c_void_p * 1)(srcs[0].data_ptr()), (C.c_int * 1)(8), None)
    assert rc == 3


def _integrate_from_dump(sx, d, mode):
    """feed the REFERENCE's step outputs (ax, ay, az, du, nc, h after the h-iteration) to the integrate kernels"""
    import torch
    from sphexa_b200 import _cabi, host
    L = _cabi.load()
    dev = torch.device("cuda:0")
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int32) if a.dtype == np.uint32
                                    else np.ascontiguousarray(a)).to(dev)  # noqa: E731
    f = {k: up(d[k]) for k in ("x", "y", "z", "x_m1", "y_m1", "z_m1", "vx", "vy", "vz", "ax", "ay", "az", "temp", "du",
                               "du_m1", "h", "nc")}
    par = d["params"]
    dt, dt1, tt = C.c_double(par[5]), C.c_double(par[6]), C.c_double(par[13])
    _cabi.check(L.sphx_compute_timestep(float(d["dts"][0]), float(d["dts"][1]), float(par[15]), C.byref(dt),
                                        C.byref(dt1), C.byref(tt), None, None))
    assert (dt.value, dt1.value, tt.value) == tuple(d["post_dt"])  # bit-exact
    a = _cabi.SphxIntegrateArgs()
    for k in _cabi.INTEGRATE_FIELDS:
        setattr(a, k, f[k].data_ptr() if k in f else None)
    n = int(d["n"][0])
    a.first, a.last, a.box = 0, n, host.make_box(d["box"], d["boundary"])
    a.dt, a.dt_m1, a.gamma, a.muiConst, a.ng0 = dt.value, dt1.value, par[3], par[4], int(d["ng0"][0])
    if mode == "fused":
        _cabi.check(L.sphx_integrate(C.byref(a)))
    else:
        _cabi.check(L.sphx_compute_positions(C.byref(a)))
        _cabi.check(L.sphx_update_smoothing_length(C.byref(a)))
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in f.items()}


@pytest.mark.parametrize("mode", ["fused", "split"])
@pytest.mark.parametrize("fname", STEP_FILES)
def test_integrate_bit_exact_vs_reference(sx, fname, mode):
    d = load_golden(fname)
    got = _integrate_from_dump(sx, d, mode)
    for k in ("x", "y", "z", "vx", "vy", "vz", "x_m1", "y_m1", "z_m1", "temp", "du_m1", "h"):
        assert np.array_equal(got[k], d["post_" + k]), k


def test_integrate_fixed_boundaries(sx):
    """particles at rest within 2h of a fixed wall keep position, velocity and x_m1 (positions.hpp:97-107); their
    energy is still advanced (updateTempHost covers the whole range)"""
    d = dict(load_golden("sedov12_step0.npz"))
    d["boundary"] = np.array([2, 1, 1], np.int32)
    free = _integrate_from_dump(sx, dict(d, boundary=np.array([1, 1, 1], np.int32)), "fused")
    got = _integrate_from_dump(sx, d, "fused")
    near = (np.abs(0.5 - d["x"]) < 2.0 * d["h"].astype(np.float64)) | (np.abs(-0.5 - d["x"]) < 2.0 * d["h"].astype(np.float64))
    atrest = (d["vx"] == 0) & (d["vy"] == 0) & (d["vz"] == 0)
    pinned = near & atrest
    assert pinned.any() and (~pinned).any()
    for k in ("x", "y", "z", "x_m1", "vx"):
        assert np.array_equal(got[k][pinned], d[k][pinned]), k
    for k in ("y", "z", "vy", "vz", "x_m1", "y_m1"):
        assert np.array_equal(got[k][~pinned], free[k][~pinned]), k
    # x of free particles: no periodic fold in a fixed dimension, otherwise identical
    assert np.array_equal(got["temp"], free["temp"]) and np.array_equal(got["h"], free["h"])


@pytest.mark.parametrize("fname", ["sedov12_step2.npz", "noh14_step0.npz", "turb12_step0.npz"])
def test_conserved_quantities(sx, fname):
    import torch
    from sphexa_b200 import _cabi
    from sphexa_b200.cases import ideal_gas_cv
    L = _cabi.load()
    d = load_golden(fname)
    dev = torch.device("cuda:0")
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int32) if a.dtype == np.uint32
                                    else np.ascontiguousarray(a)).to(dev)  # noqa: E731
    f = {k: up(d[k]) for k in ("x", "y", "z", "vx", "vy", "vz", "m", "temp", "nc")}
    scratch = torch.zeros(L.sphx_conserved_scratch_bytes(), dtype=torch.uint8, device=dev)
    out = _cabi.SphxConserved()
    n = int(d["n"][0])
    par = d["params"]
    _cabi.check(L.sphx_conserved_quantities(*[f[k].data_ptr() for k in ("x", "y", "z", "vx", "vy", "vz", "m", "temp")],
                                            None, f["nc"].data_ptr(), 0, n, par[3], par[4], 0.0, scratch.data_ptr(),
                                            None, None, C.byref(out)))
    m = d["m"].astype(np.float64)
    V = np.stack([d["vx"], d["vy"], d["vz"]]).astype(np.float64)
    X = np.stack([d["x"], d["y"], d["z"]])
    cv = np.float64(ideal_gas_cv(par[4], par[3]))
    ecin = 0.5 * np.sum(m * (V ** 2).sum(0))
    eint = np.sum(cv * d["temp"] * m)
    lin = (m * V).sum(1)
    ang = (m * np.cross(X.T, V.T).T).sum(1)
    assert out.ecin == pytest.approx(ecin, rel=1e-13, abs=1e-300)
    assert out.eint == pytest.approx(eint, rel=1e-13)
    assert out.etot == pytest.approx(ecin + eint, rel=1e-13)
    scale = np.abs(m * np.abs(V)).sum() + 1e-300
    assert np.allclose(list(out.linmom3), lin, rtol=0, atol=1e-13 * scale)
    assert np.allclose(list(out.angmom3), ang, rtol=0, atol=1e-13 * scale)
    assert out.linmom == pytest.approx(np.sqrt((np.array(list(out.linmom3)) ** 2).sum()), rel=1e-14, abs=1e-300)
    assert out.totalNeighbors == int(d["nc"].astype(np.int64).sum())
    # deterministic: the same call gives the same bits
    out2 = _cabi.SphxConserved()
    _cabi.check(L.sphx_conserved_quantities(*[f[k].data_ptr() for k in ("x", "y", "z", "vx", "vy", "vz", "m", "temp")],
                                            None, f["nc"].data_ptr(), 0, n, par[3], par[4], 0.0, scratch.data_ptr(),
                                            None, None, C.byref(out2)))
    assert bytes(out) == bytes(out2)


def test_simulation_first_steps_match_reference_dump(sx):
    """Simulation.step() from the Sedov 12^3 initial state: after 2 native steps (sync, forces, integrate each) the
    state entering step 2 must match the reference harness dump of step 2: same SFC order (ids), nc exact, positions
    and fields to fp32 accumulation noise"""
    from sphexa_b200 import cases
    d2 = load_golden("sedov12_step2.npz")
    s = cases.make_sedov_sim(sx, 12)
    rows = [s.step() for _ in range(2)]
    s.sync()
    assert np.array_equal(s.get("id").view(np.uint64), d2["id"])
    assert s.p.minDt == d2["params"][5] and s.p.minDt_m1 == d2["params"][6] and s.p.ttot == d2["params"][13]
    for k, tol in (("x", 1e-12), ("y", 1e-12), ("z", 1e-12)):
        assert np.abs(s.get(k) - d2[k]).max() <= tol, k
    np.testing.assert_allclose(s.get("h"), d2["h_in"], rtol=0, atol=0)
    np.testing.assert_allclose(s.get("temp"), d2["temp"], rtol=1e-6)
    s.compute_forces()
    assert np.array_equal(s.get("nc"), d2["nc"])
    assert rows[0][3] == pytest.approx(1.0, rel=1e-3)


def test_simulation_sedov50_energy_100_steps(sx):
    """BASELINE.json: "Sedov energy conservation must match over 100 steps": BASELINE config 0 (sedov -n 50) run entirely
    through libsphx (device sync, hydro step, conserved quantities, integrate) against the series of the reference CPU
    code (tests/golden/sedov50_energies.npz, oracle/_ref/ref_harness)."""
    from sphexa_b200 import cases
    with np.load(GOLDEN / "sedov50_energies.npz") as z:
        ref = z["series"]  # step ttot minDt etot ecin eint linmom angmom totalNeighbors
    steps = ref.shape[0]
    s = cases.make_sedov_sim(sx, 50)
    got = np.array([s.step() for _ in range(steps)], dtype=np.float64)
    # time-step sequence and energies: fp32 pair sums in a different order => relative 1e-6 class differences
    np.testing.assert_allclose(got[:, 2], ref[:, 2], rtol=2e-5)            # minDt
    np.testing.assert_allclose(got[:, 1], ref[:, 1], rtol=2e-5, atol=1e-15)  # ttot
    np.testing.assert_allclose(got[:, 3], ref[:, 3], rtol=1e-6)            # etot
    np.testing.assert_allclose(got[:, 5], ref[:, 5], rtol=1e-6)            # eint
    np.testing.assert_allclose(got[:, 4], ref[:, 4], rtol=2e-5, atol=1e-12)  # ecin
    assert np.array_equal(got[:5, 8], ref[:5, 8])                           # neighbour sums: exact at first
    np.testing.assert_allclose(got[:, 8], ref[:, 8], rtol=1e-3)
    # conservation itself, as the reference conserves it
    assert abs(got[-1, 3] / got[0, 3] - 1.0) < 5e-4
    assert abs((got[-1, 3] - got[0, 3]) - (ref[-1, 3] - ref[0, 3])) < 1e-5


@pytest.mark.parametrize("tag", ["noh14", "turb12"])
def test_simulation_noh_turbulence_energy_series(sx, tag):
    """the native loop continued from the reference's step-0 state for 40 steps: Noh (open box that follows the
    particles through makeGlobalBox in every sync, strong compression) and the turbulence box (periodic, v != 0)"""
    from sphexa_b200 import sim
    d = load_golden(f"{tag}_step0.npz")
    with np.load(GOLDEN / f"{tag}_energies.npz") as z:
        ref = z["series"]
    s = sim.simulation_from_dump(d)
    got = np.array([s.step() for _ in range(ref.shape[0])], dtype=np.float64)
    np.testing.assert_allclose(got[:, 2], ref[:, 2], rtol=2e-5)             # minDt
    np.testing.assert_allclose(got[:, 3], ref[:, 3], rtol=1e-6)             # etot
    np.testing.assert_allclose(got[:, 4], ref[:, 4], rtol=2e-5)             # ecin
    np.testing.assert_allclose(got[:, 5], ref[:, 5], rtol=2e-5, atol=1e-18)  # eint
    np.testing.assert_allclose(got[:, 6], ref[:, 6], rtol=1e-4)             # |linear momentum| (non-zero in both cases)
    assert np.array_equal(got[:3, 8], ref[:3, 8])
    np.testing.assert_allclose(got[:, 8], ref[:, 8], rtol=2e-3)


@pytest.mark.parametrize("level,boundary,nranks", [(3, [1, 1, 1], 2), (3, [0, 0, 0], 3), (4, [1, 0, 1], 8),
                                                   (4, [1, 1, 1], 5), (2, [1, 1, 1], 4), (5, [0, 1, 0], 7),
                                                   (3, [1, 1, 1], 1)])
def test_device_cell_plan_equals_host_plan(sx, level, boundary, nranks):
    """multi-rank Domain::sync, the decomposition plan: sphx_cell_plan_build_device (scans + one thread per cell) gives
    exactly what the host builder derives from the same global histogram (tests/test_dist_plan.py pins that one against
    brute force): assignment, halo cells, receive ranges, send index lists, layout sizes; plus the migration offsets"""
    import torch
    from sphexa_b200 import dist as sdist
    rng = np.random.default_rng(100 * level + nranks)
    ncell = 8 ** level
    G = rng.integers(0, 40, ncell).astype(np.uint32)
    G[rng.random(ncell) < 0.3] = 0            # empty cells
    if level >= 3:
        G[ncell // 3: ncell // 3 + ncell // 6] = 0  # an empty stretch of the curve: ranks with few or no cells near it
    Lc = (G * rng.random(ncell)).astype(np.uint32)  # this rank's share of every cell before the migration
    dev = torch.device("cuda:0")
    Gd = torch.from_numpy(G.view(np.int32)).to(dev)
    Ld = torch.from_numpy(Lc.view(np.int32)).to(dev)
    scratch = None
    # per-cell reach: one ring everywhere (None) or 1 - 3 rings (cells with large smoothing lengths reach further)
    rings = None if (level + nranks) % 2 else rng.choice(np.array([1, 1, 1, 2, 3], np.uint8), ncell)
    rings_d = None if rings is None else torch.from_numpy(rings).to(dev)
    for rank in range(nranks):
        ref = sdist.cell_plan(G, level, boundary, rank, nranks, rings=rings)
        got, send_idx, scratch = sdist.cell_plan_device(Gd, Ld, level, boundary, rank, nranks, scratch=scratch,
                                                        send_capacity=ref.send_idx.size + 8, want_recv_cells=True,
                                                        rings=rings_d, max_ring=3)
        np.testing.assert_array_equal(got.cell_splits, ref.cell_splits)
        assert (got.n_assigned, got.n_halo_left, got.n_halo_right, got.n_global) == \
               (ref.n_assigned, ref.n_halo_left, ref.n_halo_right, ref.n_global)
        np.testing.assert_array_equal(got.peers, ref.peers)
        np.testing.assert_array_equal(got.send_offsets, ref.send_offsets)
        np.testing.assert_array_equal(got.recv_begin, ref.recv_begin)
        np.testing.assert_array_equal(got.recv_count, ref.recv_count)
        np.testing.assert_array_equal(got.recv_cells, ref.recv_cells)
        assert got.num_send == ref.send_idx.size
        np.testing.assert_array_equal(send_idx[:got.num_send].cpu().numpy().view(np.uint32), ref.send_idx)
        prefix_l = np.concatenate([[0], np.cumsum(Lc.astype(np.int64))])
        np.testing.assert_array_equal(got.send_off_local, prefix_l[ref.cell_splits.astype(np.int64)])
    # a send list that does not fit is an error, not a truncation
    ref = sdist.cell_plan(G, level, boundary, 0, nranks, rings=rings)
    if ref.send_idx.size > 4:
        with pytest.raises(sx.SphxError) as e:
            sdist.cell_plan_device(Gd, Ld, level, boundary, 0, nranks, scratch=scratch, send_capacity=ref.send_idx.size - 3,
                                   rings=rings_d, max_ring=3)
        assert e.value.code == 4
