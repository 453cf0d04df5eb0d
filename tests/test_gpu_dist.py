"""Multi-GPU parity: N ranks, one per GPU, SFC domain decomposition + NCCL halo exchange inside
sphx_hydro_step_dist, compared by particle id with the single-GPU result of the same C-ABI path (which itself is
parity-tested against the reference in test_gpu_parity.py). Neighbour counts must be identical; fields agree to fp32
summation-order noise (the neighbour lists of a rank are ordered by its local candidate numbering)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

OUT = ["h", "nc", "xm", "kx", "gradh", "prho", "c", "c11", "c12", "c13", "c22", "c23", "c33", "divv", "curlv", "alpha",
       "ax", "ay", "az", "du"]


def _case(name, side):
    from sphexa_b200 import cases
    if name == "onecell":
        # all particles in ONE cell of the decomposition grid: they fill one octant of the periodic box and their
        # smoothing lengths (2 h = 0.28 > 1/4) put the plan on level 1, i.e. on the 8 octants. With two ranks one of them
        # owns every particle and the other none (more ranks than occupied cells)
        from sphexa_b200.sim import Params
        rng = np.random.default_rng(3)
        n = side ** 3
        pts = 0.5 * rng.random((3, n)) - 0.5
        p = Params(minDt=1e-4, minDt_m1=1e-4, ng0=40, ngmax=150)  # 25 - 94 neighbours at h = 0.14: no h-iteration
        # (cold gas: next to the vacuum of the other seven octants any sizeable pressure would fling the particles
        # across the box within the first time step)
        f = dict(h=np.float32(0.14), m=np.float32(1.0 / n), temp=np.float64(1e-10), alpha=np.float32(0.05),
                 vx=np.zeros(n, np.float32), vy=np.zeros(n, np.float32), vz=np.zeros(n, np.float32))
        return dict(x=pts[0].copy(), y=pts[1].copy(), z=pts[2].copy(), fields=f, params=p, box=[-0.5, 0.5] * 3,
                    boundary=[1, 1, 1])
    if name == "turbstir":  # turbulence box at rest: all motion comes from the stirring (turbulence-ve propagator)
        g = cases.turbulence_global(side)
        for k in ("vx", "vy", "vz"):
            g["fields"][k] = np.zeros_like(g["fields"][k])
        return g
    return cases.sedov_global(side) if name == "sedov" else cases.noh_global(side)


def _worker(rank, world, port, name, side, q):
    import torch
    import torch.distributed as dist
    import sphexa_b200 as sx
    from sphexa_b200 import dist as sdist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dh = sdist.DistributedHydro(sx.sim, _case(name, side), rank, world, f"cuda:{rank}")
        r = dh.step()
        out = {k: dh.assigned(k) for k in OUT}
        out["id"] = dh.assigned_ids()
        out["scalars"] = np.array([r.minDtCourant, r.minDtRho, float(r.totalNeighbors), float(r.maxNc)])
        out["n_local"] = dh.hd.n
        dh.close()
        q.put((rank, out))
    except BaseException:  # noqa: BLE001
        import traceback
        q.put((rank, {"error": traceback.format_exc()}))
        raise
    finally:
        if world > 1:
            dist.destroy_process_group()


def _run(world, name, side):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() + world) % 1500
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, side, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = _collect(procs, q)
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    ids = np.concatenate([res[r]["id"] for r in range(world)])
    merged = {}
    for k in OUT:
        v = np.concatenate([res[r][k] for r in range(world)])
        full = np.zeros(ids.size, v.dtype)
        full[ids] = v
        merged[k] = full
    assert np.array_equal(np.sort(ids), np.arange(ids.size))  # every particle assigned to exactly one rank
    return (merged, [res[r]["scalars"] for r in range(world)], [res[r]["n_local"] for r in range(world)],
            [res[r]["id"] for r in range(world)])


@pytest.mark.parametrize("name,side", [("sedov", 40), ("noh", 36)])
@pytest.mark.parametrize("world", [2])
def test_multi_gpu_equals_single_gpu(world, name, side):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    from test_gpu_parity import assert_fields_close, F32_FIELDS
    ref, ref_scal, _, _ = _run(1, name, side)
    got, scal, n_local, ids = _run(world, name, side)
    np.testing.assert_array_equal(got["nc"], ref["nc"])
    np.testing.assert_array_equal(got["h"], ref["h"])
    assert_fields_close(got, ref, F32_FIELDS, tol=1e-4)
    # Reduced scalars, as the reference reduces them: each rank turns ITS max divv into Krho / |max divv|
    # (rhoTimestep, ts_global.hpp:72-95) and the time step is the MPI_Allreduce(MIN) of the per-rank values
    # (ts_global.hpp:97-113), which for an all-negative divv field is not the single-rank value.
    krho = 0.06
    dt_rho = min(krho / abs(float(got["divv"][i].max())) if got["divv"][i].max() != 0 else np.inf for i in ids)
    for s in scal:  # every rank holds the globally reduced scalars
        np.testing.assert_allclose(s[0], ref_scal[0][0], rtol=1e-5)
        np.testing.assert_allclose(s[1], dt_rho, rtol=1e-5)
        assert s[2] == ref_scal[0][2] and s[3] == ref_scal[0][3]
    assert sum(n_local) > ref["nc"].size  # halos are really there


# ---------------------------------------------------------------------------------------------------------------------
# multi-GPU time-step loop with the dynamic decomposition (DistributedSimulation.sync = multi-rank Domain::sync)
# ---------------------------------------------------------------------------------------------------------------------
SIM_FIELDS = ["x", "y", "z", "h", "vx", "vy", "vz", "temp", "alpha", "x_m1", "du_m1"]


def _sim_worker(rank, world, port, name, side, steps, q):
    import torch
    import torch.distributed as dist
    import sphexa_b200 as sx
    from sphexa_b200 import dist as sdist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ds = sdist.DistributedSimulation(sx.sim, _case(name, side), rank, world, f"cuda:{rank}")
        if name == "turbstir":
            ds.turbulence = sx.sim.Turbulence()
        rows, stats = [], []
        for _ in range(steps):
            rows.append(ds.step())
            stats.append((ds.hd.first, ds.hd.last - ds.hd.first, ds.hd.n, ds.level))
        out = {k: ds.assigned(k) for k in SIM_FIELDS + ["nc"]}
        out["id"] = ds.assigned("id")
        out["rows"] = np.array(rows, np.float64)
        out["stats"] = np.array(stats, np.int64)
        out["keys_sorted"] = bool((ds.local_keys[1:] >= ds.local_keys[:-1]).all())
        ds.close()
        q.put((rank, out))
    except BaseException:  # noqa: BLE001  (reported to the parent, which stops the other ranks: they may sit in a collective)
        import traceback
        q.put((rank, {"error": traceback.format_exc()}))
        raise
    finally:
        if world > 1:
            dist.destroy_process_group()


def _collect(procs, q, timeout=600):
    """results of all ranks; the first rank that reports an exception (or dies silently) fails the test at once instead
    of leaving the parent waiting for peers that are blocked in a collective"""
    import queue
    import time
    res, t0 = {}, time.time()
    while len(res) < len(procs):
        try:
            rank, out = q.get(timeout=2)
        except queue.Empty:
            dead = [p for p in procs if p.exitcode not in (None, 0)]
            if dead or time.time() - t0 > timeout:
                for p in procs:
                    p.kill()
                pytest.fail(f"rank process died (exit codes {[p.exitcode for p in procs]})" if dead else "ranks timed out")
            continue
        if "error" in out:
            for p in procs:
                p.kill()
            pytest.fail(f"rank {rank} raised:\n{out['error']}")
        res[rank] = out
    return res


def _run_sim(world, name, side, steps):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 27100 + (os.getpid() + 7 * world) % 1500
    procs = [ctx.Process(target=_sim_worker, args=(r, world, port, name, side, steps, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = _collect(procs, q)
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    ids = np.concatenate([res[r]["id"] for r in range(world)])
    assert np.array_equal(np.sort(ids), np.arange(ids.size))  # every particle on exactly one rank after migration
    merged = {}
    for k in SIM_FIELDS + ["nc"]:
        v = np.concatenate([res[r][k] for r in range(world)])
        full = np.zeros(ids.size, v.dtype)
        full[ids] = v
        merged[k] = full
    return merged, res


@pytest.mark.parametrize("name,side,steps", [("sedov", 30, 12), ("noh", 30, 12)])
def test_single_rank_distributed_loop_equals_simulation(name, side, steps):
    """world = 1: the dynamic-decomposition code path (histogram, plan, reorder, presorted tree) against
    sim.Simulation, which sorts and builds its tree in one sphx_domain_sync call: identical bits"""
    import sphexa_b200 as sx
    from sphexa_b200 import cases
    merged, res = _run_sim(1, name, side, steps)
    s = getattr(cases, f"make_{name}_sim")(sx, side)
    rows = np.array([s.step() for _ in range(steps)], np.float64)
    np.testing.assert_array_equal(res[0]["rows"][:, 1:3], rows[:, 1:3])   # ttot, minDt
    np.testing.assert_array_equal(res[0]["rows"][:, 8], rows[:, 8])       # total neighbours
    np.testing.assert_allclose(res[0]["rows"][:, 3:6], rows[:, 3:6], rtol=1e-12)
    assert res[0]["keys_sorted"]


@pytest.mark.parametrize("name,side,steps", [("sedov", 30, 12), ("noh", 32, 12)])
@pytest.mark.parametrize("world", [2])
def test_multi_gpu_loop_equals_single_gpu(world, name, side, steps):
    """N ranks re-decompose the domain in every step (migration + halo discovery + local tree) and advance the same
    simulation as one rank: identical neighbour counts while the trajectories are bitwise comparable, energies and
    the final per-particle state to fp32 summation-order noise."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    ref, ref_res = _run_sim(1, name, side, steps)
    got, res = _run_sim(world, name, side, steps)
    r0, g0 = ref_res[0]["rows"], res[0]["rows"]
    for r in range(1, world):
        np.testing.assert_array_equal(res[r]["rows"], g0)              # every rank holds the reduced values
    np.testing.assert_array_equal(g0[:3, 8], r0[:3, 8])                # total neighbours
    np.testing.assert_allclose(g0[:, 8], r0[:, 8], rtol=1e-3)
    np.testing.assert_allclose(g0[:, 2], r0[:, 2], rtol=1e-5)          # minDt
    np.testing.assert_allclose(g0[:, 3], r0[:, 3], rtol=1e-6)          # etot
    np.testing.assert_allclose(g0[:, 4], r0[:, 4], rtol=1e-4, atol=1e-12)  # ecin
    for k in ("x", "y", "z"):
        assert np.abs(got[k] - ref[k]).max() < 1e-7, k
    np.testing.assert_allclose(got["h"], ref["h"], rtol=1e-6)
    np.testing.assert_allclose(got["temp"], ref["temp"], rtol=1e-4)
    assert all(res[r]["keys_sorted"] for r in range(world))
    st = np.stack([res[r]["stats"] for r in range(world)])           # [rank, step, (first, nAssigned, nLocal, level)]
    assert (st[:, :, 1].sum(0) == ref["x"].size).all()               # assigned counts always sum to N
    assert (st[:, :, 2] > st[:, :, 1]).all()                          # halos are really there
    imbalance = st[:, :, 1].max(0) / st[:, :, 1].mean(0)
    assert imbalance.max() < 1.3


def test_more_ranks_than_occupied_cells():
    """a rank may end up with NO particles (here: one occupied cell of the plan, two ranks): the sync, the distributed
    hydro step, the reductions and integrate go through on the empty rank and the loop equals the one-rank loop"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ref, ref_res = _run_sim(1, "onecell", 5, 2)
    got, res = _run_sim(2, "onecell", 5, 2)
    st = np.stack([res[r]["stats"] for r in range(2)])       # [rank, step, (first, nAssigned, nLocal, level)]
    assert (st[:, :, 3] == 1).all(), st[:, :, 3]
    assert (st[:, :, 1].min(0) == 0).all() and (st[:, :, 1].max(0) == 125).all(), st[:, :, 1]
    np.testing.assert_array_equal(got["nc"], ref["nc"])
    np.testing.assert_array_equal(res[0]["rows"][:, 8], ref_res[0]["rows"][:, 8])
    np.testing.assert_allclose(res[0]["rows"][:, 2:6], ref_res[0]["rows"][:, 2:6], rtol=1e-6)
    for k in ("x", "y", "z", "h"):
        np.testing.assert_allclose(got[k], ref[k], rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("world", [1, 2])
def test_distributed_turbulence_ve_loop(world):
    """turbulence-ve propagator (hydro step + driveTurbulence) under the dynamic decomposition: world = 1 is bitwise the
    single-rank Simulation; on 2 ranks every rank advances the same stirring state and stirs its own particles"""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import sphexa_b200 as sx
    from sphexa_b200 import sim
    side, steps = 20, 10
    merged, res = _run_sim(world, "turbstir", side, steps)
    g = _case("turbstir", side)
    s = sim.Simulation(g["x"].size, g["box"], g["boundary"], g["params"])
    n = g["x"].size
    up = {k: (v if isinstance(v, np.ndarray) and v.shape == (n,) else np.full(n, v)) for k, v in g["fields"].items()}
    s.set_fields(x=g["x"], y=g["y"], z=g["z"], **up)
    s.sync()
    s.f["id"].copy_(torch.arange(n, dtype=torch.int64, device=s.device))
    s.turbulence = sim.Turbulence()
    rows = np.array([s.step() for _ in range(steps)], np.float64)
    got = res[0]["rows"]
    assert rows[-1, 4] > 0  # the stirring has set the gas in motion
    if world == 1:
        np.testing.assert_array_equal(got[:, 1:3], rows[:, 1:3])
        np.testing.assert_array_equal(got[:, 8], rows[:, 8])
        np.testing.assert_allclose(got[:, 3:6], rows[:, 3:6], rtol=1e-12)
    else:
        np.testing.assert_allclose(got[:, 2], rows[:, 2], rtol=1e-5)
        np.testing.assert_allclose(got[:, 3], rows[:, 3], rtol=1e-9)
        np.testing.assert_allclose(got[1:, 4], rows[1:, 4], rtol=1e-4)
        np.testing.assert_array_equal(got[:3, 8], rows[:3, 8])
