"""Synthetic initial conditions of the benchmark configurations, generated on the host with numpy and
uploaded through sim.HydroData. Formulas follow the reference initialisers (field formulas only; the reference's
init/ directory is otherwise out of scope):
  Sedov       main/src/init/sedov_init.hpp:49-95, sedov_constants.hpp, grid.hpp:101-131
  Noh         main/src/init/noh_init.hpp:46-103 on a jittered lattice cut to a sphere
  turbulence  main/src/init/turbulence_init.hpp:48-102 on a jittered lattice + imposed solenoidal velocity field
"""
from __future__ import annotations

import numpy as np
import torch

from . import host
from .sim import HydroData, Params


def ideal_gas_cv(mui: float, gamma: float) -> np.float32:
    """sph::idealGasCv (sph/include/sph/eos.hpp:18-23): evaluated in double, returned in the type of mui (float)"""
    return np.float32(np.float64(np.float32(8.317e7) / np.float32(mui)) / (np.float64(gamma) - np.float64(1.0)))


def regular_grid(r: float, side: int):
    step = (2.0 * r) / side
    r_ini = -r + 0.5 * step
    c = r_ini + np.arange(side, dtype=np.float64) * step
    z, y, x = np.meshgrid(c, c, c, indexing="ij")  # index order z-major, x fastest (grid.hpp:101-131)
    return x.ravel().copy(), y.ravel().copy(), z.ravel().copy()


def jittered_lattice(r: float, side: int, seed: int = 42):
    x, y, z = regular_grid(r, side)
    rng = np.random.default_rng(seed)
    step = 2.0 * r / side
    j = rng.uniform(-0.2, 0.2, size=(x.size, 3)) * step
    return x + j[:, 0], y + j[:, 1], z + j[:, 2]


def _finish(sx, x, y, z, box_lim, boundary, p: Params, fields: dict, device, bucket_size=64) -> HydroData:
    """SFC-sort (as the reference's syncCoords / Domain::sync would), build the octree view, upload."""
    t = host.build_tree(x, y, z, box_lim, boundary, bucket_size)
    o = t.order
    n = x.size
    hd = HydroData(n, 0, n, box_lim, boundary, p, device=device)
    up = {k: (v[o] if isinstance(v, np.ndarray) and v.shape == (n,) else np.full(n, v)) for k, v in fields.items()}
    hd.set_fields(x=x[o], y=y[o], z=z[o], **up)
    hd.set_tree(t)
    hd.host_tree = t
    return hd


def sedov_fields(x, y, z, p: Params, r1=0.5, mTotal=1.0, width=0.1, u0=1e-8, energyTotal=1.0, n_total=None):
    n = x.size if n_total is None else n_total
    totalVolume = (2 * r1) ** 3
    hInit = np.cbrt(3.0 / (4 * np.pi) * p.ng0 * totalVolume / n) * 0.5
    ener0 = energyTotal / np.pi ** 1.5 / 1.0 / width ** 3.0
    cv = ideal_gas_cv(p.muiConst, p.gamma)
    r2 = x * x + y * y + z * z
    ui = ener0 * np.exp(-(r2 / (width * width))) + u0
    return dict(h=np.float32(hInit), m=np.float32(mTotal / n), vx=np.float32(0), vy=np.float32(0), vz=np.float32(0),
                temp=ui / np.float64(cv), alpha=np.float32(p.alphamin))


def regular_grid_slice(r: float, side: int, b: int, e: int):
    """particles [b, e) of regular_grid (generation order: z-major, x fastest) without building the whole lattice"""
    step = (2.0 * r) / side
    c = (-r + 0.5 * step) + np.arange(side, dtype=np.float64) * step
    idx = np.arange(b, e, dtype=np.int64)
    return c[idx % side].copy(), c[(idx // side) % side].copy(), c[idx // (side * side)].copy()


def sedov_global(side: int, part: tuple[int, int] | None = None) -> dict:
    """host-side description of the Sedov case (positions in generation order, per-particle fields, box, params):
    input of dist.DistributedHydro / dist.DistributedSimulation. part = (rank, nranks): only that rank's contiguous slice
    of the generation order is built (a 400^3 lattice is 64 M particles; no rank needs all of it)."""
    p = Params(minDt=1e-6, minDt_m1=1e-6, gamma=5.0 / 3.0, muiConst=10.0)
    n = side ** 3
    if part is None:
        x, y, z = regular_grid(0.5, side)
        return dict(x=x, y=y, z=z, fields=sedov_fields(x, y, z, p), params=p, box=[-0.5, 0.5] * 3, boundary=[1, 1, 1])
    b, e = part[0] * n // part[1], (part[0] + 1) * n // part[1]
    x, y, z = regular_grid_slice(0.5, side, b, e)
    return dict(x=x, y=y, z=z, fields=sedov_fields(x, y, z, p, n_total=n), params=p, box=[-0.5, 0.5] * 3,
                boundary=[1, 1, 1], slice=(b, e), n_global=n)


def noh_global(side: int) -> dict:
    """Noh implosion (see make_noh) as a host-side description"""
    p = Params(minDt=1e-4, minDt_m1=1e-4, gamma=5.0 / 3.0, muiConst=10.0)
    x, y, z = jittered_lattice(0.5, side)
    keep = np.sqrt(x * x + y * y + z * z) <= 0.5
    x, y, z = x[keep], y[keep], z[keep]
    n = x.size
    hInit = np.cbrt(3.0 / (4 * np.pi) * p.ng0 * (4.0 * np.pi / 3.0 * 0.5 ** 3) / n) * 0.5
    cv = ideal_gas_cv(p.muiConst, p.gamma)
    radius = np.maximum(np.sqrt(x * x + y * y + z * z), 1e-10)
    f = dict(h=np.float32(hInit), m=np.float32(1.0 / n), temp=np.float64(1e-20) / np.float64(cv),
             alpha=np.float32(p.alphamin), vx=(-1.0 * (x / radius)).astype(np.float32),
             vy=(-1.0 * (y / radius)).astype(np.float32), vz=(-1.0 * (z / radius)).astype(np.float32))
    return dict(x=x, y=y, z=z, fields=f, params=p, box=[-0.5, 0.5] * 3, boundary=[0, 0, 0])


def turbulence_global(side: int) -> dict:
    """turbulence box (see make_turbulence) as a host-side description"""
    p = Params(minDt=1e-4, minDt_m1=1e-4, gamma=1.001, muiConst=0.62, Kcour=0.4)
    x, y, z = jittered_lattice(0.5, side)
    fold = lambda a: np.where(a > 0.5, a - 1.0, np.where(a < -0.5, a + 1.0, a))  # noqa: E731  putInBox
    x, y, z = fold(x), fold(y), fold(z)
    n = x.size
    hInit = np.cbrt(3.0 / (4 * np.pi) * p.ng0 * 1.0 / n) * 0.5
    cv = ideal_gas_cv(p.muiConst, p.gamma)
    cs = np.sqrt(p.gamma * (p.gamma - 1.0) * 1000.0)
    f = dict(h=np.float32(hInit), m=np.float32(1.0 / n), temp=np.float64(1000.0) / np.float64(cv),
             alpha=np.float32(p.alphamin), vx=(0.3 * cs * np.sin(2 * np.pi * y)).astype(np.float32),
             vy=(0.3 * cs * np.sin(2 * np.pi * z)).astype(np.float32),
             vz=(0.3 * cs * np.sin(2 * np.pi * x)).astype(np.float32))
    return dict(x=x, y=y, z=z, fields=f, params=p, box=[-0.5, 0.5] * 3, boundary=[1, 1, 1])


def make_sedov(sx, side: int, device="cuda:0") -> HydroData:
    """Sedov blast wave on a side^3 lattice, periodic box (-0.5, 0.5)^3 (sedov_init.hpp:98-131)."""
    p = Params(minDt=1e-6, minDt_m1=1e-6, gamma=5.0 / 3.0, muiConst=10.0)
    x, y, z = regular_grid(0.5, side)
    return _finish(sx, x, y, z, [-0.5, 0.5] * 3, [1, 1, 1], p, sedov_fields(x, y, z, p), device)


def make_noh(sx, side: int, device="cuda:0", bucket_size: int = 64) -> HydroData:
    """Noh implosion: jittered lattice cut to the sphere r <= 0.5, open box (noh_init.hpp:46-103)."""
    p = Params(minDt=1e-4, minDt_m1=1e-4, gamma=5.0 / 3.0, muiConst=10.0)
    x, y, z = jittered_lattice(0.5, side)
    keep = np.sqrt(x * x + y * y + z * z) <= 0.5
    x, y, z = x[keep], y[keep], z[keep]
    n = x.size
    totalVolume = 4.0 * np.pi / 3.0 * 0.5 ** 3
    hInit = np.cbrt(3.0 / (4 * np.pi) * p.ng0 * totalVolume / n) * 0.5
    cv = ideal_gas_cv(p.muiConst, p.gamma)
    radius = np.maximum(np.sqrt(x * x + y * y + z * z), 1e-10)
    f = dict(h=np.float32(hInit), m=np.float32(1.0 / n), temp=np.float64(1e-20) / np.float64(cv),
             alpha=np.float32(p.alphamin), vx=(-1.0 * (x / radius)).astype(np.float32),
             vy=(-1.0 * (y / radius)).astype(np.float32), vz=(-1.0 * (z / radius)).astype(np.float32))
    return _finish(sx, x, y, z, [-0.5, 0.5] * 3, [0, 0, 0], p, f, device, bucket_size)


def make_turbulence(sx, side: int, device="cuda:0") -> HydroData:
    """Subsonic turbulence box: jittered lattice, periodic, gamma = 1.001, imposed solenoidal velocity field
    v = 0.3 c_s (sin 2 pi y, sin 2 pi z, sin 2 pi x) (stirring itself is out of scope)."""
    p = Params(minDt=1e-4, minDt_m1=1e-4, gamma=1.001, muiConst=0.62, Kcour=0.4)
    x, y, z = jittered_lattice(0.5, side)
    fold = lambda a: np.where(a > 0.5, a - 1.0, np.where(a < -0.5, a + 1.0, a))  # noqa: E731  putInBox
    x, y, z = fold(x), fold(y), fold(z)
    n = x.size
    hInit = np.cbrt(3.0 / (4 * np.pi) * p.ng0 * 1.0 / n) * 0.5
    cv = ideal_gas_cv(p.muiConst, p.gamma)
    cs = np.sqrt(p.gamma * (p.gamma - 1.0) * 1000.0)
    f = dict(h=np.float32(hInit), m=np.float32(1.0 / n), temp=np.float64(1000.0) / np.float64(cv),
             alpha=np.float32(p.alphamin), vx=(0.3 * cs * np.sin(2 * np.pi * y)).astype(np.float32),
             vy=(0.3 * cs * np.sin(2 * np.pi * z)).astype(np.float32),
             vz=(0.3 * cs * np.sin(2 * np.pi * x)).astype(np.float32))
    return _finish(sx, x, y, z, [-0.5, 0.5] * 3, [1, 1, 1], p, f, device)


def make_sedov_sim(sx, side: int, device="cuda:0"):
    """Sedov case as a sim.Simulation: particles uploaded in GENERATION order (z-major lattice); the first
    Simulation.sync() sorts them on the device and builds the tree (sedov_init.hpp:98-131 + sphexa.cpp:141)."""
    from .sim import Simulation

    g = sedov_global(side)
    return _make_sim(g, device)


def make_noh_sim(sx, side: int, device="cuda:0"):
    return _make_sim(noh_global(side), device)


def make_turbulence_sim(sx, side: int, device="cuda:0"):
    return _make_sim(turbulence_global(side), device)


def _make_sim(g: dict, device):
    from .sim import Simulation

    n = g["x"].size
    s = Simulation(n, g["box"], g["boundary"], g["params"], device=device)
    up = {k: (v if isinstance(v, np.ndarray) and v.shape == (n,) else np.full(n, v)) for k, v in g["fields"].items()}
    s.set_fields(x=g["x"], y=g["y"], z=g["z"], **up)
    if "vx" in up:  # x_m1 = v * minDt (noh_init.hpp:86-88, turbulence_init.hpp:96-98); zero for Sedov
        p = g["params"]
        s.set_fields(x_m1=(up["vx"].astype(np.float64) * p.minDt).astype(np.float32),
                     y_m1=(up["vy"].astype(np.float64) * p.minDt).astype(np.float32),
                     z_m1=(up["vz"].astype(np.float64) * p.minDt).astype(np.float32))
    # the reference initialisers SFC-sort the coordinates first (syncCoords, init/utils.hpp) and number the particles
    # afterwards (generateParticleIDs, sedov_init.hpp:76): ids are SFC ranks of the initial state
    s.sync()
    s.f["id"].copy_(torch.arange(n, dtype=torch.int64, device=s.device))
    return s
