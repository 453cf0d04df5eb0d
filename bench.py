#!/usr/bin/env python
"""bench.py — SPH-VE hydro-step throughput (particles/s) on B200, BASELINE.json metric.

    python bench.py --gpus N --steps K --warmup W [--case sedov|noh|turbulence]   # our arm (libsphx.so, sm_100a)
    python bench.py --impl reference --steps K --warmup W [--case ...]             # the reference's OpenMP CPU build

The workload is the reference's own time-step loop (main/src/sphexa/sphexa.cpp:141-170 with HydroVeProp): W + K
EVOLVING steps from the initial condition, every step = Domain::sync -> computeForces -> conserved quantities ->
integrate. A "step" of the metric is computeForces (main/src/propagator/ve_hydro.hpp:147-190: neighbour search +
h-iteration + XMass, VeDefGradh, EOS, IAD + divv/curlv, AV switches, momentum + energy, with the halo exchanges on
N > 1): `value` = particles / mean computeForces time over the K timed steps (CUDA events around it, max over ranks).
Domain::sync, conserved and integrate run between the timed intervals and are reported beside it.
Both arms run the same case, side and step count. Prints ONE JSON line (rank 0); DESIGN.md §7 defines every key.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

METRIC = "sph_hydro_step_particles_per_sec"
UNIT = "particles/s"
DTYPE = "f32+f64 (mixed, reference production types)"

# algorithmic (compulsory) bytes per particle and loop: every field read once / written once, neighbour gathers and
# the neighbour list NOT counted. Mixed precision, avClean = false, ideal gas via temp.
# The 44 B of "XMass incl. neighbour build" are split over the two kernels that do it here: the block search reads
# x, y, z, h and writes h, nc (36 B); the XMass loop adds m in, xm out (8 B; its coordinate reads are counted once).
ALGO_BYTES = {"find_neighbors": 36, "xmass": 8, "ve_def_gradh": 44, "eos": 32, "iad_divv_curlv": 80,
              "av_switches": 88, "momentum_energy": 108}
KERNEL_OF = {"find_neighbors": "blockSearchKernel", "xmass": "loopKernel<XMassOp>", "ve_def_gradh": "loopKernel<GradhOp>",
             "eos": "eosKernel", "iad_divv_curlv": "loopKernel<IadOp>", "av_switches": "loopKernel<AvOp>",
             "momentum_energy": "loopKernel<MomentumOp<0>>"}
TRAFFIC_KEY = {"find_neighbors": "block_search"}
PHASES = list(ALGO_BYTES)
CASE_TITLE = {"sedov": "Sedov blast wave", "noh": "Noh implosion", "turbulence": "Subsonic turbulence box"}
# sides of the mid-size cases of the N-rank vs 1-rank parity check (outside the timed region)
PARITY_SIDES = {"sedov": 100, "noh": 80, "turbulence": 64}
PARITY_FIELDS = ["xm", "kx", "gradh", "prho", "c", "c11", "c12", "c13", "c22", "c23", "c33", "divv", "curlv", "alpha",
                 "ax", "ay", "az", "du"]
FAMILIES = [("c11", "c12", "c13", "c22", "c23", "c33"), ("divv", "curlv"), ("ax", "ay", "az")]


def finite_json(o):
    """strict JSON has no Infinity / NaN: non-finite numbers become null (e.g. minDtRho = inf for a fluid at rest)"""
    if isinstance(o, float):
        return o if o == o and abs(o) != float("inf") else None
    if isinstance(o, dict):
        return {k: finite_json(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [finite_json(v) for v in o]
    return o


def default_side(case: str, world: int, scaling: str) -> int:
    """BASELINE configs: Sedov 200^3 on 1 GPU ... 400^3 on 8 (weak: 8 M particles per GPU); Noh 150^3 (config 3, quoted
    on 2 GPUs); turbulence 300^3 (config 4, quoted on 4 GPUs; 27 M particles also fit one). Strong scaling: Sedov 400^3."""
    if case == "sedov":
        return 400 if scaling == "strong" else int(round(200 * world ** (1.0 / 3.0)))
    return 150 if case == "noh" else 300


def make_config(case: str, side: int, steps: int, warmup: int) -> dict:
    """the workload both arms run; identical in both JSON lines"""
    return {"workload": f"{CASE_TITLE[case]} {side}^3, SPH-VE hydro step (neighbour search + 6 particle loops), "
                        f"ng0=100 ngmax=150, {warmup + steps} evolving time steps from the initial condition "
                        f"(sync -> computeForces -> integrate), the first {warmup} dropped",
            "case": case, "side": side}


def measured_peaks():
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", os.environ.get("SPHX_BENCH_CLOCK_MS", "100"), "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for nme, v in zip(names, parts[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_numa(local_rank: int):
    """pin this rank's host threads (and with them its pinned staging buffers, first touch) to the CPUs next to its GPU"""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:  # noqa: BLE001
        pass
    return None


# ------------------------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation (oracle/_ref = unmodified sources, OpenMP)
# ------------------------------------------------------------------------------------------------------------------
REF_PHASE_LABELS = ["FindNeighbors", "XMass", "Normalization & Gradh", "EquationOfState", "IadVelocityDivCurl",
                    "AVswitches", "MomentumAndEnergy"]


def cpu_model() -> str:
    try:
        return next(ln.split(":")[1].strip() for ln in open("/proc/cpuinfo") if ln.startswith("model name"))
    except (OSError, StopIteration):
        return "unknown"


def run_reference_cpu(case: str, side: int, steps: int, warmup: int) -> dict:
    """The reference's OpenMP CPU build on all host cores, `warmup + steps` evolving steps of `case` at `side`^3.
    Sedov: its own CLI (`sphexa_ref --init sedov -n side -s N --ascii`), per-phase timer lines parsed
    (main/src/propagator/ve_hydro.hpp:135-191). Noh / turbulence: the CLI needs a glass file for those, so the same
    reference headers are driven by oracle/ref_harness.cpp (timing build, synthetic jittered lattice).
    Hydro-step time = FindNeighbors + the six loops (no domain::sync, no integrate), as in our arm."""
    cores = os.cpu_count() or 1
    env = dict(os.environ, OMP_NUM_THREADS=str(cores), OMP_PROC_BIND="close")
    total = warmup + steps
    with tempfile.TemporaryDirectory() as tmp:
        if case == "sedov":
            exe = REPO / "oracle" / "_ref" / "sphexa_ref"
            cmd = [str(exe), "--init", "sedov", "-n", str(side), "-s", str(total), "--ascii"]
        else:
            exe = REPO / "oracle" / "_ref" / "ref_harness_timing"
            # one extra step: the harness dumps fields in its first and last step, both are dropped from the timing
            cmd = [str(exe), "noh" if case == "noh" else "turb", str(side), str(total + 1), str(Path(tmp) / "o"),
                   "1000000000", "0"]
        if not exe.exists():
            raise FileNotFoundError(f"{exe} missing (built by __graft_entry__.build() where /root/reference exists)")
        out = subprocess.run(cmd, cwd=tmp, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                             check=True).stdout
    iters, cur, n = [], {}, side ** 3
    for line in out.splitlines():
        if case == "sedov":
            m = re.match(r"# (.+?): ([0-9.eE+-]+)s", line)
            if m and m.group(1) in REF_PHASE_LABELS:
                cur[m.group(1)] = cur.get(m.group(1), 0.0) + float(m.group(2))
            if line.startswith("=== Total time for iteration"):
                iters.append(cur)
                cur = {}
        else:
            m = re.match(r"step (\d+) n (\d+) sync ([0-9.]+) findNeighbors ([0-9.]+) loops ([0-9.]+) integrate ([0-9.]+)", line)
            if m:
                n = int(m.group(2))
                iters.append({"FindNeighbors": float(m.group(4)), "six loops": float(m.group(5)),
                              "_sync": float(m.group(3)), "_integrate": float(m.group(6))})
    timed = iters[warmup:warmup + steps]
    if not timed:
        raise RuntimeError("no timed iterations parsed from the reference output")
    per_step = [sum(v for k, v in it.items() if not k.startswith("_")) for it in timed]
    t = sum(per_step) / len(per_step)
    labels = sorted({k for it in timed for k in it})
    phases = {k: sum(it.get(k, 0.0) for it in timed) / len(timed) * 1e3 for k in labels}
    return {"value": n / t, "ms_per_step": t * 1e3, "cores": cores, "kind": "reference", "cpu_model": cpu_model(),
            "sample": f"{' '.join([Path(cmd[0]).name] + cmd[1:(7 if case == 'sedov' else 4)])} ({n} particles, OpenMP "
                      f"{cores} threads, first {warmup} iterations dropped, {len(timed)} timed); "
                      f"step = FindNeighbors + 6 loops",
            "phases_ms": phases, "n_particles": n}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = max(args.gpus, 1)
    side = args.side or default_side(args.case, world, args.scaling)
    # the CPU cannot hold / finish the multi-GPU sizes in minutes: bounded sample of the same workload
    ref_side = args.ref_side or min(side, {"sedov": 200, "noh": 150, "turbulence": 128}[args.case])
    r = run_reference_cpu(args.case, ref_side, args.steps, args.warmup)
    cfg = make_config(args.case, side, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": DTYPE, "data": "synthetic", "config": cfg,
            "cpu_baseline": {k: r[k] for k in ("value", "cores", "kind", "sample", "cpu_model")} | {"unit": UNIT},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "phases_ms": r["phases_ms"], "sample_side": ref_side, "sample_is_full_workload": ref_side == side,
            "n_particles": r["n_particles"]}
    print(json.dumps(finite_json(line), allow_nan=False))


# ------------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------------
class Runner:
    """The reference's main loop on N GPUs, every stage a libsphx call: sim.Simulation on one rank,
    dist.DistributedSimulation (dynamic SFC decomposition, NCCL) on several. computeForces is issued loop by loop
    (one C-ABI call each, the halo exchanges between them) so that CUDA events can time the phases; the sequence is
    bit-identical to sphx_hydro_step / sphx_hydro_step_dist (tests/test_gpu_parity.py::test_fused_step_equals_loop_calls)."""

    def __init__(self, sx, case: str, side: int, rank: int, world: int, dev, ids_in_generation_order=False, part=True):
        import numpy as np
        import torch
        from sphexa_b200 import cases
        from sphexa_b200 import dist as sdist
        self.sx, self.world, self.rank, self.dev = sx, world, rank, dev
        self.L = sx.load()
        if world == 1 and not ids_in_generation_order:
            self.s = getattr(cases, f"make_{case}_sim")(sx, side, device=dev)
            self.ds = None
            self.n_global = self.s.n
        elif world == 1:
            g = getattr(cases, f"{case}_global")(side)
            n = g["x"].size
            s = sx.sim.Simulation(n, g["box"], g["boundary"], g["params"], device=dev)
            up = {k: (v if isinstance(v, np.ndarray) and v.shape == (n,) else np.full(n, v)) for k, v in g["fields"].items()}
            s.set_fields(x=g["x"], y=g["y"], z=g["z"], **up)
            if "vx" in up:
                p = g["params"]
                for a, b in (("x_m1", "vx"), ("y_m1", "vy"), ("z_m1", "vz")):
                    s.f[a].copy_((s.f[b].double() * p.minDt).float())
            self.s, self.ds, self.n_global = s, None, n  # ids = generation index (Simulation.__init__), moved by sync
        else:
            g = cases.sedov_global(side, (rank, world)) if (case == "sedov" and part) else getattr(cases, f"{case}_global")(side)
            self.ds = sdist.DistributedSimulation(sx.sim, g, rank, world, dev)
            self.s = None
            self.n_global = self.ds.n_global
        self.stream = torch.cuda.current_stream()

    @property
    def hd(self):
        return self.s if self.s is not None else self.ds.hd

    @property
    def p(self):
        return self.s.p if self.s is not None else self.ds.p

    def sync(self):
        (self.s or self.ds).sync()

    def conserved(self):
        return (self.s or self.ds).compute_conserved()

    def integrate(self):
        (self.s or self.ds).integrate()

    def calls(self):
        ds = self.ds
        X = (lambda *names: (lambda: ds.exchange(list(names)))) if ds is not None else (lambda *names: None)
        seq = [("find_neighbors", lambda: self.hd.find_neighbors_sph(sync=False), None),
               ("xmass", lambda: self.hd.xmass(), X("xm")),
               ("ve_def_gradh", lambda: self.hd.ve_def_gradh(), None),
               ("eos", lambda: self.hd.eos(), X("vx", "vy", "vz", "prho", "c", "kx")),
               ("iad_divv_curlv", lambda: self.hd.iad_divv_curlv(sync=False),
                X("c11", "c12", "c13", "c22", "c23", "c33", "divv")),
               ("av_switches", lambda: self.hd.av_switches(), X("alpha")),
               ("momentum_energy", lambda: self.hd.momentum_energy(sync=True), None)]
        out = []
        for name, fn, ex in seq:
            out.append((name, fn))
            if ex is not None:
                out.append(("halo_exchange", ex))
        return out

    def forces(self, events=None):
        """computeForces; events: list of len(calls) + 1 CUDA events recorded around the calls"""
        calls = self.calls()
        for i, (_, fn) in enumerate(calls):
            if events is not None:
                events[i].record(self.stream)
            fn()
        hd = self.hd
        if self.ds is not None:
            self.sx._cabi.check(self.L.sphx_reduce_step_result(self.ds.comm, 0, C.byref(hd.result), None))
            self.ds.result = hd.result
        if events is not None:
            events[len(calls)].record(self.stream)
        return hd.result

    def step(self, ev=None):
        """one iteration of the main loop; ev = (sync0, [forces events], forces1=cons0, cons1=int0, int1)"""
        if ev is not None:
            ev[0].record(self.stream)
        self.sync()
        self.forces(ev[1] if ev is not None else None)
        self.conserved()
        if ev is not None:
            ev[2].record(self.stream)
        self.integrate()
        if ev is not None:
            ev[3].record(self.stream)
        obj = self.s or self.ds
        obj.iteration += 1

    def assigned(self, name):
        hd = self.hd
        a = hd.f[name][hd.first:hd.last].cpu().numpy()
        import numpy as np
        return a.view(np.uint32) if name == "nc" else a

    def close(self):
        if self.ds is not None:
            self.ds.close()


def field_errors(got: dict, ref: dict) -> dict:
    """max relative error per field, |a-b| / max(|a|, |b|, floor); floor = 1e-2 x the max-norm of the field family
    (values that vanish by symmetry carry only the summation noise of their ~100 pair terms), 1e-1 for du / a whose
    pair terms cancel (tests/test_gpu_parity.py documents and measures both floors)"""
    import numpy as np
    out = {}
    for k in PARITY_FIELDS:
        fam = next((f for f in FAMILIES if k in f), (k,))
        scale = max(float(np.abs(ref[m]).max()) for m in fam)
        floor = max(scale, 1e-300) * (1e-1 if k in ("du", "ax", "ay", "az") else 1e-2)
        a, b = got[k].astype(np.float64), ref[k].astype(np.float64)
        out[k] = float((np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), floor)).max())
    return out


def dist_parity(sx, rank, world, dev, dist, tol=2e-4) -> list:
    """N ranks vs rank 0 alone, by particle id (SURVEY 8e "parity without MPI"; the reference's own multi-rank tests are
    domain/test/integration_mpi/exchange_halos_gpu.cpp and domain_nranks.cpp:64-131): one mid-size case per BASELINE
    family goes through the dynamic decomposition (migration, halo discovery, local tree) and the distributed hydro step
    on all ranks, and through the plain single-GPU path on rank 0. nc and h must be identical and the reduced scalars
    agree. The 18 fp32 fields are two evaluations by the SAME kernels with different block boundaries (hence block
    origins and summation order): each is within 1e-4 of the reference (tests/test_gpu_parity.py), so they are held to
    2e-4 of each other, with the floors of that test. Two more steps follow to compare the evolving energies."""
    import numpy as np
    import torch
    rows = []
    for case, side in PARITY_SIDES.items():
        r = Runner(sx, case, side, rank, world, dev, part=False)
        r.sync()
        res = r.forces()
        mine = {k: r.assigned(k) for k in PARITY_FIELDS + ["h", "nc", "id"]}
        scal = (res.minDtCourant, res.minDtRho, int(res.totalNeighbors), int(res.maxNc))
        r.conserved()
        r.integrate()
        ener = []
        for _ in range(2):
            r.sync()
            r.forces()
            c = r.conserved()
            ener.append((c.etot, c.ecin, c.eint, int(c.totalNeighbors)))
            r.integrate()
        n_local = r.hd.n
        r.close()
        del r
        torch.cuda.empty_cache()
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(mine, gathered, dst=0)
        nl = [None] * world if rank == 0 else None
        dist.gather_object(n_local, nl, dst=0)
        if rank != 0:
            continue
        ids = np.concatenate([g["id"] for g in gathered])
        n = ids.size
        got = {}
        for k in PARITY_FIELDS + ["h", "nc"]:
            v = np.concatenate([g[k] for g in gathered])
            full = np.zeros(n, v.dtype)
            full[ids] = v
            got[k] = full
        one = Runner(sx, case, side, 0, 1, dev, ids_in_generation_order=True)
        one.sync()
        res1 = one.forces()
        scal1 = (res1.minDtCourant, res1.minDtRho, int(res1.totalNeighbors), int(res1.maxNc))  # (res1 is overwritten below)
        oid = one.assigned("id")
        ref = {}
        for k in PARITY_FIELDS + ["h", "nc"]:
            v = one.assigned(k)
            full = np.zeros(n, v.dtype)
            full[oid] = v
            ref[k] = full
        one.conserved()
        one.integrate()
        ener1 = []
        for _ in range(2):
            one.sync()
            one.forces()
            c = one.conserved()
            ener1.append((c.etot, c.ecin, c.eint, int(c.totalNeighbors)))
            one.integrate()
        one.close()
        del one
        torch.cuda.empty_cache()
        errs = field_errors(got, ref)
        rel = lambda a, b: abs(a - b) / max(abs(a), abs(b), 1e-300)  # noqa: E731
        row = {"case": f"{case} {side}^3", "n": int(n), "ranks": world,
               "ids_complete": bool(np.array_equal(np.sort(ids), np.arange(n))),
               "nc_mismatch": int((got["nc"] != ref["nc"]).sum()), "h_mismatch": int((got["h"] != ref["h"]).sum()),
               "max_rel_err": max(errs.values()), "fields": errs, "tol": tol,
               "total_neighbors": [scal[2], scal1[2]], "max_nc": [scal[3], scal1[3]],
               "minDtCourant_rel": rel(scal[0], scal1[0]),
               "etot_rel_steps_2_3": [rel(a[0], b[0]) for a, b in zip(ener, ener1)],
               "neighbor_sum_rel_steps_2_3": [rel(a[3], b[3]) for a, b in zip(ener, ener1)],
               "local_particles_incl_halos": [int(v) for v in nl]}
        row["ok"] = bool(row["ids_complete"] and row["nc_mismatch"] == 0 and row["h_mismatch"] == 0 and
                         row["max_rel_err"] <= tol and scal[2] == scal1[2] and scal[3] == scal1[3] and
                         row["minDtCourant_rel"] <= 1e-5 and max(row["etot_rel_steps_2_3"]) <= 1e-6)
        rows.append(row)
    return rows


def our_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libsphx has no CPU fallback")
    numa_cpus = bind_to_gpu_numa(local_rank) if world > 1 else None
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    if world > 1:
        # NCCL for the device-side plumbing of the bench (barriers, max-over-ranks), gloo for host-side object gathers;
        # the data path (halo exchange, migration, reductions) is libsphx's own NCCL communicator
        dist.init_process_group("cpu:gloo,cuda:nccl", device_id=torch.device(dev))

    import sphexa_b200 as sx
    sx.load()
    case, K, W = args.case, args.steps, max(args.warmup, 3)
    side = args.side or default_side(case, world, args.scaling)
    t0 = time.time()
    run = Runner(sx, case, side, rank, world, dev)
    n_global = run.n_global
    stream = run.stream
    call_names = [c[0] for c in run.calls()]
    num_exchanges = call_names.count("halo_exchange")
    # reset-scalars, block search (standard + overflow instantiation), EOS, 5 x (work-counter reset + persistent loop
    # kernel), pack kernel per exchange
    launches_per_step = 14 + num_exchanges

    # ---- continuity with round 1: the frozen initial state (h and alpha restored before every repetition) ----------
    run.sync()
    setup_s = time.time() - t0
    static_ms = None
    if not args.no_static:
        hd = run.hd
        h0, a0 = hd.f["h"].clone(), hd.f["alpha"].clone()
        reps = []
        for k in range(6):
            hd.f["h"].copy_(h0)
            hd.f["alpha"].copy_(a0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            run.forces()
            e1.record(stream)
            torch.cuda.synchronize()
            reps.append(e0.elapsed_time(e1))
        hd.f["h"].copy_(h0)
        hd.f["alpha"].copy_(a0)
        sm = torch.tensor([sum(reps[1:]) / 5], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(sm, op=dist.ReduceOp.MAX)
        static_ms = float(sm.item())

    # ---- the evolving loop: W warm-up steps, then K timed steps ------------------------------------------------------
    for _ in range(W):
        run.step()
    torch.cuda.synchronize()
    mk = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    ev = [(mk(), [mk() for _ in range(len(call_names) + 1)], mk(), mk()) for _ in range(K)]
    nstat = []
    sampler = ClockSampler(local_rank)
    if rank == 0 and os.environ.get("SPHX_BENCH_CLOCK_MS") != "0":
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    wall0 = time.perf_counter()
    for k in range(K):
        run.step(ev[k])
        nstat.append((run.hd.last - run.hd.first, run.hd.n, int(run.hd.result.numHIterated)))
    torch.cuda.synchronize()
    wall_loop = time.perf_counter() - wall0
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None

    forces_ms = [e[1][0].elapsed_time(e[1][-1]) for e in ev]
    static_after_ms = None
    if not args.no_static and world == 1:
        # the same frozen-state repetitions once more, now on the evolved state and a warm GPU: separates what the
        # evolution costs from what the clocks do
        hd = run.hd
        run.sync()
        h1, a1 = hd.f["h"].clone(), hd.f["alpha"].clone()
        reps = []
        for k in range(6):
            hd.f["h"].copy_(h1)
            hd.f["alpha"].copy_(a1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            run.forces()
            e1.record(stream)
            torch.cuda.synchronize()
            reps.append(e0.elapsed_time(e1))
        hd.f["h"].copy_(h1)
        hd.f["alpha"].copy_(a1)
        static_after_ms = sum(reps[1:]) / 5
    phase_ms = {}
    for i, name in enumerate(call_names):
        phase_ms[name] = phase_ms.get(name, 0.0) + sum(e[1][i].elapsed_time(e[1][i + 1]) for e in ev) / K
    other = {"domain_sync": sum(e[0].elapsed_time(e[1][0]) for e in ev) / K,
             "conserved": sum(e[1][-1].elapsed_time(e[2]) for e in ev) / K,
             "integrate": sum(e[2].elapsed_time(e[3]) for e in ev) / K}
    red = torch.tensor([sum(forces_ms), sorted(forces_ms)[K // 2], max(forces_ms), wall_loop * 1e3] +
                       [other[k] for k in ("domain_sync", "conserved", "integrate")], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
    total_ms, median_ms, worst_ms, wall_ms = (float(v) for v in red[:4])
    other = dict(zip(("domain_sync", "conserved", "integrate"), (float(v) for v in red[4:])))
    ms_per_step = total_ms / K
    value = n_global / (ms_per_step * 1e-3)
    res = run.hd.result
    n_assigned = run.hd.last - run.hd.first
    cons = run.conserved()
    check = {"total_neighbors": int(res.totalNeighbors), "mean_nc": res.totalNeighbors / n_global, "max_nc": int(res.maxNc),
             "minDtCourant": res.minDtCourant, "minDtRho": res.minDtRho, "minDt": run.p.minDt, "ttot": run.p.ttot,
             "etot": cons.etot, "ecin": cons.ecin, "eint": cons.eint,
             "h_iterated_last_step": nstat[-1][2], "static_ms_per_step": static_ms,
             "static_ms_per_step_evolved_state": static_after_ms, "ms_all_steps": [round(v, 3) for v in forces_ms],
             "phases_ms_slowest_step": {n: round(ev[forces_ms.index(max(forces_ms))][1][i].elapsed_time(
                 ev[forces_ms.index(max(forces_ms))][1][i + 1]), 3) for i, n in enumerate(call_names)},
             "phases_ms_fastest_step": {n: round(ev[forces_ms.index(min(forces_ms))][1][i].elapsed_time(
                 ev[forces_ms.index(min(forces_ms))][1][i + 1]), 3) for i, n in enumerate(call_names)},
             "static_note": "round-1 definition, kept for continuity: the frozen initial state, h and alpha restored "
                            "before every repetition"}

    # ---- end-to-end: host buffers in, host results out, through the same C-ABI calls -----------------------------
    run.sync()  # the last integrate moved the particles: tree and order of the state the batches start from
    hd = run.hd
    nloc = hd.n
    bs = hd.block_stats() if world == 1 else None
    in_names = ["x", "y", "z", "h", "temp", "vx", "vy", "vz", "alpha"]  # m is step-invariant: resident
    out_names = ["ax", "ay", "az", "du", "h", "nc"]
    host_in = {k: hd.f[k][:nloc].cpu().pin_memory() for k in in_names}
    host_out = {k: torch.empty(nloc, dtype=hd.f[k].dtype).pin_memory() for k in out_names}
    h2d = sum(t.numel() * t.element_size() for t in host_in.values())
    d2h = sum(t.numel() * t.element_size() for t in host_out.values()) + C.sizeof(sx._cabi.SphxStepResult)

    # Every batch has its own inputs in pinned host memory and returns its results to pinned host memory; consecutive
    # batches are independent (each starts from the host copy of the same state), so they are pipelined over two sets of
    # device buffers for the uploaded and downloaded fields: the upload of batch k+1 (copy stream) and the download of
    # batch k-1's results (third stream) overlap the kernels of batch k. Inside a batch the inputs are uploaded in the
    # order the loops consume them and every loop waits only for the fields it reads. The C ABI re-reads the field
    # pointers on every call, so switching sets is just handing it the other pointers. All copies of all batches,
    # including the first upload and the last download, are inside the timed region.
    first_use = {"find_neighbors": "h", "eos": "vz", "av_switches": "alpha"}
    up_stream, down_stream = torch.cuda.Stream(), torch.cuda.Stream()
    io_names = list(dict.fromkeys(in_names + out_names))
    sets = [{k: hd.f[k] for k in io_names}, {k: torch.empty_like(hd.f[k]) for k in io_names}]
    uploaded = [None, None]     # per set: events of the upload in flight
    set_free = [None, None]     # per set: its results have reached the host (the set may be overwritten)
    calls = run.calls()

    def upload(s):
        if set_free[s] is not None:
            up_stream.wait_event(set_free[s])
        else:
            up_stream.wait_stream(stream)  # whatever ran before the pipeline no longer reads the fields
        done = {}
        with torch.cuda.stream(up_stream):
            for k in in_names:
                sets[s][k][:nloc].copy_(host_in[k], non_blocking=True)
                done[k] = up_stream.record_event()
        uploaded[s] = done

    def e2e_pipeline(nsteps):
        upload(0)
        for k in range(nsteps):
            s = k % 2
            if k + 1 < nsteps:
                upload(1 - s)  # enqueued before this batch's kernels: runs beside them
            hd.f.update(sets[s])
            done = uploaded[s]
            for name, fn in calls:
                if name in first_use:
                    stream.wait_event(done[first_use[name]])
                fn()  # the loops and, on N > 1, the NCCL halo exchanges between them
                if name == "find_neighbors":
                    down_stream.wait_event(stream.record_event())
                    with torch.cuda.stream(down_stream):
                        for f in ("h", "nc"):
                            host_out[f].copy_(sets[s][f][:nloc], non_blocking=True)
            if world > 1:
                sx._cabi.check(run.L.sphx_reduce_step_result(run.ds.comm, 0, C.byref(hd.result), None))
            down_stream.wait_event(stream.record_event())
            with torch.cuda.stream(down_stream):
                for f in ("ax", "ay", "az", "du"):
                    host_out[f].copy_(sets[s][f][:nloc], non_blocking=True)
                set_free[s] = down_stream.record_event()
        stream.wait_stream(down_stream)
        stream.wait_stream(up_stream)

    e2e_pipeline(2)
    torch.cuda.synchronize()
    # the pipelined batches return what the device-resident step returns
    got = {k: host_out[k].clone() for k in ("ax", "du", "nc")}
    hd.f.update(sets[0])
    for k in in_names:
        hd.f[k][:nloc].copy_(host_in[k])
    run.forces()
    torch.cuda.synchronize()
    for k, v in got.items():  # outputs exist for the assigned particles only
        assert torch.equal(v[hd.first:hd.last], hd.f[k][hd.first:hd.last].cpu()), \
            f"e2e pipeline: {k} differs from the device-resident step"
    e2e_steps = max(4, min(K, 10))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    set_free[0] = set_free[1] = None
    e0, e1 = mk(), mk()
    e0.record(stream)
    e2e_pipeline(e2e_steps)
    e1.record(stream)
    torch.cuda.synchronize()
    hd.f.update(sets[0])
    e2e_ms = torch.tensor([e0.elapsed_time(e1) / e2e_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = n_global / (float(e2e_ms.item()) * 1e-3)
    del sets, host_in, host_out

    # ---- N > 1: parity of the distributed path, where the driver can see it (outside every timed region) -----------
    parity = None
    if world > 1 and not args.no_parity:
        run.close()
        del run, hd
        torch.cuda.empty_cache()
        try:
            parity = dist_parity(sx, rank, world, dev, dist)
        except Exception as e:  # noqa: BLE001
            parity = [{"case": "error", "ok": False, "error": repr(e)}]

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    dom = max(PHASES, key=lambda k: phase_ms[k])
    n_rank = float(np.mean([s[0] for s in nstat]))  # algorithmic bytes count the particles a rank computes
    achieved = ALGO_BYTES[dom] * n_rank / (phase_ms[dom] * 1e-3) / 1e9
    traffic = None
    tfile = REPO / "profiles" / "traffic.json"
    tj_all = json.loads(tfile.read_text()) if tfile.exists() else {}
    tag = f"{case}{side}"
    traffic = tj_all.get(f"{TRAFFIC_KEY.get(dom, dom)}@{tag}")
    mean_nc = res.totalNeighbors / n_global
    # bytes this design moves on top of the compulsory field traffic ("if the implementation materialises
    # neighbour lists in HBM, add the list term it actually needs"): 2 B per stored neighbour (16-bit block-local
    # indices) and one 16 B candidate record per block-local candidate, written once by the search and read once per
    # consuming loop (the two passes of IAD + divv/curlv read the list twice)
    cand_pp = (bs["candTop"] / n_rank) if bs is not None else None
    list_b = 2.0 * (mean_nc - 1)
    cand_b = 16.0 * cand_pp if cand_pp is not None else None
    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    issue_peak = sm_count * 4 * sm_mhz * 1e6  # warp instructions / s

    def per_kernel(k):
        t = phase_ms[k] * 1e-3
        out = {"ms": phase_ms[k], "GBps": ALGO_BYTES[k] * n_rank / t / 1e9, "frac": ALGO_BYTES[k] * n_rank / t / 1e9 / peak}
        if cand_b is not None:
            design = ALGO_BYTES[k] + (0 if k == "eos" else list_b * (2 if k == "iad_divv_curlv" else 1) + cand_b)
            out |= {"design_bytes_per_particle": design, "design_GBps": design * n_rank / t / 1e9,
                    "design_frac": design * n_rank / t / 1e9 / peak}
        inst = tj_all.get(f"inst:{TRAFFIC_KEY.get(k, k)}@{tag}") if world == 1 else None
        if inst:
            out["warp_inst_per_launch"] = inst
            out["issue_frac"] = inst / t / issue_peak
        return out

    roofline = {"bound": "hbm", "kernel": KERNEL_OF[dom], "phase": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_particle": ALGO_BYTES[dom],
                "neighbor_list_bytes_per_particle": list_b, "candidate_bytes_per_particle": cand_b,
                "step": {"algorithmic_bytes_per_particle": sum(ALGO_BYTES.values()),
                         "GBps": sum(ALGO_BYTES.values()) * n_rank / (ms_per_step * 1e-3) / 1e9,
                         "frac": sum(ALGO_BYTES.values()) * n_rank / (ms_per_step * 1e-3) / 1e9 / peak},
                "note": "achieved/frac: algorithmic bytes = compulsory field traffic only (SURVEY 8d); design_*: "
                        "compulsory + the 16-bit neighbour list and candidate records this design stores in HBM; "
                        "issue_frac: warp instructions (ncu capture under profiles/) / time / (4 per SM and cycle), the "
                        "bound that actually binds these kernels",
                "issue_peak_warp_inst_per_s": issue_peak,
                "per_kernel": {k: per_kernel(k) for k in PHASES}}
    pk = roofline["per_kernel"]
    if all("warp_inst_per_launch" in pk[k] for k in PHASES):
        roofline["issue_frac_step"] = sum(pk[k]["warp_inst_per_launch"] for k in PHASES) / (ms_per_step * 1e-3) / issue_peak

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            cs = {"sedov": 128, "noh": 100, "turbulence": 100}[case]
            r = run_reference_cpu(case, cs, 5, 1)
            cpu_baseline = {k: r[k] for k in ("value", "cores", "kind", "sample", "cpu_model")} | {"unit": UNIT}
        except Exception as e:  # noqa: BLE001
            cpu_baseline = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference",
                            "sample": f"unavailable: {e}"}

    if bs is not None:
        check |= {"candidates_per_block_mean": float(bs["numCand"].mean()), "candidates_per_block_max": int(bs["numCand"].max()),
                  "candidates_per_particle": cand_pp, "fold_blocks": int((bs["flags"] & 1).sum())}
    if parity is not None:
        check["dist_parity"] = parity
        check["dist_parity_ok"] = all(r.get("ok") for r in parity)
    loop_ms = ms_per_step + sum(other.values())
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": DTYPE, "data": "synthetic",
            "config": make_config(case, side, K, W),
            "n_particles": n_global, "particles_per_gpu": n_global // world,
            "assigned_rank0": n_assigned, "local_rank0_incl_halos": nstat[-1][1],
            "cache": "inputs (fields + neighbour list, > 3 GB per GPU) exceed the 126 MB L2; every step works on a new state",
            "parallelism": (f"dynamic SFC (Hilbert) domain decomposition over {world} GPUs redone every step, 4 NCCL halo "
                            f"exchanges per hydro step (send/recv), one process per GPU") if world > 1 else "single GPU",
            "timing": "value = particles / mean over the K timed steps of the CUDA-event interval around computeForces "
                      "(max over ranks of the sum); sync / conserved / integrate run between the intervals (next_rows)",
            "ms_per_step_median": median_ms, "ms_per_step_max": worst_ms,
            "roofline": roofline, "cpu_baseline": cpu_baseline,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": float(e2e_ms.item()),
                    "host_GBps_per_rank": (h2d + d2h) / (float(e2e_ms.item()) * 1e-3) / 1e9,
                    "note": "independent batches of the last evolved state through the C ABI with pinned HOST buffers: "
                            "x,y,z,h,temp,v,alpha up, ax,ay,az,du,h,nc + scalars down, every batch; m is step-invariant "
                            "and stays resident", "numa_cpus_bound": numa_cpus},
            "gpu_launches": launches_per_step * K, "clocks": clocks, "phases_ms": phase_ms, "check": check,
            "next_rows": {"ms": other, "loop_ms_per_step": loop_ms, "wall_ms_per_step": wall_ms / K,
                          "particles_per_sec_whole_loop": n_global / (loop_ms * 1e-3),
                          "note": "the callers of the hot path, executed between the timed intervals of every step: "
                                  "Domain::sync (keys, radix sort, octree, field reorder; on N > 1 also the global cell "
                                  "histogram, decomposition plan, migration, halo exchange of x,y,z,h,m), conserved "
                                  "quantities, integrate"},
            "setup_s": setup_s}
    print(json.dumps(finite_json(line), allow_nan=False))
    if world > 1:
        dist.destroy_process_group()
    if parity is not None and not check["dist_parity_ok"]:
        sys.exit(3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--case", default="sedov", choices=["sedov", "noh", "turbulence"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: 8 M particles per GPU (Sedov 200^3 x N); strong: Sedov 400^3 on any N")
    ap.add_argument("--side", type=int, default=0,
                    help="global lattice side; default: Sedov 200 * gpus^(1/3) (BASELINE configs 1 and 4), Noh 150, "
                         "turbulence 300")
    ap.add_argument("--ref-side", type=int, default=0,
                    help="lattice side the CPU reference arm runs (default: the workload's own side, capped at Sedov "
                         "200^3 / Noh 150^3 / turbulence 128^3 so that the arm ends within minutes)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-static", action="store_true", help="skip the frozen-state timing kept for continuity")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the N-rank vs 1-rank parity check")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        our_arm(args)


if __name__ == "__main__":
    main()
