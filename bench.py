#!/usr/bin/env python
"""bench.py — SPH-VE hydro-step throughput (particles/s) on B200, BASELINE.json metric.

    python bench.py --gpus N --steps K --warmup W            # our arm (libsphx.so, CUDA sm_100a)
    python bench.py --impl reference --steps K --warmup W    # the reference's own OpenMP CPU build on the host cores

A "step" is one pass of the hot path over all particles: neighbour search + h-iteration + XMass, VeDefGradh, EOS,
IAD + divv/curlv, AV switches, momentum + energy (HydroVeProp::computeForces without Domain::sync,
main/src/propagator/ve_hydro.hpp:147-190), on the synthetic Sedov lattice of SURVEY §8(d).
Prints ONE JSON line (rank 0). See DESIGN.md §Measurement for the definition of every key.
"""
from __future__ import annotations

import argparse
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


def finite_json(o):
    """strict JSON has no Infinity / NaN: non-finite numbers become null (e.g. minDtRho = inf for a fluid at rest)"""
    if isinstance(o, float):
        return o if o == o and abs(o) != float("inf") else None
    if isinstance(o, dict):
        return {k: finite_json(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [finite_json(v) for v in o]
    return o


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
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
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


# ------------------------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation (oracle/_ref/sphexa_ref = unmodified sources, OpenMP)
# ------------------------------------------------------------------------------------------------------------------
REF_PHASE_LABELS = ["FindNeighbors", "XMass", "Normalization & Gradh", "EquationOfState", "IadVelocityDivCurl",
                    "AVswitches", "MomentumAndEnergy"]


def run_reference_cpu(side: int, steps: int, warmup: int) -> dict:
    """Time `sphexa_ref --init sedov -n side -s (warmup+steps) --ascii` and parse its per-phase timer lines
    (main/src/propagator/ve_hydro.hpp:135-191). Hydro-step time = FindNeighbors + the six loops (no domain::sync)."""
    exe = REPO / "oracle" / "_ref" / "sphexa_ref"
    cores = os.cpu_count() or 1
    if not exe.exists():
        raise FileNotFoundError(f"{exe} missing (built by __graft_entry__.build() where /root/reference exists)")
    env = dict(os.environ, OMP_NUM_THREADS=str(cores), OMP_PROC_BIND="close")
    with tempfile.TemporaryDirectory() as tmp:
        out = subprocess.run([str(exe), "--init", "sedov", "-n", str(side), "-s", str(warmup + steps), "--ascii"],
                             cwd=tmp, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                             check=True).stdout
    iters, cur = [], {}
    for line in out.splitlines():
        m = re.match(r"# (.+?): ([0-9.eE+-]+)s", line)
        if m and m.group(1) in REF_PHASE_LABELS:
            cur[m.group(1)] = cur.get(m.group(1), 0.0) + float(m.group(2))
        if line.startswith("=== Total time for iteration"):
            iters.append(cur)
            cur = {}
    timed = iters[warmup:warmup + steps]
    if not timed:
        raise RuntimeError("no timed iterations parsed from sphexa_ref output")
    per_step = [sum(it.values()) for it in timed]
    t = sum(per_step) / len(per_step)
    n = side ** 3
    phases = {k: sum(it.get(k, 0.0) for it in timed) / len(timed) * 1e3 for k in REF_PHASE_LABELS}
    try:
        cpu_model = next(l.split(":")[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name"))
    except (OSError, StopIteration):
        cpu_model = "unknown"
    return {"value": n / t, "ms_per_step": t * 1e3, "cores": cores, "kind": "reference", "cpu_model": cpu_model,
            "sample": f"sphexa_ref --init sedov -n {side} -s {warmup + steps} --ascii ({n} particles, OpenMP "
                      f"{cores} threads, first {warmup} iterations dropped); step = FindNeighbors + 6 loops",
            "phases_ms": phases, "n_particles": n}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    side = args.ref_side
    r = run_reference_cpu(side, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64 (mixed, reference production types)",
            "data": "synthetic",
            "config": {"workload": f"Sedov blast wave {args.side or int(round(200 * args.gpus ** (1.0 / 3.0)))}^3 "
                                   f"(reference CPU arm runs the bounded sample "
                                   f"Sedov {side}^3, same per-particle work)", "sample_side": side},
            "cpu_baseline": {k: r[k] for k in ("value", "cores", "kind", "sample", "cpu_model")} | {"unit": UNIT},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "phases_ms": r["phases_ms"]}
    print(json.dumps(finite_json(line), allow_nan=False))


def measure_next_rows(sx, cases, side, dev, peak, reps=5) -> dict:
    """"report domain::sync and integrate separately". The callers of the hot path (8f rows 1-3) timed on
    their own Simulation object: device Domain::sync (keys, radix sort, octree), the field reorder, integrate
    (positions + energy + smoothing length, one fused pass) and the conserved-quantity reduction. Not part of `value`."""
    import torch
    s = cases.make_sedov_sim(sx, side, device=dev)
    s.step()  # sorts the lattice, gives every field a value
    torch.cuda.synchronize()
    names = ["domain_sync", "hydro_step", "conserved", "integrate"]
    acc = {k: 0.0 for k in names}
    for _ in range(reps):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        ev[0].record()
        s.sync()
        ev[1].record()
        s.compute_forces()
        ev[2].record()
        s.compute_conserved()
        ev[3].record()
        s.integrate()
        ev[4].record()
        torch.cuda.synchronize()
        for i, k in enumerate(names):
            acc[k] += ev[i].elapsed_time(ev[i + 1]) / reps
    n = s.n
    # integrate: reads x,y,z (24) x_m1.. (12) a (12) temp (8) du (8) du_m1 (4) h (4) nc (4) = 76 B,
    # writes x,y,z (24) x_m1.. (12) v (12) temp (8) du_m1 (4) h (4) = 64 B per particle
    integ_bytes = 140
    # conserved: x,y,z (24) v (12) m (4) temp (8) nc (4)
    cons_bytes = 52
    # sync: keys (read 24, write 8+4) + sort (4 passes x 2 x 12 B) + reorder of 15 fields (100 B read + 100 B write + 4)
    out = {"ms": acc, "n_particles": n, "tree_nodes": s.tree.num_nodes, "tree_leaves": s.tree.num_leaves,
           "integrate": {"bytes_per_particle": integ_bytes, "GBps": integ_bytes * n / (acc["integrate"] * 1e-3) / 1e9,
                         "frac": integ_bytes * n / (acc["integrate"] * 1e-3) / 1e9 / peak},
           "conserved": {"bytes_per_particle": cons_bytes, "GBps": cons_bytes * n / (acc["conserved"] * 1e-3) / 1e9,
                         "frac": cons_bytes * n / (acc["conserved"] * 1e-3) / 1e9 / peak},
           "note": "one rank; domain_sync = Hilbert keys + radix sort + octree build + reorder of 15 fields; "
                   "integrate = computeTimestep + fused positions/energy/h update; the reference's own CUDA build "
                   "needs 6.1 ms (domain::sync) + 0.5 ms (Timestep + UpdateQuantities) for the same (profiles/)"}
    del s
    torch.cuda.empty_cache()
    return out


def measure_dist_loop(sx, sdist, glob, rank, world, dev, dist, reps=4) -> dict:
    """N > 1: DistributedSimulation = the reference's whole main loop with the domain re-decomposed in every step
    (global cell histogram over NCCL, migration, halo discovery, local tree). Times sync / forces / conserved /
    integrate per step, max over ranks. Not part of `value` (domain::sync is reported separately)."""
    import torch
    ds = sdist.DistributedSimulation(sx.sim, glob, rank, world, dev)
    ds.step()
    ds.step()
    torch.cuda.synchronize()
    names = ["domain_sync", "hydro_step", "conserved", "integrate"]
    acc = torch.zeros(len(names), dtype=torch.float64, device=dev)
    for _ in range(reps):
        dist.barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        ev[0].record()
        ds.sync()
        ev[1].record()
        ds.compute_forces()
        ev[2].record()
        ds.compute_conserved()
        ev[3].record()
        ds.integrate()
        ev[4].record()
        torch.cuda.synchronize()
        acc += torch.tensor([ev[i].elapsed_time(ev[i + 1]) for i in range(4)], dtype=torch.float64, device=dev) / reps
    dist.all_reduce(acc, op=dist.ReduceOp.MAX)
    sizes = torch.tensor([ds.hd.last - ds.hd.first, ds.hd.n], dtype=torch.int64, device=dev)
    smax = sizes.clone()
    dist.all_reduce(smax, op=dist.ReduceOp.MAX)
    out = {"ms": dict(zip(names, acc.tolist())), "loop_ms_per_step": float(acc.sum()),
           "particles_per_sec_whole_loop": ds.n_global / (float(acc.sum()) * 1e-3),
           "cell_level": ds.level, "max_assigned_per_rank": int(smax[0]), "max_local_per_rank": int(smax[1]),
           "note": "dynamic SFC decomposition redone every step, all on the device: keys + radix sort, cell histogram + "
                   "ncclAllReduce, decomposition plan (sphx_cell_plan_build_device; only a 2 KB summary reaches the host), "
                   "slice migration (ncclSend/Recv), merge sort, halo exchange of x,y,z,h,m, local octree"}
    ds.close()
    return out


# ------------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------------
def our_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libsphx has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    if world > 1:
        # NCCL for the device-side plumbing of the bench (barriers, max-over-ranks), gloo for the host-side object
        # gathers of the domain decomposition; the halo exchange itself is libsphx's own NCCL communicator
        dist.init_process_group("cpu:gloo,cuda:nccl", device_id=torch.device(dev))

    import sphexa_b200 as sx
    from sphexa_b200 import cases
    from sphexa_b200 import dist as sdist

    sx.load()
    # weak scaling: 8 M particles per GPU. N = 1 is BASELINE config 1 (Sedov 200^3), N = 8 is config 4 (Sedov 400^3);
    # the global lattice is cut into N contiguous Hilbert-key ranges, each rank holds its range plus halos.
    side = args.side if args.side else int(round(200 * world ** (1.0 / 3.0)))
    t0 = time.time()
    dh = None
    if world == 1:
        hd = cases.make_sedov(sx, side, device=dev)
        n_assigned = hd.n
    else:
        glob = cases.sedov_global(side)
        dh = sdist.DistributedHydro(sx.sim, glob, rank, world, dev)
        hd = dh.hd
        n_assigned = dh.n_assigned
    setup_s = time.time() - t0
    n_global = side ** 3
    n = hd.n  # local particles including halos
    stream = torch.cuda.current_stream()

    # HydroVeProp::computeForces (ve_hydro.hpp:147-190): one C-ABI call per loop, halo exchanges in between
    def X(*names):
        return (lambda: dh.exchange(list(names))) if dh is not None else None

    seq = [("find_neighbors", lambda: hd.find_neighbors_sph(sync=False), None),
           ("xmass", hd.xmass, X("xm")),
           ("ve_def_gradh", hd.ve_def_gradh, None),
           ("eos", hd.eos, X("vx", "vy", "vz", "prho", "c", "kx")),
           ("iad_divv_curlv", lambda: hd.iad_divv_curlv(sync=False), X("c11", "c12", "c13", "c22", "c23", "c33", "divv")),
           ("av_switches", hd.av_switches, X("alpha")),
           ("momentum_energy", lambda: hd.momentum_energy(sync=False), None)]
    calls = []
    for name, fn, ex in seq:
        calls.append((name, fn))
        if ex is not None:
            calls.append(("halo_exchange", ex))
    num_exchanges = sum(1 for c in calls if c[0] == "halo_exchange")
    # reset-scalars, block search (standard + overflow instantiation), EOS, 5 x (work-counter reset + persistent loop
    # kernel), pack kernel per exchange
    launches_per_step = 14 + num_exchanges

    h0 = hd.f["h"].clone()
    alpha0 = hd.f["alpha"].clone()

    def one_step(events=None):
        # every step starts from the same state (h and alpha are in/out fields of the step)
        hd.f["h"].copy_(h0)
        hd.f["alpha"].copy_(alpha0)
        for i, (_, fn) in enumerate(calls):
            if events is not None:
                events[i].record(stream)
            fn()
        if events is not None:
            events[len(calls)].record(stream)

    for _ in range(max(args.warmup, 3)):
        one_step()
    torch.cuda.synchronize()

    # ---- device-resident timing -----------------------------------------------------------------------------------
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(len(calls) + 1)] for _ in range(args.steps)]
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    for k in range(args.steps):
        one_step(ev[k])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None

    step_ms = [ev[k][0].elapsed_time(ev[k][-1]) for k in range(args.steps)]
    total_ms = sum(step_ms)
    phase_ms = {}
    for i, (name, _) in enumerate(calls):
        phase_ms[name] = phase_ms.get(name, 0.0) + sum(ev[k][i].elapsed_time(ev[k][i + 1])
                                                       for k in range(args.steps)) / args.steps
    tmax = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms = float(tmax.item())
    ms_per_step = total_ms / args.steps
    value = n_global / (ms_per_step * 1e-3)

    # step result (also proves the step ran): neighbour statistics + time steps
    a = hd.args()
    import ctypes as C
    res = sx._cabi.SphxStepResult()
    sx._cabi.check(hd.L.sphx_momentum_energy(C.byref(a), C.byref(res)))

    bs = hd.block_stats()

    # ---- end-to-end: host buffers in, host results out, through the same C-ABI calls -----------------------------
    in_names = ["x", "y", "z", "h", "m", "temp", "vx", "vy", "vz", "alpha"]
    out_names = ["ax", "ay", "az", "du", "h", "nc"]
    host_in = {k: hd.f[k].cpu().pin_memory() for k in in_names}
    host_in["h"] = h0.cpu().pin_memory()
    host_in["alpha"] = alpha0.cpu().pin_memory()
    host_out = {k: torch.empty_like(hd.f[k], device="cpu").pin_memory() for k in out_names}
    h2d = sum(t.numel() * t.element_size() for t in host_in.values())
    d2h = sum(t.numel() * t.element_size() for t in host_out.values()) + C.sizeof(sx._cabi.SphxStepResult)

    # Every step has its own inputs in pinned host memory and returns its results to pinned host memory; consecutive
    # steps are independent batches (each restarts from the same h and alpha), so they are pipelined over two sets of
    # device buffers for the uploaded and downloaded fields: the upload of step k+1 (copy stream) and the download of
    # step k-1's results (third stream) overlap the kernels of step k. Inside a step the inputs are uploaded in the
    # order the loops consume them and every loop waits only for the fields it reads. The C ABI re-reads the field
    # pointers on every call, so switching sets is just handing it the other pointers. All copies of all steps,
    # including the first upload and the last download, are inside the timed region.
    first_use = {"find_neighbors": "h", "xmass": "m", "eos": "vz", "av_switches": "alpha"}
    up_stream, down_stream = torch.cuda.Stream(), torch.cuda.Stream()
    io_names = list(dict.fromkeys(in_names + out_names))
    sets = [{k: hd.f[k] for k in io_names}, {k: torch.empty_like(hd.f[k]) for k in io_names}]
    uploaded = [None, None]     # per set: events of the upload in flight
    set_free = [None, None]     # per set: its results have reached the host (the set may be overwritten)

    def upload(s):
        if set_free[s] is not None:
            up_stream.wait_event(set_free[s])
        else:
            up_stream.wait_stream(stream)  # whatever ran before the pipeline no longer reads the fields
        done = {}
        with torch.cuda.stream(up_stream):
            for k in in_names:
                sets[s][k].copy_(host_in[k], non_blocking=True)
                done[k] = up_stream.record_event()
        uploaded[s] = done

    def e2e_pipeline(nsteps):
        results = []
        upload(0)
        for k in range(nsteps):
            s = k % 2
            if k + 1 < nsteps:
                upload(1 - s)  # enqueued before this step's kernels: runs beside them
            hd.f.update(sets[s])
            done = uploaded[s]
            r = sx._cabi.SphxStepResult()
            for name, fn in calls[:-1]:
                if name in first_use:
                    stream.wait_event(done[first_use[name]])
                fn()  # the loops and, on N > 1, the NCCL halo exchanges between them
                if name == "find_neighbors":
                    down_stream.wait_event(stream.record_event())
                    with torch.cuda.stream(down_stream):
                        for f in ("h", "nc"):
                            host_out[f].copy_(sets[s][f], non_blocking=True)
            aa = hd.args()
            sx._cabi.check(hd.L.sphx_momentum_energy(C.byref(aa), C.byref(r)))  # synchronises, returns dt scalars
            down_stream.wait_event(stream.record_event())
            with torch.cuda.stream(down_stream):
                for f in ("ax", "ay", "az", "du"):
                    host_out[f].copy_(sets[s][f], non_blocking=True)
                set_free[s] = down_stream.record_event()
            results.append(r)
        stream.wait_stream(down_stream)
        stream.wait_stream(up_stream)
        return results

    e2e_pipeline(2)
    torch.cuda.synchronize()
    # the pipelined steps return what the device-resident step returned
    got = {k: host_out[k].clone() for k in ("ax", "du", "nc")}
    hd.f.update(sets[0])
    one_step()
    torch.cuda.synchronize()
    for k, v in got.items():
        assert torch.equal(v, hd.f[k].cpu()), f"e2e pipeline: {k} differs from the device-resident step"
    e2e_steps = max(4, min(args.steps, 10))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    set_free[0] = set_free[1] = None
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    e2e_pipeline(e2e_steps)
    e1.record(stream)
    torch.cuda.synchronize()
    hd.f.update(sets[0])
    e2e_ms = torch.tensor([e0.elapsed_time(e1) / e2e_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = n_global / (float(e2e_ms.item()) * 1e-3)

    # ---- N > 1: the dynamic decomposition (multi-rank Domain::sync) and the whole loop, reported separately ---------
    dist_rows = None
    if world > 1 and not args.no_next_rows:
        try:
            dist_rows = measure_dist_loop(sx, sdist, glob, rank, world, dev, dist)
        except Exception as e:  # noqa: BLE001
            dist_rows = {"error": repr(e)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    dom = max(PHASES, key=lambda k: phase_ms[k])
    n = n_assigned  # algorithmic bytes count the particles a rank computes
    algo = ALGO_BYTES[dom] * n
    achieved = algo / (phase_ms[dom] * 1e-3) / 1e9
    traffic = None
    tfile = REPO / "profiles" / "traffic.json"
    if tfile.exists():
        tj = json.loads(tfile.read_text())
        traffic = tj.get(f"{TRAFFIC_KEY.get(dom, dom)}@sedov{side}")
    mean_nc = res.totalNeighbors / n_assigned
    # bytes this design moves on top of the compulsory field traffic ("if the implementation materialises
    # neighbour lists in HBM, add the list term it actually needs"): 2 B per stored neighbour (16-bit block-local
    # indices) and one 16 B candidate record per block-local candidate, written once by the search and read once per
    # consuming loop (the two passes of IAD + divv/curlv read the list twice)
    cand_pp = bs["candTop"] / n
    list_b, cand_b = 2.0 * (mean_nc - 1), 16.0 * cand_pp
    design = {k: ALGO_BYTES[k] + (0 if k == "eos" else list_b * (2 if k == "iad_divv_curlv" else 1) + cand_b)
              for k in PHASES}

    # the bound that binds the neighbour kernels: warp-instruction issue (4 per SM and cycle). Instructions per launch
    # come from the committed ncu capture of the same command (profiles/traffic.json, "inst:<phase>@sedov<side>").
    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    issue_peak = sm_count * 4 * sm_mhz * 1e6  # warp instructions / s
    tj_all = json.loads(tfile.read_text()) if tfile.exists() else {}

    def per_kernel(k):
        t = phase_ms[k] * 1e-3
        out = {"ms": phase_ms[k], "GBps": ALGO_BYTES[k] * n / t / 1e9, "frac": ALGO_BYTES[k] * n / t / 1e9 / peak,
               "design_bytes_per_particle": design[k], "design_GBps": design[k] * n / t / 1e9,
               "design_frac": design[k] * n / t / 1e9 / peak}
        inst = tj_all.get(f"inst:{TRAFFIC_KEY.get(k, k)}@sedov{side}") if world == 1 else None
        if inst:
            out["warp_inst_per_launch"] = inst
            out["issue_frac"] = inst / t / issue_peak
        return out

    roofline = {"bound": "hbm", "kernel": KERNEL_OF[dom], "phase": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_particle": ALGO_BYTES[dom],
                "neighbor_list_bytes_per_particle": list_b, "candidate_bytes_per_particle": cand_b,
                "note": "achieved/frac: algorithmic bytes = compulsory field traffic only (SURVEY 8d); design_*: "
                        "compulsory + the 16-bit neighbour list and candidate records this design stores in HBM; "
                        "issue_frac: warp instructions (ncu capture under profiles/) / time / (4 per SM and cycle), the "
                        "bound that actually binds these kernels",
                "issue_peak_warp_inst_per_s": issue_peak,
                "per_kernel": {k: per_kernel(k) for k in PHASES}}
    pk = roofline["per_kernel"]
    if all("warp_inst_per_launch" in pk[k] for k in PHASES):
        roofline["issue_frac_step"] = sum(pk[k]["warp_inst_per_launch"] for k in PHASES) / (ms_per_step * 1e-3) / issue_peak

    next_rows = None
    if world == 1 and not args.no_next_rows:
        try:
            next_rows = measure_next_rows(sx, cases, side, dev, peak)
        except Exception as e:  # noqa: BLE001
            next_rows = {"error": str(e)}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            r = run_reference_cpu(args.ref_side, 8, 1)
            cpu_baseline = {k: r[k] for k in ("value", "cores", "kind", "sample", "cpu_model")} | {"unit": UNIT}
        except Exception as e:  # noqa: BLE001
            cpu_baseline = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference",
                            "sample": f"unavailable: {e}"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32+f64 (mixed, reference production types)", "data": "synthetic",
            "config": {"workload": f"Sedov blast wave {side}^3 ({n_global} particles), VE hydro step, ng0=100 ngmax=150, "
                                   f"periodic box", "particles_per_gpu": n_global // world,
                       "local_particles_rank0_incl_halos": hd.n,
                       "cache": "inputs (>3 GB of fields + neighbour list per GPU) exceed the 126 MB L2",
                       "parallelism": (f"SFC (Hilbert) domain decomposition over {world} GPUs, 4 NCCL halo exchanges "
                                       f"per step (send/recv), one process per GPU") if world > 1 else "single GPU"},
            "roofline": roofline, "cpu_baseline": cpu_baseline,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": float(e2e_ms.item())},
            "gpu_launches": launches_per_step * args.steps, "clocks": clocks, "phases_ms": phase_ms,
            "check": {"total_neighbors": int(res.totalNeighbors), "mean_nc": mean_nc, "max_nc": int(res.maxNc),
                      "minDtCourant": res.minDtCourant, "minDtRho": res.minDtRho,
                      "candidates_per_block_mean": float(bs["numCand"].mean()),
                      "candidates_per_block_max": int(bs["numCand"].max()),
                      "candidates_per_particle": bs["candTop"] / n, "fold_blocks": int((bs["flags"] & 1).sum())},
            "next_rows": next_rows if world == 1 else dist_rows, "setup_s": setup_s}
    print(json.dumps(finite_json(line), allow_nan=False))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--side", type=int, default=0,
                    help="global Sedov lattice side; default 200 * gpus^(1/3): 200 on 1 GPU, 400 on 8 (BASELINE configs)")
    ap.add_argument("--ref-side", type=int, default=128,
                    help="lattice side of the bounded CPU-reference sample (128^3 = 2.1 M particles, ~1.5 s/step on 16 cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-next-rows", action="store_true", help="skip timing Domain::sync / integrate / conserved")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        our_arm(args)


if __name__ == "__main__":
    main()
