"""Long multi-GPU run of the native loop with the dynamic decomposition (torchrun, N ranks): the energy series of the
N-rank run against the single-rank Simulation run on rank 0, and the per-rank load balance over time.
usage: torchrun --nproc-per-node N tools/long_run_dist.py [side=100] [steps=600] [case=sedov]"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sphexa_b200 as sx  # noqa: E402
from sphexa_b200 import cases, dist as sdist  # noqa: E402

world, rank, lr = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = f"cuda:{lr}"
if world > 1:
    dist.init_process_group("cpu:gloo,cuda:nccl", device_id=torch.device(dev))
side = int(sys.argv[1]) if len(sys.argv) > 1 else 100
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 600
case = sys.argv[3] if len(sys.argv) > 3 else "sedov"
glob = {"sedov": cases.sedov_global, "noh": cases.noh_global, "turbulence": cases.turbulence_global}[case](side)
ds = sdist.DistributedSimulation(sx.sim, glob, rank, world, dev)
rows, loads = [], []
for k in range(steps):
    rows.append(ds.step())
    if k % max(1, steps // 6) == 0 or k == steps - 1:
        n = torch.tensor([ds.hd.last - ds.hd.first, ds.hd.n], dtype=torch.int64, device=dev)
        mx, mn = n.clone(), n.clone()
        if world > 1:
            dist.all_reduce(mx, op=dist.ReduceOp.MAX), dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        loads.append({"step": k, "assigned_min": int(mn[0]), "assigned_max": int(mx[0]), "local_max": int(mx[1]),
                      "cell_level": ds.level})
rows = np.array(rows, dtype=np.float64)

# halo completeness on the evolved state: one more sync + search on N ranks, then the same search on one rank over the
# gathered particle set; the search is bit-exact, so nc and h must agree particle by particle
ds.sync()
mine = {k: ds.assigned(k).copy() for k in ("id", "x", "y", "z", "h")}
ds.compute_forces()
mine["nc"], mine["h_out"] = ds.assigned("nc").copy(), ds.assigned("h").copy()
parts = [None] * world
if world > 1:
    dist.all_gather_object(parts, mine)
else:
    parts = [mine]
halo_check = None
if rank == 0:
    from sphexa_b200 import host
    from sphexa_b200.sim import HydroData
    g = {k: np.concatenate([p[k] for p in parts]) for k in mine}
    t = host.build_tree(g["x"], g["y"], g["z"], ds.box_lim, ds.boundary, bucket_size=64)
    o = t.order
    hd = HydroData(g["x"].size, 0, g["x"].size, ds.box_lim, ds.boundary, ds.p, device=dev)
    hd.set_fields(x=g["x"][o], y=g["y"][o], z=g["z"][o], h=g["h"][o], m=np.full(o.size, 1.0 / o.size, np.float32))
    hd.set_tree(t)
    hd.find_neighbors_sph()
    halo_check = {"particles": int(o.size), "nc_mismatches": int((hd.get("nc") != g["nc"][o]).sum()),
                  "h_mismatches": int((hd.get("h") != g["h_out"][o]).sum())}
if rank == 0:
    make = {"sedov": cases.make_sedov_sim, "noh": cases.make_noh_sim, "turbulence": cases.make_turbulence_sim}[case]
    s = make(sx, side, device=dev)
    ref = np.array([s.step() for _ in range(steps)], dtype=np.float64)
    rel = lambda a, b: float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))  # noqa: E731
    print(json.dumps({"case": case, "side": side, "world": world, "steps": steps, "t_end": rows[-1, 1],
                      "max_rel_diff_vs_single_rank": {"dt": rel(rows[:, 2], ref[:, 2]), "etot": rel(rows[:, 3], ref[:, 3]),
                                                      "ecin": rel(rows[1:, 4], ref[1:, 4]), "eint": rel(rows[:, 5], ref[:, 5])},
                      "neighbour_totals_equal_steps": int(np.sum(rows[:, 8] == ref[:, 8])),
                      "etot_first_last": [rows[0, 3], rows[-1, 3]], "halo_check_on_final_state": halo_check,
                      "load": loads}))
ds.close()
