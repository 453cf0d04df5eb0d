"""Per-stage wall times of DistributedSimulation.sync (multi-rank Domain::sync). Launch with torchrun on N GPUs."""
import json
import os
import sys

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
side = int(sys.argv[1]) if len(sys.argv) > 1 else int(round(200 * world ** (1 / 3)))
glob = cases.sedov_global(side)
ds = sdist.DistributedSimulation(sx.sim, glob, rank, world, dev)
ds.step(), ds.step()
ds.profile = {}
reps = 4
for _ in range(reps):
    ds.step()
if rank == 0:
    print(json.dumps({"side": side, "world": world, "level": ds.level,
                      "sync_ms": {k: round(1e3 * v / reps, 2) for k, v in ds.profile.items()},
                      "total_ms": round(1e3 * sum(ds.profile.values()) / reps, 2)}))
