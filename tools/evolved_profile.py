"""ncu target: evolve Sedov side^3 for `steps` steps of the native loop, then run the block search inside a
cudaProfilerStart/Stop bracket (ncu --profile-from-start off). usage: python tools/evolved_profile.py [side] [steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sphexa_b200 as sx  # noqa: E402
from sphexa_b200 import cases  # noqa: E402

side = int(sys.argv[1]) if len(sys.argv) > 1 else 128
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 600
s = cases.make_sedov_sim(sx, side)
for _ in range(steps):
    s.step()
s.sync()
s.hydro_step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
s.hydro_step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
