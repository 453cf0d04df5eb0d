"""Per-phase timing of the hydro step on an EVOLVED state (Sedov n^3 after k steps of the native loop): the blast wave has
density contrasts, non-zero velocities and a spread of neighbour counts that the initial lattice does not have.
usage: python tools/evolved_timing.py [side=128] [steps=300]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import sphexa_b200 as sx  # noqa: E402
from sphexa_b200 import cases  # noqa: E402
from case_timings import time_case  # noqa: E402

side = int(sys.argv[1]) if len(sys.argv) > 1 else 128
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
s = cases.make_sedov_sim(sx, side)
out = []
for k in range(steps + 1):
    if k in (0, steps // 3, steps):
        s.sync()
        r = time_case(f"sedov {side}^3 after {k} steps (t = {s.p.ttot:.3e})", s, steps=3, warmup=1)
        out.append(r)
        print(json.dumps(r))
    if k < steps:
        s.step()
