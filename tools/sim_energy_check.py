"""GPU-box helper: native Simulation loop (Sedov side^3, N steps) vs the reference CPU energy series; prints the
max relative deviations per column and the per-stage timings. usage: python tools/sim_energy_check.py [side] [steps]"""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import sphexa_b200 as sx  # noqa: E402
from sphexa_b200 import cases  # noqa: E402

side = int(sys.argv[1]) if len(sys.argv) > 1 else 50
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
s = cases.make_sedov_sim(sx, side)
rows = []
t0 = time.time()
for k in range(steps):
    rows.append(s.step())
torch.cuda.synchronize()
print(f"side {side} steps {steps}: {time.time() - t0:.2f} s wall; nodes {s.tree.num_nodes} leaves {s.tree.num_leaves}")
got = np.array(rows, dtype=np.float64)
g = Path(__file__).resolve().parent.parent / "tests" / "golden" / f"sedov{side}_energies.npz"
if g.exists():
    ref = np.load(g)["series"][:steps]
    cols = "step ttot minDt etot ecin eint linmom angmom totalNeighbors".split()
    for c in range(1, 9):
        den = np.maximum(np.abs(ref[:, c]), 1e-300)
        rel = np.abs(got[: ref.shape[0], c] - ref[:, c]) / den
        print(f"  {cols[c]:15s} max rel dev {rel.max():.3e} at step {rel.argmax()}  (last: got {got[ref.shape[0]-1, c]:.10g} ref {ref[-1, c]:.10g})")
print("etot drift", got[-1, 3] / got[0, 3] - 1.0)
