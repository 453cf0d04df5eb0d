"""Long native run (sedov | noh | turbulence): watches the per-block search tables (leaves in reach, precise walks) as the blast wave evolves.
usage: python tools/long_run.py [side=100] [steps=2000] [every=250] [case=sedov]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sphexa_b200 as sx  # noqa: E402
from sphexa_b200 import cases  # noqa: E402

side = int(sys.argv[1]) if len(sys.argv) > 1 else 100
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
every = int(sys.argv[3]) if len(sys.argv) > 3 else 250
case = sys.argv[4] if len(sys.argv) > 4 else "sedov"
s = {"sedov": cases.make_sedov_sim, "noh": cases.make_noh_sim, "turbulence": cases.make_turbulence_sim}[case](sx, side)
t0 = time.time()
for k in range(steps):
    try:
        row = s.step()
    except sx.SphxError as e:
        print(json.dumps({"step": k, "error": str(e)}))
        break
    if k % every == 0 or k == steps - 1:
        bs = s.block_stats()
        print(json.dumps({"step": k, "t": row[1], "dt": row[2], "etot": row[3], "ecin": row[4],
                          "max_nc": int(s.result.maxNc), "leaves_max": int(bs["numLeaves"].max()),
                          "leaves_mean": float(bs["numLeaves"].mean()), "precise_blocks": int(bs["precise"].sum()),
                          "cand_max": int(bs["numCand"].max()), "tiles_max": int(bs["numTiles"].max()),
                          "wall_s": round(time.time() - t0, 1)}), flush=True)
