import sys, json
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tools')
import sphexa_b200 as sx
from sphexa_b200 import cases
from case_timings import time_case
hd = cases.make_sedov(sx, 128)
for chunk in (0, 512):
    sx.load().sphx_debug_candidate_chunk(chunk)
    r = time_case(f"sedov 128^3 chunk={chunk}", hd, steps=3, warmup=1)
    print(chunk, round(r["ms_per_step"], 2), {k: round(v, 2) for k, v in r["phases_ms"].items()})
sx.load().sphx_debug_candidate_chunk(0)
