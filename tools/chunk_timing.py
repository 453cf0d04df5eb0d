"""Cost of the candidate-chunk path of the loops (test hook sphx_debug_candidate_chunk): Sedov 128^3 with every block
staged in two chunks against the unchunked step. usage: python tools/chunk_timing.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import sphexa_b200 as sx
from sphexa_b200 import cases
from case_timings import time_case
hd = cases.make_sedov(sx, 128)
for chunk in (0, 512):
    sx.load().sphx_debug_candidate_chunk(chunk)
    r = time_case(f"sedov 128^3 chunk={chunk}", hd, steps=3, warmup=1)
    print(chunk, round(r["ms_per_step"], 2), {k: round(v, 2) for k, v in r["phases_ms"].items()})
sx.load().sphx_debug_candidate_chunk(0)
