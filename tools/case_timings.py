"""Per-phase timing of the hydro step on the other BASELINE cases (single GPU): Noh 150^3 (jittered lattice cut to a
sphere, open box, v = -r/|r|), turbulence box 200^3 (jittered lattice, periodic, imposed solenoidal velocity field), next to
the Sedov lattice. Each step restarts from the same h and alpha. usage: python tools/case_timings.py [out.json]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sphexa_b200 as sx  # noqa: E402
from sphexa_b200 import cases  # noqa: E402


def time_case(name, hd, steps=5, warmup=2):
    seq = [("find_neighbors", lambda: hd.find_neighbors_sph(sync=False)), ("xmass", hd.xmass),
           ("ve_def_gradh", hd.ve_def_gradh), ("eos", hd.eos), ("iad_divv_curlv", lambda: hd.iad_divv_curlv(sync=False)),
           ("av_switches", hd.av_switches), ("momentum_energy", lambda: hd.momentum_energy(sync=False))]
    h0, a0 = hd.f["h"].clone(), hd.f["alpha"].clone()
    hd.hydro_step()  # converge h once: the timed steps start from the converged smoothing lengths
    h0 = hd.f["h"].clone()
    stream = torch.cuda.current_stream()
    acc = {k: 0.0 for k, _ in seq}
    for it in range(warmup + steps):
        hd.f["h"].copy_(h0), hd.f["alpha"].copy_(a0)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(seq) + 1)]
        for i, (_, fn) in enumerate(seq):
            ev[i].record(stream)
            fn()
        ev[-1].record(stream)
        torch.cuda.synchronize()
        if it >= warmup:
            for i, (k, _) in enumerate(seq):
                acc[k] += ev[i].elapsed_time(ev[i + 1]) / steps
    hd.momentum_energy()
    bs = hd.block_stats()
    n = hd.last - hd.first
    tot = sum(acc.values())
    return {"case": name, "particles": n, "ms_per_step": tot, "particles_per_sec": n / (tot * 1e-3),
            "phases_ms": acc, "mean_nc": hd.result.totalNeighbors / n, "max_nc": hd.result.maxNc,
            "h_iterated": hd.result.numHIterated,
            "candidates_per_block": {"mean": float(bs["numCand"].mean()), "max": int(bs["numCand"].max())},
            "fold_blocks": int((bs["flags"] & 1).sum()), "precise_blocks": int(bs["precise"].sum()),
            "leaves_per_block": {"mean": float(bs["numLeaves"].mean()), "max": int(bs["numLeaves"].max())},
            "tiles_per_block": {"mean": float(bs["numTiles"].mean()), "max": int(bs["numTiles"].max())}}


if __name__ == "__main__":
    out = []
    for name, make in (("sedov 200^3", lambda: cases.make_sedov(sx, 200)), ("noh 150^3", lambda: cases.make_noh(sx, 150)),
                       ("turbulence 200^3", lambda: cases.make_turbulence(sx, 200))):
        hd = make()
        r = time_case(name, hd)
        out.append(r)
        print(json.dumps(r))
        del hd
        torch.cuda.empty_cache()
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"), indent=1)
