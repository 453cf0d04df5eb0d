"""GPU-side diagnostic (run under gpurun): stage-by-stage comparison of the CUDA path with golden fixtures, with
neighbour-set mismatch details and candidate statistics. Not a test; prints, never asserts."""
import sys
import traceback

sys.path.insert(0, "tests")
sys.path.insert(0, ".")
import numpy as np

import sphexa_b200 as sx
from refdata import csr_sorted_neighbors, load_golden
from test_gpu_parity import F32_FIELDS, field_floor


def check(fname):
    ref = load_golden(fname)
    hd = sx.sim.from_dump(ref)
    ngmax = hd.p.ngmax
    try:
        hd.find_neighbors_xmass()
    except Exception as e:  # noqa: BLE001
        print("  search raised:", e)
    bs = hd.block_stats()
    print(f"{fname}: n={hd.n} blocks={bs['numCand'].size} cand mean {bs['numCand'].mean():.1f} max {bs['numCand'].max()} "
          f"candTop {bs['candTop']} cap {bs['candCapacity']} fold {int((bs['flags'] & 1).sum())} err {bs['errFlags']}")
    got_nc, got_h = hd.get("nc"), hd.get("h")
    print("  h equal", np.array_equal(got_h, ref["h"]), " nc equal", np.array_equal(got_nc, ref["nc"]),
          " nc mismatches", int((got_nc != ref["nc"]).sum()))
    if not np.array_equal(got_nc, ref["nc"]):
        bad = np.nonzero(got_nc != ref["nc"])[0][:10]
        for i in bad:
            print(f"    i={i} got {got_nc[i]} ref {ref['nc'][i]} block {i // 128}")
    else:
        nb = hd.export_neighbors()
        off, idx = csr_sorted_neighbors(nb, got_nc, ngmax)
        same = np.array_equal(idx, ref["nb_sorted"])
        print("  neighbour sets equal", same)
        if not same:
            d = np.nonzero(idx != ref["nb_sorted"])[0]
            print("    first diffs at csr positions", d[:10], "of", idx.size)
            for pos in d[:5]:
                i = np.searchsorted(off, pos, side="right") - 1
                print(f"    target {i}: got {idx[off[i]:off[i+1]][:12]} ref {ref['nb_sorted'][off[i]:off[i+1]][:12]}")
    stages = [("ve_def_gradh", hd.ve_def_gradh), ("eos", hd.eos), ("iad_divv_curlv", hd.iad_divv_curlv),
              ("av_switches", hd.av_switches), ("momentum_energy", hd.momentum_energy)]
    for name, fn in stages:
        try:
            fn()
        except Exception as e:  # noqa: BLE001
            print("  ", name, "raised:", e)
    for k in F32_FIELDS:
        a, b = hd.get(k).astype(np.float64), ref[k].astype(np.float64)
        fl = field_floor(ref, k)
        err = np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), fl)
        i = int(np.nanargmax(err)) if np.isfinite(err).any() else 0
        flag = "OK " if np.nanmax(err) <= 1e-4 and np.isfinite(a).all() else "BAD"
        print(f"  {flag} {k:6s} max|ref| {np.abs(b).max():.4e} maxrel {np.nanmax(err):.3e} at {i}: {a[i]:.8e} vs "
              f"{b[i]:.8e} rms {np.sqrt(np.nanmean(err ** 2)):.2e} nan {int(np.isnan(a).sum())}")
    print(f"  dts got {hd.result.minDtCourant:.8e} {hd.result.minDtRho:.8e} ref {ref['dts']}")


if __name__ == "__main__":
    for f in sys.argv[1:]:
        try:
            check(f)
        except Exception:  # noqa: BLE001
            traceback.print_exc()
