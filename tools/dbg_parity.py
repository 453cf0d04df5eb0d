import sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np
import sphexa_b200 as sx
from refdata import load_golden
from test_gpu_parity import run_step_by_loops, field_floor, F32_FIELDS
for f in sys.argv[1:]:
    ref = load_golden(f)
    got, hd = run_step_by_loops(sx, ref)
    print(f, "h equal", np.array_equal(got["h"], ref["h"]), "nc equal", np.array_equal(got["nc"], ref["nc"]))
    for k in F32_FIELDS:
        a, b = got[k].astype(np.float64), ref[k].astype(np.float64)
        fl = field_floor(ref, k)
        err = np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), fl)
        i = err.argmax()
        print(f"  {k:6s} max|ref| {np.abs(b).max():.4e} floor {fl:.3e} maxrel {err.max():.3e} at {i}: {a[i]:.8e} vs {b[i]:.8e}  abs {np.abs(a-b).max():.3e}  rms-rel {np.sqrt((err**2).mean()):.2e}")
