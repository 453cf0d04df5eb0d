"""Where do the comparison floors of du, ax, ay, az come from (tests/test_gpu_parity.py FLOOR_FRACTION)?
For live dumps of the compiled reference: (1) the reference's OWN fp32 noise = |reference (production fp32) - the same
loop in all-double on the same inputs| (oracle.momentum_fields_d), (2) the error of the CUDA path against the same
all-double values, both as a fraction of the field family's max-norm, plus the unfloored relative errors.
Run on the GPU box: python tools/floor_analysis.py > gpurun_out/floor_analysis.json"""
import json
import pathlib
import sys
import tempfile

import numpy as np

REPO = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))
import oracle_lib as O  # noqa: E402
from refdata import run_ref_harness  # noqa: E402
import sphexa_b200 as sx  # noqa: E402

out = []
for case, n, steps, hs in (("sedov", 64, 1, 1.0), ("sedov", 50, 3, 1.0), ("noh", 40, 2, 1.0), ("turb", 32, 2, 1.4)):
    with tempfile.TemporaryDirectory() as t:
        d = run_ref_harness(case, n, steps, pathlib.Path(t) / "o", hscale=hs)[-1]
    r64 = O.momentum_fields_d(d)
    hd = sx.sim.from_dump(d)
    hd.hydro_step()
    fam = max(np.abs(r64[k]).max() for k in ("ax", "ay", "az"))
    for k in ("ax", "ay", "az", "du"):
        ref32, exact, gpu = d[k].astype(np.float64), r64[k], hd.get(k).astype(np.float64)
        scale = np.abs(exact).max() if k == "du" else fam
        if scale == 0:
            continue
        noise, err, pair = np.abs(ref32 - exact) / scale, np.abs(gpu - exact) / scale, np.abs(gpu - ref32) / scale
        big = np.abs(exact) > 0.1 * scale
        rel = lambda a, b: np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), 1e-300)  # noqa: E731
        q = lambda v: [float(np.percentile(v, p)) for p in (50, 99, 99.9, 100)]  # noqa: E731
        out.append({"case": f"{case} {n}^3 step {d['_step']}", "field": k, "scale": float(scale),
                    "pct": [50, 99, 99.9, 100],
                    "reference_fp32_vs_fp64_over_scale": q(noise), "gpu_vs_fp64_over_scale": q(err),
                    "gpu_vs_reference_fp32_over_scale": q(pair),
                    "unfloored_rel_gpu_vs_reference": q(rel(gpu, ref32)),
                    "unfloored_rel_reference_fp32_vs_fp64": q(rel(ref32, exact)),
                    "rel_gpu_vs_reference_where_value_above_0.1_scale": q(rel(gpu, ref32)[big]) if big.any() else None,
                    "n": int(ref32.size)})
print(json.dumps(out, indent=1))
