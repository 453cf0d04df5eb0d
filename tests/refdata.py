"""Readers for reference-harness dumps (oracle/ref_harness.cpp) and the committed golden fixtures.

Test infrastructure only.
"""
from __future__ import annotations

import os
import subprocess
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
GOLDEN = REPO / "tests" / "golden"
REF_HARNESS = REPO / "oracle" / "_ref" / "ref_harness"

_DT = {"f8": np.float64, "f4": np.float32, "u4": np.uint32, "i4": np.int32, "u8": np.uint64}


def load_dump_dir(path) -> dict:
    """Load one <outdir>/step<k>/ directory written by ref_harness into {name: ndarray}."""
    path = Path(path)
    out = {}
    for line in (path / "manifest.txt").read_text().splitlines():
        name, dt, n = line.split()
        a = np.fromfile(path / f"{name}.bin", dtype=_DT[dt])
        assert a.size == int(n), (name, a.size, n)
        out[name] = a
    return out


def load_golden(name: str) -> dict:
    with np.load(GOLDEN / name) as z:
        return {k: z[k] for k in z.files}


def have_ref_harness() -> bool:
    return REF_HARNESS.exists() and os.access(REF_HARNESS, os.X_OK)


def run_ref_harness(case: str, n: int, steps: int, outdir, dump_every: int = 1, dump_neighbors: bool = True,
                    threads: int | None = None, hscale: float = 1.0, av_clean: bool = False,
                    stir: int = 0) -> list[dict]:
    """Run the compiled reference (oracle/_ref/ref_harness) and return one dict per dumped step."""
    outdir = Path(outdir)
    env = dict(os.environ)
    if threads:
        env["OMP_NUM_THREADS"] = str(threads)
    subprocess.run([str(REF_HARNESS), case, str(n), str(steps), str(outdir), str(dump_every),
                    "1" if dump_neighbors else "0", repr(float(hscale)), "1" if av_clean else "0", str(int(stir))], check=True,
                   env=env,
                   stdout=subprocess.PIPE)
    steps_out = []
    k = 0
    while (outdir / f"step{k}").exists() or k < steps:
        if (outdir / f"step{k}").exists():
            d = load_dump_dir(outdir / f"step{k}")
            d["_step"] = k
            steps_out.append(d)
        k += 1
        if k > steps:
            break
    return steps_out


def csr_sorted_neighbors(neighbors: np.ndarray, nc: np.ndarray, ngmax: int):
    """ngmax-strided reference lists -> (offsets, sorted CSR indices); nc includes self."""
    n = nc.size
    cnt = np.minimum(nc.astype(np.int64) - 1, ngmax)
    nb = neighbors.reshape(n, ngmax)
    mask = np.arange(ngmax)[None, :] < cnt[:, None]
    big = np.where(mask, nb, np.iinfo(np.uint32).max)
    big = np.sort(big, axis=1)
    offsets = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(cnt, out=offsets[1:])
    return offsets, big[mask.sum(axis=1)[:, None] > np.arange(ngmax)[None, :]].astype(np.uint32)
