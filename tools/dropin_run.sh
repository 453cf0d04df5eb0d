#!/bin/bash
# GPU-box helper: run the reference's own CUDA build (oracle/_ref/sphexa_cuda_ref), the same application on libsphx
# (sphexa_cuda_sphx) and print per-phase timer medians + final energies. usage: tools/dropin_run.sh <side> <steps>
side=${1:-100}; steps=${2:-6}
for exe in sphexa_cuda_ref sphexa_cuda_sphx; do
  d=$(mktemp -d); ( cd $d && /root/repo/oracle/_ref/$exe --init sedov -n $side -s $steps --ascii > log.txt 2>&1; echo "== $exe rc=$? side=$side steps=$steps"
  python3 - <<PY
import re, statistics as st
ph = {}
for line in open("log.txt"):
    m = re.match(r"# (.+?): ([0-9.eE+-]+)s", line)
    if m: ph.setdefault(m.group(1), []).append(float(m.group(2)))
tot = 0.0
for k, v in ph.items():
    if k.startswith("Total execution"): continue
    med = st.median(v[1:]) if len(v) > 1 else v[0]
    if k in ("FindNeighbors","XMass","Normalization & Gradh","EquationOfState","IadVelocityDivCurl","AVswitches","MomentumAndEnergy"): tot += med
    print(f"   {k:28s} median {med*1e3:9.3f} ms  (n={len(v)})")
print(f"   hydro step (FindNeighbors + 6 loops) {tot*1e3:.3f} ms -> {$side**3/tot/1e6:.1f} M particles/s")
PY
  tail -1 constants.txt ) ; rm -rf $d
done
