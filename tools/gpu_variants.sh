#!/bin/bash
# GPU-box helper: rebuild libsphx with different -D settings and time the hydro step with each.
# usage: tools/gpu_variants.sh "<EXTRA1>" "<EXTRA2>" ...   ("" = defaults). CASES="sedov noh" selects the cases.
for extra in "$@"; do
  rm -f sphexa_b200/csrc/build/search.o sphexa_b200/csrc/build/loops.o
  make -s -C sphexa_b200/csrc EXTRA="$extra" > /dev/null 2>&1 || { echo "build failed: $extra"; cat sphexa_b200/csrc/build/loops.ptxas.log | grep -i error | head; continue; }
  for c in ${CASES:-sedov}; do
    python bench.py --case $c --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline --no-static > /tmp/b.json 2> /tmp/b.err || tail -3 /tmp/b.err
    python -c "
import json;d=json.load(open('/tmp/b.json'));print('[$extra] $c', round(d['ms_per_step'],3), 'median', round(d['ms_per_step_median'],3), {k:round(v,3) for k,v in d['phases_ms'].items()})"
  done
done
rm -f sphexa_b200/csrc/build/search.o sphexa_b200/csrc/build/loops.o
make -s -C sphexa_b200/csrc > /dev/null 2>&1
