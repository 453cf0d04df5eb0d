#!/bin/bash
# GPU-box helper: rebuild libsphx with different -D settings and time the search on an evolved state.
for extra in "$@"; do
  rm -f sphexa_b200/csrc/build/search.o sphexa_b200/csrc/build/loops.o
  make -s -C sphexa_b200/csrc EXTRA="$extra" > /dev/null 2>&1 || { echo "build failed: $extra"; continue; }
  python tools/evolved_timing.py 128 600 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$extra', d['case'][:30], round(d['ms_per_step'],2), round(d['phases_ms']['find_neighbors'],3))"
done
