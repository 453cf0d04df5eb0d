#!/bin/bash
# GPU-box helper: rebuild libsphx with different -D settings and bench each. usage: tools/gpu_variants.sh "<EXTRA1>" "<EXTRA2>" ...
for extra in "$@"; do
  rm -f sphexa_b200/csrc/build/search.o sphexa_b200/csrc/build/loops.o
  make -s -C sphexa_b200/csrc EXTRA="$extra" > /dev/null 2>&1 || { echo "build failed: $extra"; continue; }
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-next-rows > /tmp/b.json 2> /tmp/b.err
  python -c "
import json;d=json.load(open('/tmp/b.json'));print('$extra', round(d['ms_per_step'],3),{k:round(v['ms'],3) for k,v in d['roofline']['per_kernel'].items()})"
done
