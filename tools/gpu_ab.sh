#!/bin/bash
# GPU-box helper: A/B of two prebuilt libraries on the same box. usage: tools/gpu_ab.sh <tag> [libA libB ...]
# (libs are file names under sphexa_b200/, default: libsphx_base.so libsphx_new.so); CASES="sedov:200 turbulence:200"
tag=$1; shift
libs=${@:-"libsphx_base.so libsphx_new.so"}
cd sphexa_b200 && cp libsphx.so libsphx_new.so && cd ..
for lib in $libs; do
  cp sphexa_b200/$lib sphexa_b200/libsphx.so
  for cs in ${CASES:-sedov:200 turbulence:200}; do
    c=${cs%%:*}; n=${cs##*:}
    timeout 600 python bench.py --case $c --side $n --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline --no-static > /tmp/b.json 2> /tmp/b.err || tail -3 /tmp/b.err
    python -c "
import json;d=json.load(open('/tmp/b.json'));print('[$lib] $c $n', round(d['ms_per_step'],3), 'median', round(d['ms_per_step_median'],3), {k:round(v,3) for k,v in d['phases_ms'].items()}, 'mean_nc', d['check']['mean_nc'], 'etot', d['check']['etot'])" | tee -a gpurun_out/${tag}_ab.log
  done
done
cp sphexa_b200/libsphx_new.so sphexa_b200/libsphx.so
