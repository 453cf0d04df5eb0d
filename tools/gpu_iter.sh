#!/bin/bash
# GPU-box helper for one tuning iteration: parity tests, a short bench, optionally an ncu --set full capture.
# usage: tools/gpu_iter.sh <tag> [kernel-regex for ncu]
tag=$1; pat=$2
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python -c "
import json;d=json.load(open('gpurun_out/bench_$tag.json'));print(d['ms_per_step'],{k:round(v,3) for k,v in d['phases_ms'].items()},d['check'])"
tail -3 gpurun_out/bench_$tag.err
if [ -n "$pat" ]; then
  ncu --set full --clock-control none --import-source on -k regex:$pat -s ${3:-3} -c ${4:-1} -o gpurun_out/prof_$tag python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
fi
