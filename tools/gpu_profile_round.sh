#!/bin/bash
# GPU-box helper: the measurement set committed under profiles/ for one kernel version.
# usage: tools/gpu_profile_round.sh <tag>   (outputs go to gpurun_out/<tag>_*)
tag=$1
timeout 900 python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference_cpu.json 2> gpurun_out/${tag}_bench_reference_cpu.err
# launch list of two steps (per-launch times under ncu are cold-cache and serialised: shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"loopKernel|blockSearchKernel|eosKernel" -s 27 -c 18 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-static > /dev/null 2>&1
# full capture of one step's kernels (search x2, five loops, eos)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"loopKernel|blockSearch|eosKernel" -s 16 -c 8 -o gpurun_out/${tag}_full \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-static > gpurun_out/${tag}_ncu.log 2>&1
timeout 600 tools/dropin_run.sh 200 6 > gpurun_out/${tag}_dropin_sedov200_refcuda_vs_sphx.log 2>&1
tail -c 600 gpurun_out/${tag}_bench_n1.json; echo; grep "hydro step\|==" gpurun_out/${tag}_dropin_sedov200_refcuda_vs_sphx.log
