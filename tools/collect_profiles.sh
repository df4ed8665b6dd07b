#!/bin/bash
# Run on the GPU box (under gpurun): the round's evidence set.
#   gpurun_out/launches_rN.csv      every launch of a short bench run with its device time
#   gpurun_out/prof_rN_all.ncu-rep  one ncu --set full capture of each of our 3DGS kernels
#   gpurun_out/prof_rN_surfel.ncu-rep  the same for the surfel (2DGS) kernels
# usage: tools/collect_profiles.sh <round-tag>
tag=${1:-r2}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 120 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'project_kernel|tile_sort|blend_|gauss_backward' \
    -s 10 -c 5 -f -o gpurun_out/prof_${tag}_all python tools/prof_step.py --views 2 > gpurun_out/ncu_${tag}_all.log 2>&1
tail -2 gpurun_out/ncu_${tag}_all.log
ncu --set full --clock-control none --import-source on -k regex:'surfel_' \
    -s 4 -c 4 -f -o gpurun_out/prof_${tag}_surfel python tools/prof_surfel.py --views 2 > gpurun_out/ncu_${tag}_surfel.log 2>&1
tail -2 gpurun_out/ncu_${tag}_surfel.log
