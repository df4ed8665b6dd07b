#!/bin/bash
# Run on the GPU box (under gpurun): the round's evidence set.
#   gpurun_out/launches_<tag>.csv            every launch of a short bench run with its device time (ours)
#   gpurun_out/launches_<tag>_reference.csv  the same for the unmodified reference (bench.py --impl reference)
#   gpurun_out/prof_<tag>_all.ncu-rep        one ncu --set full capture of each of our 3DGS kernels (bench workload)
#   gpurun_out/prof_<tag>_c5.ncu-rep         the same at 2M Gaussians / 1600^2 (BASELINE configs[4])
#   gpurun_out/prof_<tag>_surfel.ncu-rep     the same for the surfel (2DGS) kernels
# usage: tools/collect_profiles.sh <round-tag>
tag=${1:-r2}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 120 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_${tag}.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file gpurun_out/launches_${tag}_reference.csv \
    python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/launches_${tag}_reference.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'project_kernel|tile_sort|blend_|gauss_backward' \
    -s 10 -c 5 -f -o gpurun_out/prof_${tag}_all python tools/prof_step.py --views 2 > gpurun_out/ncu_${tag}_all.log 2>&1
tail -2 gpurun_out/ncu_${tag}_all.log
ncu --set full --clock-control none -k regex:'project_kernel|tile_sort|blend_|gauss_backward' \
    -s 10 -c 5 -f -o gpurun_out/prof_${tag}_c5 python tools/prof_step.py --views 2 --gaussians 2000000 --res 1600 > gpurun_out/ncu_${tag}_c5.log 2>&1
tail -2 gpurun_out/ncu_${tag}_c5.log
ncu --set full --clock-control none --import-source on -k regex:'surfel_' \
    -s 4 -c 4 -f -o gpurun_out/prof_${tag}_surfel python tools/prof_surfel.py --views 2 > gpurun_out/ncu_${tag}_surfel.log 2>&1
tail -2 gpurun_out/ncu_${tag}_surfel.log
