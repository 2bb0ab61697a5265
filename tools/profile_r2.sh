#!/bin/bash
# ncu evidence of one round (run under gpurun, one GPU): tools/profile_r2.sh TAG
#   gpurun_out/TAG_plain.log          the command without a profiler (must exit 0 first)
#   gpurun_out/TAG_launches.csv       every launch with its device time (cold caches, serialised: compare shares)
#   gpurun_out/TAG_mgsolve.ncu-rep    ncu --set full of one k_mg_solve launch
#   gpurun_out/TAG_particles.ncu-rep  ncu --set full of the fused cell push (both species) and the mover insertion
tag=${1:-r2}
cmd="python bench.py --steps 2 --warmup 3 --e2e-steps 2 --no-cpu-baseline"
$cmd > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${tag}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${tag}_launches.csv $cmd > gpurun_out/${tag}_ncu0.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_mg_solve -s 6 -c 1 -o gpurun_out/${tag}_mgsolve -f $cmd > gpurun_out/${tag}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_cell_push|k_mv_insert' -s 12 -c 4 -o gpurun_out/${tag}_particles -f $cmd > gpurun_out/${tag}_ncu2.log 2>&1
ls -la gpurun_out/${tag}_*
