#!/bin/bash
# Runs ON THE GPU BOX (under gpurun): plain bench first, then the two ncu passes of B200_PROFILING.md on the same command.
set -x
CMD="python bench.py --steps 3 --warmup 3"
$CMD > gpurun_out/prof_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/prof_launches.csv $CMD > gpurun_out/prof_ncu1.log 2>&1
$CMD > gpurun_out/prof_plain2.log 2>&1 || exit 1
# world-step kernels of one timed step: skip warm-up steps (3 x 12 launches) then take GJK + the EPA tiers
ncu --set full --clock-control none --import-source on -k regex:"k_gjk|k_epa|k_integrate|k_world_verts" -s 12 -c 6 -o gpurun_out/prof_full $CMD > gpurun_out/prof_ncu2.log 2>&1
tail -2 gpurun_out/prof_ncu2.log
