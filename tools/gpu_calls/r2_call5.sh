#!/bin/bash
mkdir -p gpurun_out
python tools/prof_world.py 1000000 6 c3 2>&1 | tail -1
ncu --set full --clock-control none --cache-control none --graph-profiling node --import-source on -k regex:"k_gjk|k_epa" -s 10 -c 5 -o gpurun_out/r2e_full_1m python tools/prof_world.py 1000000 4 c3 > gpurun_out/r2e_ncu2.log 2>&1
tail -2 gpurun_out/r2e_ncu2.log
