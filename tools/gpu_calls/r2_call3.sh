#!/bin/bash
mkdir -p gpurun_out
for cfg in "NP_GJK1_OLD=1 NP_EPA_FIRST=0" "NP_EPA_FIRST=0" "NP_X=0"; do
  echo "== $cfg"; env $cfg python tools/prof_world.py 1000000 6 c3 2>&1 | tail -1
done
ncu --set full --clock-control none --cache-control none --graph-profiling node --import-source on -k regex:"k_gjk|k_epa" -s 8 -c 4 -o gpurun_out/r2c_full_1m python tools/prof_world.py 1000000 4 c3 > gpurun_out/r2c_ncu2.log 2>&1
tail -2 gpurun_out/r2c_ncu2.log
