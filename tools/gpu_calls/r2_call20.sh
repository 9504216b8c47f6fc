#!/bin/bash
mkdir -p gpurun_out
for t in 1 0; do
  echo "== NP_PREFETCH=$t"
  NP_PREFETCH=$t python tools/prof_world.py 1000000 6 c3 2>&1 | tail -1 | cut -c1-140
done
ncu --set full --clock-control none --import-source on -k regex:"k_gjk2" -c 1 -o gpurun_out/r2r_world1m -f python tools/prof_world.py 1000000 1 c3 > gpurun_out/r2r_ncu.log 2>&1
tail -1 gpurun_out/r2r_ncu.log
