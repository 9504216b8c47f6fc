#!/bin/bash
mkdir -p gpurun_out
for m in box sphere; do
  ncu --set full --clock-control none --import-source on -k regex:"k_gjk1|k_epa" -c 4 -o gpurun_out/r2n_pairs_$m -f python tools/prof_pairs.py $m 1000000 1 > gpurun_out/r2n_ncu_$m.log 2>&1
  tail -2 gpurun_out/r2n_ncu_$m.log
done
