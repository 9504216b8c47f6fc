#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2b_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r2b_tests.log
for cfg in "NP_GJK1_OLD=1 NP_EPA_FIRST=0" "NP_GJK1_OLD=1" "NP_EPA_FIRST=0" "NP_CBIG=0" "NP_X=0"; do
  echo "== $cfg"; env $cfg python tools/prof_world.py 1000000 6 c3 2>&1 | tail -1
done
for cfg in "NP_GJK1_OLD=1 NP_EPA_FIRST=0" "NP_X=0"; do
  echo "== grid $cfg"; env $cfg python tools/prof_world.py 1000000 6 c3 2.0 grid 2>&1 | tail -1
  echo "== 4096x256-like 1M c2 $cfg"; env $cfg python tools/prof_world.py 1000000 6 c2 2>&1 | tail -1
  echo "== box $cfg"; env $cfg python tools/prof_world.py 1000000 6 box 2>&1 | tail -1
done
ncu --set full --clock-control none --cache-control none --graph-profiling node --import-source on -k regex:"k_gjk|k_epa|k_integrate|k_resolve|k_link|k_world" -s 16 -c 8 -o gpurun_out/r2b_full_1m python tools/prof_world.py 1000000 4 c3 > gpurun_out/r2b_ncu2.log 2>&1
tail -2 gpurun_out/r2b_ncu2.log
