#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2d_tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/r2d_tests.log
for cfg in "NP_SEED=0" "NP_SEED=1" "NP_SEED=1 NETPHYS_B200_LIB=/root/repo/gpurun_in/lib_mb5.so" "NP_SEED=0 NETPHYS_B200_LIB=/root/repo/gpurun_in/lib_mb5.so"; do
  echo "== $cfg"; env $cfg python tools/prof_world.py 1000000 6 c3 2>&1 | tail -1
done
echo "== grid"; python tools/prof_world.py 1000000 6 c3 2.0 grid 2>&1 | tail -1
echo "== box"; python tools/prof_world.py 1000000 6 box 2>&1 | tail -1
echo "== c2-1M"; python tools/prof_world.py 1000000 6 c2 2>&1 | tail -1
