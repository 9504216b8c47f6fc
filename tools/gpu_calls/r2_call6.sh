#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2f_tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/r2f_tests.log
python tools/prof_world.py 1000000 6 c3 2>&1 | tail -1
echo "== box"; python tools/prof_world.py 1000000 6 box 2>&1 | tail -1
echo "== c2-1M"; python tools/prof_world.py 1000000 6 c2 2>&1 | tail -1
python -c "
import netphys_b200 as npb
w=npb.World(0); print('peaks', w.fp32_peaks_tflops())"
ncu --set full --clock-control none --cache-control none --graph-profiling node --import-source on -k regex:"k_gjk|k_epa" -s 10 -c 5 -o gpurun_out/r2f_full_1m python tools/prof_world.py 1000000 4 c3 > gpurun_out/r2f_ncu2.log 2>&1
