#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_gjk1|k_epa" -c 5 -o gpurun_out/r2p_world1m -f python tools/prof_world.py 1000000 1 c3 > gpurun_out/r2p_ncu.log 2>&1
tail -2 gpurun_out/r2p_ncu.log
