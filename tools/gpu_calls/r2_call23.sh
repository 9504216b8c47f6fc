#!/bin/bash
mkdir -p gpurun_out
NP_G_EPA=1 ncu --set full --clock-control none --import-source on -k regex:"k_epa" -c 3 -o gpurun_out/r2u_pairs_box_g1 -f python tools/prof_pairs.py box 1000000 1 > gpurun_out/r2u_ncu.log 2>&1
tail -1 gpurun_out/r2u_ncu.log
