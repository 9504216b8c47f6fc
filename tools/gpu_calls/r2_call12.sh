#!/bin/bash
mkdir -p gpurun_out
NP_BENCH_VERBOSE=1 python bench.py --steps 20 --warmup 5 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench rc=$?"; tail -c 800 gpurun_out/r2k_bench.err
python tools/prof_world.py 10000 30 c2 2>&1 | tail -1
python tools/prof_world.py 131072 10 c2 2>&1 | tail -1
