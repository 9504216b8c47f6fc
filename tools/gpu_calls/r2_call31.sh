#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
NP_BENCH_VERBOSE=1 python bench.py > gpurun_out/r3b_bench.json 2> gpurun_out/r3b_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r3b_bench.err; wc -c gpurun_out/r3b_bench.json
