#!/bin/bash
# final single-GPU pass at HEAD: the GPU suite, the driver's bench command, the reference arm's line, the launch list of the bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err || { tail -5 gpurun_out/r2_bench_n1.err; }
wc -c gpurun_out/r2_bench_n1.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/r2_ncu1.log 2>&1
wc -l gpurun_out/r2_launches.csv
