#!/bin/bash
# the driver's single-GPU command at HEAD
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err || tail -5 gpurun_out/r2_bench_n1.err
wc -c gpurun_out/r2_bench_n1.json
