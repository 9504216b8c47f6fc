#!/bin/bash
# Runs ON THE GPU BOX (under gpurun): the bench first (plain), then the two ncu passes of B200_PROFILING.md on the same command.
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3"
$CMD > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err || { tail -5 gpurun_out/r2_bench_n1.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -c 400 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu1.log 2>&1
# full capture of the narrow-phase + integrate + resolve kernels of the headline world's steps, L2 left as the step's earlier kernels left it
ncu --set full --clock-control none --cache-control none --graph-profiling node \
    -k regex:"k_gjk2|k_epa|k_integrate|k_resolve|k_link" -s 18 -c 9 -o gpurun_out/r2_full -f $CMD > gpurun_out/r2_ncu2.log 2>&1
tail -2 gpurun_out/r2_ncu2.log
# two GPUs are a separate call: tools/gpu_calls/r2_call_n2.sh
