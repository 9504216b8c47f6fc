#!/bin/bash
mkdir -p gpurun_out
NP_BENCH_VERBOSE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "bench n=8 rc=$?"; tail -c 300 gpurun_out/r2_bench_n8.err; wc -c gpurun_out/r2_bench_n8.json
