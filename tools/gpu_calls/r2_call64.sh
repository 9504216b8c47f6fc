#!/bin/bash
# two GPUs at HEAD, the driver's command (full-size legs)
NP_BENCH_VERBOSE=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench n=2 rc=$?"; wc -c gpurun_out/r2_bench_n2.json
