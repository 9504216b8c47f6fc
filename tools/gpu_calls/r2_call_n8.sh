#!/bin/bash
mkdir -p gpurun_out
for n in 8 4; do
NP_BENCH_VERBOSE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r2j_bench_n$n.json 2> gpurun_out/r2j_bench_n$n.err; echo "bench n=$n rc=$?"; tail -c 600 gpurun_out/r2j_bench_n$n.err; wc -c gpurun_out/r2j_bench_n$n.json
done
free -g | head -2
