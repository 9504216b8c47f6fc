#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_comm_cpp.py tests/test_gpu_sharded.py -x -q 2>&1 | tail -3
NP_BENCH_VERBOSE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench n=2 rc=$?"; tail -c 400 gpurun_out/r2_bench_n2.err; wc -c gpurun_out/r2_bench_n2.json
