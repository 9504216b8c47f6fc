#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2i_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2i_tests.log
NP_BENCH_VERBOSE=1 NP_BENCH_BODIES=200000 NP_BENCH_SHARDED_BODIES_PER_GPU=200000 python -X faulthandler -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2i_bench_n2_small.json 2> gpurun_out/r2i_bench_n2_small.err; echo "bench small rc=$?"; tail -c 2500 gpurun_out/r2i_bench_n2_small.err; wc -c gpurun_out/r2i_bench_n2_small.json
NP_BENCH_VERBOSE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2i_bench_n2.json 2> gpurun_out/r2i_bench_n2.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r2i_bench_n2.err; wc -c gpurun_out/r2i_bench_n2.json
