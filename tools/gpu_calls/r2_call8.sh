#!/bin/bash
# 2 GPUs: the pure C++ program over the library's NCCL path, then the bench under torchrun
mkdir -p gpurun_out
./tests/cpp/comm_two_ranks.bin netphys_b200/libnetphys_b200.so 2 4000 6 > gpurun_out/r2h_cpp.log 2>&1; echo "cpp rc=$?"; tail -5 gpurun_out/r2h_cpp.log
python -m pytest tests/test_gpu_comm_cpp.py tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2h_bench_n2.json 2> gpurun_out/r2h_bench_n2.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r2h_bench_n2.err
