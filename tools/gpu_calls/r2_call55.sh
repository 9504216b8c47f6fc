#!/bin/bash
# four GPUs: the C++ program with four ranks (three peers per rank), then the headline with the copy-engine gather
mkdir -p gpurun_out
g++ -std=c++17 -O1 -o /tmp/comm_ranks tests/cpp/comm_two_ranks.cpp -ldl && timeout 300 /tmp/comm_ranks netphys_b200/libnetphys_b200.so 4 4000 6; echo "cpp rc=$?"
NP_BENCH_VERBOSE=1 NP_BENCH_SHARDED_BODIES_PER_GPU=100000 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r4i_bench_n4.json 2> gpurun_out/r4i_bench_n4.err; echo "bench n=4 rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r4i_bench_n4.json') if l.startswith('{')][0])
print("value %.4g ms %.4f per_rank %s phases %s" % (d['value'], d['ms_per_step'], [round(x,3) for x in d['per_rank_ms']], {k: round(v,4) for k,v in d['phase_ms'].items()}))
print(d['method']['collective'][:90])
PY
