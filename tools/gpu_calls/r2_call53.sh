#!/bin/bash
# two GPUs: the C++ two-rank program (copy-engine gather and NCCL gather), the sharded tests, the full-size bench with both gathers
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_comm_cpp.py tests/test_gpu_sharded.py -x -q 2>&1 | tail -5
for mode in 0 1; do
  NP_GATHER_NCCL=$mode NP_BENCH_VERBOSE=1 NP_BENCH_SHARDED_BODIES_PER_GPU=200000 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$mode bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r4g_bench_n2_nccl$mode.json 2> gpurun_out/r4g_bench_n2_nccl$mode.err; echo "bench n=2 NP_GATHER_NCCL=$mode rc=$?"; tail -c 300 gpurun_out/r4g_bench_n2_nccl$mode.err
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r4g_bench_n2_nccl$mode.json') if l.startswith('{')][0])
print("value %.4g ms %.4f per_rank %s phases %s" % (d['value'], d['ms_per_step'], d['per_rank_ms'], d['phase_ms']))
print(d['method']['collective'][:90])
PY
done
