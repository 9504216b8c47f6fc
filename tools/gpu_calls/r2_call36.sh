#!/bin/bash
mkdir -p gpurun_out
NP_BENCH_VERBOSE=1 NP_BENCH_BODIES=200000 NP_BENCH_SHARDED_BODIES_PER_GPU=200000 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r3f_bench_n2_small.json 2> gpurun_out/r3f_bench_n2_small.err; echo "rc=$?"; tail -c 1200 gpurun_out/r3f_bench_n2_small.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r3f_bench_n2_small.json') if l.startswith('{')][0])
s=d['other_configs']['sharded_world']; print(s['ms_per_step'], s['phase_ms'], s['state_bit_identical_on_all_ranks']); print(json.dumps(s['partitioned'], indent=1)[:1500])
PY
