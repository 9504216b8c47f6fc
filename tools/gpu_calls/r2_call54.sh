#!/bin/bash
# eight GPUs: the headline (one 1M-body world per rank + snapshot gather) with the copy-engine gather and with ncclAllGather
mkdir -p gpurun_out
for mode in 0 1; do
  NP_GATHER_NCCL=$mode NP_BENCH_VERBOSE=1 NP_BENCH_SHARDED_BODIES_PER_GPU=100000 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2952$mode bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r4h_bench_n8_nccl$mode.json 2> gpurun_out/r4h_bench_n8_nccl$mode.err; echo "bench n=8 NP_GATHER_NCCL=$mode rc=$?"; grep "timed region" gpurun_out/r4h_bench_n8_nccl$mode.err | head -8
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r4h_bench_n8_nccl$mode.json') if l.startswith('{')][0])
print("value %.4g ms %.4f per_rank %s phases %s" % (d['value'], d['ms_per_step'], [round(x,3) for x in d['per_rank_ms']], {k: round(v,4) for k,v in d['phase_ms'].items()}))
print(d['method']['collective'][:90])
PY
done
