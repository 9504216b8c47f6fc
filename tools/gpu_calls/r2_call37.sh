#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
ncu --metrics smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:"k_gjk2|k_epa|k_integrate|k_resolve" -c 8 --csv --log-file gpurun_out/r2_fp32_mix.csv python tools/prof_world.py 1000000 1 c3 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_fp32_mix.csv')) if len(r)>5]
hdr=rows[0]; ik=hdr.index('Kernel Name'); im=hdr.index('Metric Name'); iv=hdr.index('Metric Value')
for r in rows[1:]: print(r[ik].split('(')[0][:40], r[im], r[iv])
PY
