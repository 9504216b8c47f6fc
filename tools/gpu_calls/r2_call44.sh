#!/bin/bash
python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py tests/test_gpu_ode_shim.py -x -q 2>&1 | tail -3
python tools/prof_grid.py 1000000 2.0 2>&1 | tail -2 | cut -c1-200
python tools/prof_world.py 10000 30 c2 2>&1 | tail -1 | cut -c1-150
NP_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r3m_grid_launches.csv python tools/prof_grid.py 1000000 2.0 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r3m_grid_launches.csv')) if len(r)>5]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value')
L=[(r[ik].split('(')[0].replace('void ','')[:40], float(r[iv])/1e3) for r in rows[1:]]
for n,v in L[-19:-9]: print("%-42s %8.1f us"%(n,v))
PY
