#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for hs in 1.0 1.25 1.5 2.0 2.5; do echo "h_scale $hs"; NP_GRID_H_SCALE=$hs python tools/prof_grid.py 1000000 2.0 gridonly 2>&1 | tail -1 | cut -c1-230; done
echo "hashed"; NP_GRID_HASH=1 python tools/prof_grid.py 1000000 2.0 gridonly 2>&1 | tail -1 | cut -c1-230
NP_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r4a_grid_launches.csv python tools/prof_grid.py 1000000 2.0 gridonly > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r4a_grid_launches.csv')) if len(r)>5]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value')
L=[(r[ik].split('(')[0].replace('void ','')[:40], float(r[iv])/1e3) for r in rows[1:]]
for n,v in L[-21:]: print("%-42s %8.1f us"%(n,v))
PY
