#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/prof_grid.py 1000000 2.0 2>&1 | tail -2 | cut -c1-230
python tools/prof_world.py 10000 30 c2 2>&1 | tail -1 | cut -c1-150
NP_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:"k_world_verts" -s 3 -c 1 -o gpurun_out/r4b_wv -f python tools/prof_grid.py 1000000 2.0 gridonly > gpurun_out/r4b_ncu.log 2>&1
NP_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r4b_grid_launches.csv python tools/prof_grid.py 1000000 2.0 gridonly > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r4b_grid_launches.csv')) if len(r)>5]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value')
L=[(r[ik].split('(')[0].replace('void ','')[:40], float(r[iv])/1e3) for r in rows[1:]]
for n,v in L[-18:]: print("%-42s %8.1f us"%(n,v))
PY
