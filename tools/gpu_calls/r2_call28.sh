#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/prof_world.py 1000000 6 c3 2>&1 | tail -1 | cut -c1-170
python tools/prof_world.py 131072 20 c2 2>&1 | tail -1 | cut -c1-170
python tools/prof_pairs.py mix 2000000 3 2>&1 | tail -1 | cut -c1-100
python tools/prof_pairs.py box 2000000 3 2>&1 | tail -1 | cut -c1-100
python tools/prof_pairs.py sphere 1000000 3 2>&1 | tail -1 | cut -c1-100
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2y_launches_1m.csv python tools/prof_world.py 1000000 2 c3 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2y_launches_1m.csv')) if len(r)>5]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value')
for r in rows[-11:]: print(r[ik][:70], r[iv])
PY
