#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for t in 1 0; do
  echo "== NP_GJK2=$t"
  NP_GJK2=$t python tools/prof_world.py 1000000 6 c3 2>&1 | tail -1 | cut -c1-200
  NP_GJK2=$t python tools/prof_world.py 300000 6 c3 2>&1 | tail -1 | cut -c1-200
  NP_GJK2=$t python tools/prof_pairs.py mix 2000000 3 2>&1 | tail -1 | cut -c1-200
  NP_GJK2=$t python tools/prof_pairs.py box 2000000 3 2>&1 | tail -1 | cut -c1-200
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2q_launches_1m.csv python tools/prof_world.py 1000000 2 c3 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2q_launches_1m.csv')) if len(r)>5]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value')
for r in rows[-9:]: print(r[ik][:70], r[iv])
PY
