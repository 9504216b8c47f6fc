#!/bin/bash
mkdir -p gpurun_out
for e in 4 1; do
  echo "== NP_G_EPA=$e"
  NP_G_EPA=$e python tools/prof_world.py 1000000 6 c3 2>&1 | tail -1 | cut -c1-160
  NP_G_EPA=$e python tools/prof_pairs.py mix 2000000 3 2>&1 | tail -1 | cut -c1-200
  NP_G_EPA=$e python tools/prof_pairs.py box 2000000 3 2>&1 | tail -1 | cut -c1-200
  NP_G_EPA=$e python tools/prof_pairs.py sphere 1000000 3 2>&1 | tail -1 | cut -c1-200
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2o_launches_1m.csv python tools/prof_world.py 1000000 3 c3 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2o_launches_1m.csv')) if len(r)>5]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value')
for r in rows[-16:]: print(r[ik][:70], r[iv])
PY
