#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for t in 32 4; do
  echo "== NP_EPA2_TAIL=$t"
  NP_EPA2_TAIL=$t python tools/prof_world.py 1000000 6 c3 2>&1 | tail -1 | cut -c1-170
  NP_EPA2_TAIL=$t python tools/prof_pairs.py mix 2000000 3 2>&1 | tail -1 | cut -c1-100
  NP_EPA2_TAIL=$t python tools/prof_pairs.py box 2000000 3 2>&1 | tail -1 | cut -c1-100
done
ncu --set full --clock-control none --import-source on -k regex:"k_epa2" -c 1 -o gpurun_out/r2w_world1m -f python tools/prof_world.py 1000000 1 c3 > gpurun_out/r2w_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2w_launches_1m.csv python tools/prof_world.py 1000000 2 c3 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2w_launches_1m.csv')) if len(r)>5]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value')
for r in rows[-11:]: print(r[ik][:70], r[iv])
PY
