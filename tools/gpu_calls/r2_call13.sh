#!/bin/bash
for n in 40000 65536 131072 262144; do
  for cfg in "NP_X=0" "NP_G_GJK=1 NP_G_EPA=4" "NP_G_GJK=1 NP_G_EPA=32" "NP_G_GJK=8 NP_G_EPA=4"; do
    echo "== n=$n $cfg"; env $cfg python tools/prof_world.py $n 12 c2 2>&1 | tail -1 | cut -c1-150
  done
done
