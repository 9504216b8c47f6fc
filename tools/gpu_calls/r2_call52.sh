#!/bin/bash
for e in "" "NP_G_EPA=4" "NP_G_EPA=1" "NP_G_EPA=8" "NP_G_GJK=1 NP_G_EPA=4"; do
  echo "== $e"; env $e python tools/prof_world.py 10000 30 c2 2>&1 | tail -1 | cut -c1-150
done
for e in "" "NP_G_EPA=4"; do
  echo "== 20k $e"; env $e python tools/prof_world.py 20000 30 c2 2>&1 | tail -1 | cut -c1-150
  echo "== 40k $e"; env $e python tools/prof_world.py 40000 30 c2 2>&1 | tail -1 | cut -c1-150
done
