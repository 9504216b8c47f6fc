#!/bin/bash
for n in 10000 20000 40000 80000 131072 200000; do
  a=$(python tools/prof_world.py $n 20 c2 2>&1 | tail -1 | grep -o "narrow resolve snap total\] \[[^]]*\]")
  b=$(NP_G_GJK=1 NP_G_EPA=4 python tools/prof_world.py $n 20 c2 2>&1 | tail -1 | grep -o "narrow resolve snap total\] \[[^]]*\]")
  c=$(NP_G_GJK=1 NP_G_EPA=4 python tools/prof_world.py $n 20 c3 2>&1 | tail -1 | grep -o "narrow resolve snap total\] \[[^]]*\]")
  d=$(python tools/prof_world.py $n 20 c3 2>&1 | tail -1 | grep -o "narrow resolve snap total\] \[[^]]*\]")
  echo "$n c2 auto: $a"; echo "$n c2 thread: $b"; echo "$n c3 auto: $d"; echo "$n c3 thread: $c"
done
