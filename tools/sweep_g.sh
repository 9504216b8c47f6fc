#!/bin/bash
# G sweep of the narrow-phase kernels on fixed pair workloads (tuning aid)
for mix in box mix sphere; do
  for g in 1 4 8 32; do
    for e in 1 4 8 32; do
      echo -n "mix=$mix G_gjk=$g G_epa=$e : "
      NP_G_GJK=$g NP_G_EPA=$e timeout 120 python tools/prof_pairs.py $mix ${1:-1000000} 3 2>&1 | tail -1
    done
  done
done
