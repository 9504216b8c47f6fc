#!/bin/bash
# A/B of tuning builds: world 10k (C2), world 1M (C3 mix), pairs box / mix; then overflow counters
for lib in "$@"; do
  echo "== $lib"
  export NETPHYS_B200_LIB=$PWD/netphys_b200/$lib
  python tools/prof_world.py 10000 6 c2 | tail -1
  python tools/prof_world.py 1000000 4 c3 | tail -1
  python tools/prof_pairs.py box 1000000 3 | tail -1
  python tools/prof_pairs.py mix 1000000 3 | tail -1
  NP_COUNT=1 python tools/prof_pairs.py mix 1000000 1 | tail -1 | grep -o "'hits': [0-9]*\|'epa_overflow': [0-9]*" | paste - -
done
