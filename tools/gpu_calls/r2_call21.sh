#!/bin/bash
mkdir -p gpurun_out
for b in 4 5 6; do
  echo "== minb $b"
  NETPHYS_B200_LIB=$PWD/build/lib_minb$b.so python tools/prof_world.py 1000000 6 c3 2>&1 | tail -1 | cut -c1-170
  NETPHYS_B200_LIB=$PWD/build/lib_minb$b.so python tools/prof_pairs.py mix 2000000 3 2>&1 | tail -1 | cut -c1-100
done
