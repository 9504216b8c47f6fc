#!/bin/bash
for l in "" build/lib_e2t64.so; do echo "== lib $l"
  if [ -n "$l" ]; then export NETPHYS_B200_LIB=$PWD/$l; fi
  python tools/prof_world.py 1000000 6 c3 2>&1 | tail -1 | cut -c1-150
  python tools/prof_pairs.py mix 2000000 3 2>&1 | tail -1 | cut -c1-100
  python tools/prof_pairs.py sphere 1000000 3 2>&1 | tail -1 | cut -c1-100
done
