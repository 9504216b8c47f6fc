#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for t in 5 9 17; do
  echo "== NP_TAB_MIN=$t"
  NP_TAB_MIN=$t python tools/prof_world.py 1000000 6 c3 2>&1 | tail -1 | cut -c1-170
  NP_TAB_MIN=$t python tools/prof_pairs.py mix 2000000 3 2>&1 | tail -1 | cut -c1-100
  NP_TAB_MIN=$t python tools/prof_pairs.py box 2000000 3 2>&1 | tail -1 | cut -c1-100
done
for e in 1 4; do echo "== NP_G_EPA=$e";  NP_G_EPA=$e python tools/prof_world.py 1000000 6 c3 2>&1 | tail -1 | cut -c1-170;  NP_G_EPA=$e python tools/prof_pairs.py box 2000000 3 2>&1 | tail -1 | cut -c1-100; NP_G_EPA=$e python tools/prof_pairs.py mix 2000000 3 2>&1 | tail -1 | cut -c1-100; done
