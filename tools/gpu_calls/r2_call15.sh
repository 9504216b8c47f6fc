#!/bin/bash
# support-direction tables: correctness first, then A/B on the headline world, the C2 world and the pair lists
mkdir -p gpurun_out
python -m pytest tests/test_gpu_suptab.py -x -q 2>&1 | tail -15
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for t in 1 0; do
  echo "== NP_TABLES=$t"
  NP_TABLES=$t python tools/prof_world.py 1000000 6 c3 2>&1 | tail -1 | cut -c1-200
  NP_TABLES=$t python tools/prof_world.py 10000 30 c2 2>&1 | tail -1 | cut -c1-200
  NP_TABLES=$t python tools/prof_pairs.py mix 2000000 3 2>&1 | tail -1 | cut -c1-200
  NP_TABLES=$t python tools/prof_pairs.py box 2000000 3 2>&1 | tail -1 | cut -c1-200
  NP_TABLES=$t python tools/prof_pairs.py sphere 1000000 3 2>&1 | tail -1 | cut -c1-200
done
