#!/bin/bash
python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
for v in 1 0; do echo "== NP_TABC=$v"
for n in 10000 20000 40000; do NP_TABC=$v python tools/prof_world.py $n 30 c2 2>&1 | tail -1 | cut -c1-150; done
NP_TABC=$v python tools/prof_world.py 10000 30 c3 2>&1 | tail -1 | cut -c1-150
done
